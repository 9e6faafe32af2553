"""Pins the CPU oracle (and the host Graph mirror) against the reference's own known-answer tests."""
import numpy as np
import pytest

import fdgraph_b200 as fd
from oracle import oracle as O


def _emit(graphs, leaf, root=None):
    raw, nodes = fd.flatten(graphs, root)
    orc = O.Oracle(raw)
    out = orc.eval(np.asarray(leaf, np.float64).reshape(-1, 1), mode="emitter")
    return orc, out[:, 0], nodes


def test_compile_directly_4p5():
    # reference test/compiler.jl:2-16: FeynmanGraph([ext_vertex, ext_vertex]; factor=1.5), leaf=[1,2] -> 4.5
    v1, v2 = fd.FeynmanGraph([]), fd.FeynmanGraph([])
    g = fd.FeynmanGraph([v1, v2], factor=1.5)
    orc, out, nodes = _emit([g], [1.0, 2.0])
    assert out[0] == 4.5 == (1.0 + 2.0) * 1.5
    # leaf numbering: first visited leaf is column 0 (static.jl:117-119); return value = last root (static.jl:127)
    assert [nodes[i] for i in orc.leaf_nodes] == [v1, v2]
    assert orc.last_root == 0
    # the interpreter agrees (eval.jl)
    assert O.eval_interp_py(g, {v1.id: 0, v2.id: 1}, [1.0, 2.0]) == 4.5


def test_eval_26_27_702():
    # reference test/computational_graph.jl:874-887
    g1 = fd.Graph([])
    g2 = fd.Graph([], factor=2)
    g3 = 2 * (3 * g1 + 5 * g2)
    g4 = g1 + 2 * (3 * g1 + 5 * g2)
    g5 = g4 * g3
    assert O.eval_interp_py(g3) == 26
    assert O.eval_interp_py(g4) == 27
    assert O.eval_interp_py(g5) == 27 * 26
    orc, out, _ = _emit([g3, g4, g5], [1.0, 1.0])
    assert list(out) == [26.0, 27.0, 702.0]
    for mode in ("emitter", "interp"):
        r = orc.eval(np.ones((2, 5)), mode=mode)
        assert (r == np.array([[26.0], [27.0], [702.0]])).all()


def test_graph_constructor_conventions():
    # graph.jl:69-73 factor wrapping, :136-147 scalar product merges trivial unary chains
    g1 = fd.Graph([])
    w = fd.Graph([], factor=2)
    assert isinstance(w.operator, fd.Prod) and w.subgraph_factors == [2.0] and w.subgraphs[0].isleaf()
    h = 5 * w
    assert h.subgraphs[0] is w.subgraphs[0] and h.subgraph_factors == [10.0]
    s = g1 + g1  # linear_combination merges equal ids (graph.jl:199-201)
    assert len(s.subgraphs) == 1 and s.subgraph_factors == [2.0]
    p = g1 * g1  # multi_product -> Power(2) (graph.jl:318-319)
    assert isinstance(p.operator, fd.Power) and p.operator.N == 2
    with pytest.raises(AssertionError):
        fd.Power(1)
    lc = fd.linear_combination([g1, w, g1], [2, 1, 3])  # vector form merges duplicates, unwraps w
    assert [x.id for x in lc.subgraphs] == [g1.id, w.subgraphs[0].id] and lc.subgraph_factors == [5.0, 2.0]
    mp = fd.multi_product([g1, g1, w])
    assert isinstance(mp.operator, fd.Prod) and isinstance(mp.subgraphs[0].operator, fd.Power)


def test_root_semantics():
    # static.jl:111-114,126-128: only ids in `root` are written, at the first position of the id;
    # the return value is the last root assigned.
    a, b = fd.Graph([]), fd.Graph([])
    s = fd.Graph([a, b], operator=fd.Sum(), subgraph_factors=[2.0, 3.0])
    p = fd.Graph([s, a], operator=fd.Prod())
    raw, _ = fd.flatten([p], root=[s.id, 12345, p.id, s.id])
    orc = O.Oracle(raw)
    out = orc.eval(np.array([[2.0], [5.0]]), root=np.full((4, 1), -7.0))
    assert list(out[:, 0]) == [19.0, -7.0, 38.0, -7.0]
    assert orc.last_root == 2


def test_dedupe_by_id_first_visit_wins():
    # two distinct leaf objects with the same id are one variable g<ID> (static.jl:116,122)
    a = fd.Graph([])
    a2 = fd.Graph([])
    a2.id = a.id
    s = fd.Graph([a, a2], operator=fd.Sum())
    raw, nodes = fd.flatten([s])
    orc = O.Oracle(raw)
    assert orc.n_leaves == 1
    assert orc.eval(np.array([[3.0]]))[0, 0] == 6.0


def test_emitter_vs_interpreter_rounding_differs_for_prod():
    # eval! computes prod(g_i*f_i), the emitter ((g1*f1)*g2)*f2: same value, maybe different bits
    rng = np.random.default_rng(0)
    a, b, c = fd.Graph([]), fd.Graph([]), fd.Graph([])
    p = fd.Graph([a, b, c], operator=fd.Prod(), subgraph_factors=[1.0 / 3.0, 0.7, 1.1])
    raw, _ = fd.flatten([p])
    orc = O.Oracle(raw)
    leaf = 0.5 + rng.random((3, 4096))
    e, i = orc.eval(leaf, "emitter"), orc.eval(leaf, "interp")
    assert np.allclose(e, i, rtol=1e-14)
    assert (e != i).any()
    l0 = leaf[:, 0]
    assert e[0, 0] == ((l0[0] * (1.0 / 3.0)) * l0[1]) * 0.7 * l0[2] * 1.1
    assert i[0, 0] == (l0[0] * (1.0 / 3.0)) * (l0[1] * 0.7) * (l0[2] * 1.1)


def test_exact_rational_agrees():
    import graphgen

    roots = graphgen.random_dag(3, n_leaves=5, n_inner=25, n_roots=3)
    raw, _ = fd.flatten(roots)
    orc = O.Oracle(raw)
    leaf = graphgen.leaf_values(1, orc.n_leaves, 4, signed=True)
    out = orc.eval(leaf)
    for b in range(4):
        exact = O.eval_exact(orc, leaf[:, b])
        bound = O.eval_abs_bound(orc, leaf[:, b])
        for r in range(orc.n_roots):
            assert abs(out[r, b] - float(exact[r])) <= 64 * 2.3e-16 * bound[r] * orc.n_stmts


# ---- Taylor-mode AD (BASELINE config 5): the restated front end against the reference's own known answers ------------


def test_taylor_series_known_answers():
    # reference test/taylor.jl:42-56
    from oracle.frontend import taylor as T

    a, b, c, d, e = T.set_variables([3, 3, 3, 3, 3])
    F1 = (a + b) * (a + b) * (a + b)
    assert F1.coeffs[(2, 1, 0, 0, 0)] == 3.0 and F1.coeffs[(1, 2, 0, 0, 0)] == 3.0
    assert F1.coeffs[(3, 0, 0, 0, 0)] == 1.0 and F1.coeffs[(0, 3, 0, 0, 0)] == 1.0
    F2 = (1 + a) * (3 + 2 * c)
    assert F2.coeffs[(0, 0, 0, 0, 0)] == 3.0 and F2.coeffs[(1, 0, 0, 0, 0)] == 3.0
    assert F2.coeffs[(0, 0, 1, 0, 0)] == 2.0 and F2.coeffs[(1, 0, 1, 0, 0)] == 2.0
    F3 = (a + b) ** 3
    assert F3.coeffs[(2, 1, 0, 0, 0)] == 3.0 and F3.coeffs[(0, 3, 0, 0, 0)] == 1.0
    assert (4, 0, 0, 0, 0) not in ((a + b) ** 4).coeffs  # truncated at the maximal orders (arithmetic.jl:175)


def _graphs_from_raw(d):
    """Graph objects from flattened arrays (children come before parents in fd.flatten's order)."""
    nodes = []
    ops = {0: fd.Unitary(), 1: fd.Sum(), 2: fd.Prod()}
    for i in range(len(d["node_id"])):
        lo, hi = d["child_ptr"][i], d["child_ptr"][i + 1]
        subs = [nodes[c] for c in d["child_node"][lo:hi]]
        op = fd.Power(d["node_pow"][i]) if d["node_op"][i] == 3 else ops[d["node_op"][i]]
        g = fd.Graph(subs, subgraph_factors=d["child_factor"][lo:hi], operator=op if subs else fd.Sum())
        nodes.append(g)
    return nodes, [nodes[i] for i in d["graphs"]]


def test_taylor_expansion_equals_counterterm_diagram_files():
    """reference test/taylor.jl:96-112: the Taylor coefficients [g, v] of the two order-2 self-energy graphs equal the
    diagrams of the counter-term files Sigma2_<v>_<g>.diag (all leaves one).  Fixture: tools/gen_golden_taylor.py."""
    import json
    import os

    from oracle.frontend import taylor as T

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "taylor_sigma2.json")) as fh:
        gold = json.load(fh)
    nodes, roots = _graphs_from_raw(gold["graph"])
    var = {n.id: [k == "G", k == "W"] for n, k in zip(nodes, gold["node_kind"]) if n.isleaf()}
    series, _ = T.taylorexpansion(roots, var, [2, 2])
    assert len(gold["expected"]) == 8
    for key, want in gold["expected"].items():
        g, v = (int(x) for x in key.split(","))
        coeffs = [series[0].coeffs[(g, v)], series[1].coeffs[(g, v)]]
        raw, _ = fd.flatten(coeffs)
        orc = O.Oracle(raw)
        got = orc.eval(np.ones((orc.n_leaves, 1)))[:, 0]
        assert list(got) == want, (key, list(got), want)
        assert [O.eval_interp_py(c) for c in coeffs] == want  # eval! with the default leaf weight... every leaf 1.0


def test_taylor_workload_checksums():
    """The committed config-5 graphs: all-leaves-one values of the 6 x 3 coefficient roots of the order-3 self-energy
    scale with the number of ways to place the counter-term orders (binomial pattern of the order-(0,0) values)."""
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    man = json.load(open(os.path.join(root, "workloads", "MANIFEST.json")))
    raw = fd.RawGraph.load(os.path.join(root, "workloads", "taylor_sigma_o3.npz"))
    orc = O.Oracle(raw)
    ones = orc.eval(np.ones((orc.n_leaves, 1)))[:, 0]
    assert list(ones) == man["taylor_sigma_o3"]["all_leaves_one"]
    base = ones[:3]
    # order-3 self-energy: 5 propagators, 3 interactions -> (0,1): C(3,1)=3, (1,0): C(5,1)=5, (2,0): C(6,2)=15, products for mixed
    for k, mult in enumerate([1, 3, 5, 15, 15, 45]):
        assert list(ones[3 * k:3 * k + 3]) == list(mult * base)


def test_derivative_known_answers_through_the_taylor_series():
    """reference test/computational_graph.jl:930-1071 (`forwardAD_root!`, `build_derivative_graph`): explicit leaf vectors
    -> 120 / 5 / 1, 570 / 3 / 1, 120 / 2 / 0, 300, 3840, 480, 1002, 426, 90, 0, 5568, 1003, 3708, 1638, 234.  The
    reference's graph-level AD is restated and pinned further down; here the same numbers are pinned on the restated Taylor-series expansion, which
    is what the Taylor-AD workloads (BASELINE config 5) are made with: the derivative of order n is n! times the Taylor
    coefficient, a leaf's own first derivative is the seed of the test's leaf vector and its higher derivatives are 0."""
    import math

    import fdgraph_b200 as fd
    from oracle.frontend import taylor as T

    fd.uidreset()
    g1, g2 = fd.Graph([]), fd.Graph([])
    g3 = fd.Graph([], factor=2.0)
    l3 = g3.eldest()
    F3 = g1 + g2
    F2 = fd.graph.linear_combination([g1, g3, F3], [2, 1, 3])
    F1 = fd.Graph([g1, F2, F3], operator=fd.Prod(), subgraph_factors=[3.0, 1.0, 1.0])
    F0 = F1 * F3
    F0_r1 = F1 + F3
    dep = {g1.id: [True, False, False], g2.id: [False, True, False], l3.id: [False, False, True]}
    (s1, s2, s3, s0, s0r), _ = T.taylorexpansion([F1, F2, F3, F0, F0_r1], dep, [3, 2, 2])
    var_of = {g1.id: 0, g2.id: 1, l3.id: 2}

    def derivative(series, order, leaf):
        """leaf = [g1, g2, l3, seed1, seed2, seed3] as in the reference's leaf vectors"""
        val = {}

        def leaf_value(node):
            orders = tuple(getattr(node, "orders", None) or (0, 0, 0))
            if sum(orders) == 0:
                return leaf[var_of[node.id]]
            v = next(i for i, o in enumerate(orders) if o)
            return leaf[3 + v] if sum(orders) == 1 else 0.0

        g = series.coeffs.get(order)
        if g is None:
            return 0.0
        for node in fd.graph.post_order_unique([g]):
            if not node.subgraphs:
                val[id(node)] = leaf_value(node)
                continue
            terms = [val[id(s)] * f for s, f in zip(node.subgraphs, node.subgraph_factors)]
            val[id(node)] = sum(terms) if isinstance(node.operator, fd.Sum) else math.prod(terms)
        return val[id(g)] * math.prod(math.factorial(o) for o in order)

    x, y, z = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    # forwardAD_root!: first derivatives along g1, g2, eldest(g3)
    leaf = [1.0, 1.0, 1.0, 1.0, 0.0, 0.0]
    assert (derivative(s1, x, leaf), derivative(s2, x, leaf), derivative(s3, x, leaf)) == (120.0, 5.0, 1.0)
    leaf = [5.0, -1.0, 2.0, 0.0, 1.0, 0.0]
    assert (derivative(s1, y, leaf), derivative(s2, y, leaf), derivative(s3, y, leaf)) == (570.0, 3.0, 1.0)
    leaf = [5.0, -1.0, 2.0, 0.0, 0.0, 1.0]
    assert (derivative(s1, z, leaf), derivative(s2, z, leaf), derivative(s3, z, leaf)) == (120.0, 2.0, 0.0)
    assert derivative(s0, x, [1.0, 1.0, 1.0, 1.0, 0.0, 0.0]) == 300.0
    assert derivative(s0, y, [5.0, -1.0, 2.0, 0.0, 1.0, 0.0]) == 3840.0
    assert derivative(s0, z, [5.0, -1.0, 2.0, 0.0, 0.0, 1.0]) == 480.0
    assert derivative(s0r, z, [5.0, -1.0, 2.0, 0.0, 0.0, 1.0]) == 120.0
    # build_derivative_graph: orders up to (3, 2, 2), every first derivative seeded with 1
    leaf = [5.0, -1.0, 2.0, 1.0, 1.0, 1.0]
    assert [derivative(s1, o, leaf) for o in ((1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 1, 0))] == [1002.0, 426.0, 90.0, 0.0]
    assert [derivative(s0, o, leaf) for o in ((1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 1, 0), (3, 2, 0))] == [5568.0, 3708.0, 1638.0, 234.0, 0.0]
    assert [derivative(s0r, o, leaf) for o in ((1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 1, 0), (3, 2, 0))] == [1003.0, 426.0, 90.0, 0.0, 0.0]


def _eval_by_id(graph, leafmap, leaf):
    """Graphs.eval!(g, leafmap, leaf) (eval.jl:15-39) on Graph objects: a node without subgraphs reads leaf[leafmap[id]]."""
    import math

    import fdgraph_b200 as fd

    val = {}
    for node in fd.graph.post_order_unique([graph]):
        if not node.subgraphs:
            val[id(node)] = leaf[leafmap[node.id] - 1]
            continue
        terms = [val[id(s)] * f for s, f in zip(node.subgraphs, node.subgraph_factors)]
        if isinstance(node.operator, fd.Sum):
            val[id(node)] = sum(terms)
        elif isinstance(node.operator, fd.Prod):
            val[id(node)] = math.prod(terms)
        else:
            val[id(node)] = terms[0] ** node.operator.N
    return val[id(graph)]


def _compiled_by_id(graphs, leafmap, leaf):
    """The same values through flatten + the C oracle in emitter order (what Compilers.compile would evaluate)."""
    import fdgraph_b200 as fd

    raw, nodes = fd.flatten(list(graphs))
    orc = O.Oracle(raw)
    vec = np.array([[leaf[leafmap[nodes[i].id] - 1]] for i in orc.leaf_nodes], dtype=np.float64).reshape(orc.n_leaves, 1)
    return [float(x) for x in orc.eval(vec)[:, 0]]


def test_forwardAD_root_known_answers():
    """reference test/computational_graph.jl:930-987, with the restated forwardAD_root! (oracle/frontend/ad.py)."""
    import fdgraph_b200 as fd
    from oracle.frontend import ad

    fd.uidreset()
    g1, g2 = fd.Graph([]), fd.Graph([])
    g3 = fd.Graph([], factor=2.0)
    F3 = g1 + g2
    F2 = fd.graph.linear_combination([g1, g3, F3], [2, 1, 3])
    F1 = fd.Graph([g1, F2, F3], operator=fd.Prod(), subgraph_factors=[3.0, 1.0, 1.0])
    k = (True,)
    kg1, kg2, kg3 = (g1.id, k), (g2.id, k), (g3.eldest().id, k)
    kF1, kF2, kF3 = (F1.id, k), (F2.id, k), (F3.id, k)
    dual = ad.forwardAD_root([F1])
    assert dual[kF3].subgraphs == [dual[kg1], dual[kg2]]
    assert dual[kF2].subgraphs == [dual[kg1], dual[kg3], dual[kF3]]
    assert all(d.name != ad.UNDEFINED for d in dual.values())
    leafmap = {g1.id: 1, g2.id: 2, g3.eldest().id: 3, dual[kg1].id: 4, dual[kg2].id: 5, dual[kg3].id: 6}
    for leaf, want in (([1.0, 1.0, 1.0, 1.0, 0.0, 0.0], (120.0, 5.0, 1.0)), ([5.0, -1.0, 2.0, 0.0, 1.0, 0.0], (570.0, 3.0, 1.0)),
                       ([5.0, -1.0, 2.0, 0.0, 0.0, 1.0], (120.0, 2.0, 0.0))):
        assert tuple(_eval_by_id(dual[key], leafmap, leaf) for key in (kF1, kF2, kF3)) == want
        assert tuple(_compiled_by_id([dual[kF1], dual[kF2], dual[kF3]], leafmap, leaf)) == want
    F0 = F1 * F3
    kF0 = (F0.id, k)
    dual1 = ad.forwardAD_root([F0])
    leafmap.update({dual1[kg1].id: 4, dual1[kg2].id: 5, dual1[kg3].id: 6})
    assert _eval_by_id(dual1[kF0], leafmap, [1.0, 1.0, 1.0, 1.0, 0.0, 0.0]) == 300.0
    assert _eval_by_id(dual1[kF0], leafmap, [5.0, -1.0, 2.0, 0.0, 1.0, 0.0]) == 3840.0
    leaf = [5.0, -1.0, 2.0, 0.0, 0.0, 1.0]
    assert _eval_by_id(dual1[kF0], leafmap, leaf) == 480.0
    F0_r1 = F1 + F3
    dual2 = ad.forwardAD_root([F0, F0_r1])
    leafmap.update({dual2[kg1].id: 4, dual2[kg2].id: 5, dual2[kg3].id: 6})
    assert _eval_by_id(dual2[kF0], leafmap, leaf) == 480.0
    assert _eval_by_id(dual2[(F0_r1.id, k)], leafmap, leaf) == 120.0
    assert _compiled_by_id([dual2[kF0], dual2[(F0_r1.id, k)]], leafmap, leaf) == [480.0, 120.0]


def test_build_derivative_graph_known_answers():
    """reference test/computational_graph.jl:988-1071: orders (3, 2, 2), before and after burn_from_targetleaves!."""
    import itertools

    import fdgraph_b200 as fd
    from oracle.frontend import ad

    fd.uidreset()
    g1, g2 = fd.Graph([]), fd.Graph([])
    g3 = fd.Graph([], factor=2.0)
    l3 = g3.eldest()
    F3 = g1 + g2
    F2 = fd.graph.linear_combination([g1, g3, F3], [2, 1, 3])
    F1 = fd.Graph([g1, F2, F3], operator=fd.Prod(), subgraph_factors=[3.0, 1.0, 1.0])
    orders = (3, 2, 2)
    leaf = [5.0, -1.0, 2.0, 1.0, 1.0, 1.0, 0.0]

    def leafmap_of(dual):
        leafmap = {g1.id: 1, g2.id: 2, l3.id: 3, dual[(g1.id, (1, 0, 0))].id: 4, dual[(g2.id, (0, 1, 0))].id: 5,
                   dual[(l3.id, (0, 0, 1))].id: 6}
        burn = []
        for order in itertools.product(*[range(o + 1) for o in orders]):
            if order == (0, 0, 0):
                continue
            for g in (g1, g2, l3):
                if dual[(g.id, order)].id not in leafmap:
                    leafmap[dual[(g.id, order)].id] = 7
                    burn.append(dual[(g.id, order)].id)
        return leafmap, burn

    dual = ad.build_derivative_graph(F1, orders)
    leafmap, burn = leafmap_of(dual)
    keys = [(F1.id, o) for o in ((1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 1, 0))]
    want = [1002.0, 426.0, 90.0, 0.0]
    assert [_eval_by_id(dual[key], leafmap, leaf) for key in keys] == want
    assert _compiled_by_id([dual[key] for key in keys], leafmap, leaf) == want
    c0 = ad.burn_from_targetleaves([dual[key] for key in keys], burn)
    if c0 is not None:
        leafmap[c0] = 7
    assert [_eval_by_id(dual[key], leafmap, leaf) for key in keys] == want
    # a vector of graphs
    F0 = F1 * F3
    F0_r1 = F1 + F3
    dual = ad.build_derivative_graph([F0, F0_r1], orders)
    leafmap, burn = leafmap_of(dual)
    olist = ((1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 1, 0), (3, 2, 0))
    want0, want1 = [5568.0, 3708.0, 1638.0, 234.0, 0.0], [1003.0, 426.0, 90.0, 0.0, 0.0]
    assert [_eval_by_id(dual[(F0.id, o)], leafmap, leaf) for o in olist] == want0
    assert [_eval_by_id(dual[(F0_r1.id, o)], leafmap, leaf) for o in olist] == want1
    roots = [dual[(F0.id, o)] for o in olist] + [dual[(F0_r1.id, o)] for o in olist]
    assert _compiled_by_id(roots, leafmap, leaf) == want0 + want1
    c0 = ad.burn_from_targetleaves(roots, burn)
    if c0 is not None:
        leafmap[c0] = 7
    assert [_eval_by_id(dual[(F0.id, o)], leafmap, leaf) for o in olist] == want0
    assert [_eval_by_id(dual[(F0_r1.id, o)], leafmap, leaf) for o in olist] == want1


def test_node_derivative_forwardAD_and_backAD_known_answers():
    """reference test/computational_graph.jl:887-928 (testsets "node_derivative" and "Eval"), every leaf = 1."""
    import fdgraph_b200 as fd
    from oracle.frontend import ad

    def ev(x):  # eval!(::Number) = the number; eval!(graph) with all leaves one; eval!(nothing) does not exist: None
        if x is None or isinstance(x, (int, float)):
            return x
        ids = {n.id for n in fd.graph.post_order_unique([x]) if not n.subgraphs}
        return _eval_by_id(x, {i: 1 for i in ids}, [1.0])

    fd.uidreset()
    g1, g2 = fd.Graph([]), fd.Graph([])
    g3 = fd.Graph([], factor=2.0)
    G3 = g1
    G4 = 4 * g1 * g1
    G5 = 4 * (2 * G3 + 3 * G4)
    G6 = (2 * g1 + 3 * g2) * (4 * g1 + g3) * g1
    G7 = (3 * g1 + 4 * g2 + 5 * g3) * 3 * g1
    F1 = g1 * g1
    F2 = (3 * g1) * (4 * g1)
    F3 = (2 * g1 * g2) * (3 * g1)
    F4 = (2 * g1 + 3 * g2) + g1
    assert ev(ad.node_derivative(F1, g1)) == 2
    assert ev(ad.node_derivative(F2, g1)) == 24
    assert ev(ad.node_derivative(F1, g2)) is None
    assert ev(ad.node_derivative(F3, g1)) == 6  # local: only the children of the root count
    assert ev(ad.node_derivative(F4, g1)) == 1
    assert ev(ad.forwardAD(G3, g1.id)) == 1
    assert ev(ad.forwardAD(G4, g1.id)) == 8
    assert ev(ad.forwardAD(G5, g1.id)) == 104
    assert ev(ad.forwardAD(G6, g1.id)) == 62
    assert ev(ad.forwardAD(G6, g2.id)) == 18
    assert ev(ad.forwardAD(ad.forwardAD(G6, g1.id), g2.id)) == 30
    assert ev(ad.forwardAD(G6, g3.id)) == 0
    for G in (G3, G4, G5, G6, G7):
        back = ad.backAD(G)
        assert back
        for (_, leaf_id), value_back in back.items():
            assert ev(value_back) == ev(ad.forwardAD(G, leaf_id))
