"""The restated front ends (oracle/frontend/: the producers of every workload graph) pinned on the reference's own golden
numbers.  They are test / workload infrastructure -- nothing on the evaluation path imports them -- but the headline
workload is their output, so their known answers belong in the suite:

  * closed-form diagram counts, all leaves = 1 (test/front_end.jl:600-656 with
    src/frontend/parquet/benchmark/diagram_count.jl:53-66): Sigma in the G^2 v expansion, loops 1-4 = 1, 1+s, 4+5s+s^2,
    27+40s+14s^2+s^3;
  * leaf identity of interactions (equal-time rule), parameter equality, ordered partitions, first loop / time indices
    (test/front_end.jl:70-102, :126-183);
  * which Green's functions / self-energies a filter allows (test/front_end.jl:186-219, :660-700);
  * the RPA chain of the 4-point vertex, +-2^3 2^4 (test/front_end.jl:398-443);
  * value invariance under optimize! for the 4-point vertex, loops 1-3, every channel (test/front_end.jl:446-598) and
    for a hand-built graph (test/computational_graph.jl:471-491);
  * the hand-built diagram of test/front_end.jl:221-310 and its Taylor coefficients, closed forms (-2+spin) f {1, 2, 2}.

  * the 3-point vertex and the polarisation (vertex3.jl, polarization.jl; no BASELINE configuration needs them, they
    are here for their counts, test/front_end.jl:701-825 with diagram_count.jl:23-36, :73-123): Gamma3 in the G^2 v
    expansion, loops 1-3 = 1, 4+3s, 27+31s+5s^2; polarisation in the G^2 v expansion, loops 1-4 = s x Gamma3(loops-1);
    in the g^2 v expansion without Fock sub-diagrams, loops 1-4 = 2, 2, 32, 326 (up-up + up-down) and 2, 2, 28, 274
    (up-up alone)."""
import math

import numpy as np
import pytest

import fdgraph_b200 as fd
from fdgraph_b200.graph import Graph, Power, Prod, Sum
from oracle import oracle as O
from oracle.frontend import optimize as opt
from oracle.frontend import parquet as pq
from oracle.frontend import taylor
from oracle.frontend.ids import (Alli, BareGreenId, BareInteractionId, ChargeCharge, Dynamic, GenericId, Girreducible, Instant,
                                 NoFock, NoHartree, PHEr, PHr, PPr, Proper, UpDown, UpUp)


def _eval(graphs, leaf_value=lambda leaf: 1.0):
    """Graphs.eval! restated on Graph objects (eval.jl:15-39): post-order, Prod as prod(g_i * f_i), each object once."""
    val = {}
    for node in fd.graph.post_order_unique(list(graphs)):
        if not node.subgraphs:
            val[id(node)] = leaf_value(node)
            continue
        terms = [val[id(s)] * f for s, f in zip(node.subgraphs, node.subgraph_factors)]
        if isinstance(node.operator, Sum):
            val[id(node)] = sum(terms)
        elif isinstance(node.operator, Prod):
            val[id(node)] = math.prod(terms)
        else:
            assert isinstance(node.operator, Power)
            val[id(node)] = terms[0] ** node.operator.N
    return [val[id(g)] for g in graphs]


def _leaf_by_content(leaf):
    """A reproducible value that depends on WHAT the leaf is (its properties), so that optimize!'s merging of equal leaves
    keeps the assignment: the physical-propagator evaluation of the reference tests, without the physics."""
    h = hash(repr(opt._prop_key(leaf.properties))) & 0xFFFFFFFF
    return 0.5 + (h % 100003) / 100003.0


def count_sigma_G2v(loops, spin):  # benchmark/diagram_count.jl:53-66
    return {1: 1, 2: 1 + spin, 3: 4 + 5 * spin + spin ** 2, 4: 27 + 40 * spin + 14 * spin ** 2 + spin ** 3}[loops]


@pytest.mark.parametrize("loops", [1, 2, 3, 4])
def test_sigma_diagram_counts(loops):
    # test/front_end.jl:600-656
    fd.uidreset()
    pq._ver4I.clear()
    spin = 2
    para = pq.DiagPara(type=pq.SigmaDiag, hasTau=True, innerLoopNum=loops, totalLoopNum=loops + 1, totalTauNum=loops, isFermi=False,
                       spin=spin, firstLoopIdx=2, firstTauIdx=1, filter=(NoHartree, Girreducible),
                       interaction=(pq.Interaction(ChargeCharge, Instant),), extra=pq.ParquetBlocks(phi=[PHEr, PPr], ppi=[PHr, PHEr]))
    ext_k = [0.0] * para.totalLoopNum
    ext_k[0] = 1.0
    rows = pq.sigma(para, ext_k, False)
    merged = pq.mergeby_vec([r["diagram"] for r in rows])
    assert len(merged) == 1
    (w,) = _eval(merged)
    assert w * (-1) ** loops == pytest.approx(count_sigma_G2v(loops, spin), rel=0, abs=1e-9)
    assert [1, 3, 18, 171][loops - 1] == count_sigma_G2v(loops, 2)
    # the evaluator's own interpreter agrees (flattened graph, all leaves one)
    raw, _ = fd.flatten(merged)
    orc = O.Oracle(raw)
    assert orc.eval(np.ones((orc.n_leaves, 1)), "eval")[0, 0] == w


def count_ver3_G2v(loops, spin):  # benchmark/diagram_count.jl:23-36
    return {0: 1, 1: 1, 2: 4 + 3 * spin, 3: 27 + 31 * spin + 5 * spin ** 2}[loops]


def _gamma3(loops, flt):
    # getGamma3, test/front_end.jl:702-729
    fd.uidreset()
    pq._ver4I.clear()
    para = pq.DiagPara(type=pq.Ver3Diag, innerLoopNum=loops, isFermi=False, hasTau=True, filter=flt,
                       interaction=(pq.Interaction(ChargeCharge, Instant),))
    q, k_in = [0.0] * para.totalLoopNum, [0.0] * para.totalLoopNum
    q[0], k_in[1] = 1.0, 1.0
    return para, pq.vertex3(para, [q, k_in])


@pytest.mark.parametrize("loops", [1, 2, 3])
def test_vertex3_diagram_counts(loops):
    # test/front_end.jl:732-750
    para, rows = _gamma3(loops, (NoHartree, Girreducible, Proper))
    assert all(r["extT"][0] == para.firstTauIdx and r["extT"][1] == para.firstTauIdx + 1 for r in rows)
    merged = pq.mergeby_df(rows, [])
    assert len(merged) == 1
    (w,) = _eval([merged[0]["diagram"]])
    assert w * (-1) ** loops == pytest.approx(count_ver3_G2v(loops, para.spin), rel=0, abs=1e-9)
    assert [1, 10, 109][loops - 1] == count_ver3_G2v(loops, 2)
    raw, _ = fd.flatten([merged[0]["diagram"]])
    orc = O.Oracle(raw)
    assert orc.eval(np.ones((orc.n_leaves, 1)), "eval")[0, 0] == w


def _polar(loops, flt):
    # getPolar, test/front_end.jl:759-781
    fd.uidreset()
    pq._ver4I.clear()
    para = pq.DiagPara(type=pq.PolarDiag, innerLoopNum=loops, isFermi=False, hasTau=True, filter=flt,
                       interaction=(pq.Interaction(ChargeCharge, Instant),))
    q = [0.0] * para.totalLoopNum
    q[0] = 1.0
    return para, pq.polarization(para, q)


@pytest.mark.parametrize("loops", [1, 2, 3, 4])
def test_polarization_diagram_counts(loops):
    # test/front_end.jl:783-825
    spin = 2
    sign = (-1) ** (loops - 1)
    para, rows = _polar(loops, (NoHartree, Girreducible))  # G^2 v expansion
    assert all(r["extT"] == (para.firstTauIdx, para.firstTauIdx + 1) for r in rows)
    (w,) = _eval([pq.mergeby_df(rows, [])[0]["diagram"]])
    assert w * spin * sign == pytest.approx(spin * count_ver3_G2v(loops - 1, spin), rel=0, abs=1e-9)
    para, rows = _polar(loops, (NoHartree, NoFock))  # g^2 v expansion, up-up + up-down and up-up alone
    total = pq.mergeby_df(rows, [])[0]["diagram"]
    w, upup = _eval([total, rows[0]["diagram"]])
    assert rows[0]["response"] == UpUp
    assert w * spin * sign == pytest.approx([2, 2, 32, 326][loops - 1], rel=0, abs=1e-9)
    assert upup * spin * sign == pytest.approx([2, 2, 28, 274][loops - 1], rel=0, abs=1e-9)
    raw, _ = fd.flatten([total])
    orc = O.Oracle(raw)
    assert orc.eval(np.ones((orc.n_leaves, 1)), "eval")[0, 0] == w


def test_polarization_with_an_explicit_proper_filter():
    # test/front_end.jl:784: the builder accepts a filter that already holds Proper
    from oracle.frontend.ids import Proper
    para, rows = _polar(1, (Proper, NoHartree, NoFock))
    assert len(rows) == 1 and not rows[0]["diagram"].subgraphs[0].subgraphs  # Pi0 = G G, two bare propagators


@pytest.mark.parametrize("kind", [Instant, Dynamic])
def test_interaction_ids_equal_time_rule(kind):
    # test/front_end.jl:70-102 (diagram_id.jl:49-69): leaf identity of interactions -- equal-time labels are interchangeable
    a = BareInteractionId(UpUp, kind, k=[1.0, 1.0], t=(1, 1))
    b = BareInteractionId(UpUp, kind, k=[1.0, 1.0], t=(1, 1))
    c = BareInteractionId(UpUp, kind, k=[2.0, 2.0], t=(2, 2))
    d = BareInteractionId(UpUp, kind, k=[1.0, 1.0], t=(2, 2))
    e = BareInteractionId(UpUp, kind, k=[1.0, 1.0], t=(1, 2))
    f = BareInteractionId(UpUp, kind, k=[1.0, 1.0], t=(1, 2))
    assert a == b and a != c and a == d and a != e and e == f
    assert hash(a) == hash(d) and opt._prop_key(a) == opt._prop_key(d)  # what remove_duplicated_leaves compares


def test_parameters_partitions_and_first_indices():
    # test/front_end.jl:126-147 "Parameter"
    p, q, a = (pq.DiagPara(type=pq.Ver4Diag, innerLoopNum=n) for n in (1, 2, 2))
    assert p.key() != q.key() and q.key() == a.key()
    assert a.key() != a.reconstruct(transferLoop=[0.0, 0.0, 0.0]).key()
    assert a.key() != a.reconstruct(interaction=[]).key()
    assert a.reconstruct(type=pq.SigmaDiag).key() != pq.DiagPara(type=pq.SigmaDiag, innerLoopNum=2).key()
    # :149-157 "Partition"
    assert set(pq.ordered_partition(5, 2)) == {(4, 1), (1, 4), (2, 3), (3, 2)}
    assert set(pq.ordered_partition(3, 2, 0)) == {(3, 0), (0, 3), (1, 2), (2, 1)}
    # :159-183 "FindFirstIdx"
    for partition, first, expected in (([1, 1, 2, 1], 1, [1, 2, 3, 5]), ([1, 1, 2, 1], 0, [0, 1, 2, 4]), ([1, 0, 2, 0], 1, [1, 2, 2, 4]),
                                       ([1], 1, [1])):
        idx, total = pq.find_first_loop_idx(partition, first)
        assert idx == expected and total == sum(partition) + first - 1
    kinds = [pq.Ver4Diag, pq.GreenDiag, pq.Ver4Diag, pq.GreenDiag]
    for partition, first, expected in (([1, 1, 2, 1], 1, [1, 3, 4, 7]), ([1, 1, 2, 1], 0, [0, 2, 3, 6]), ([1, 0, 2, 0], 1, [1, 3, 3, 6])):
        assert pq.find_first_tau_idx(partition, kinds, first, 1)[0] == expected


def test_which_green_functions_and_self_energies_a_filter_allows():
    # test/front_end.jl:186-219
    assert pq.is_valid_g((Girreducible,), 0) is True and pq.is_valid_g((Girreducible,), 1) is False and pq.is_valid_g((Girreducible,), 2) is False
    assert pq.is_valid_g((NoFock,), 0) is True and pq.is_valid_g((NoFock,), 1) is True and pq.is_valid_g((NoFock, NoHartree), 1) is False
    assert pq.is_valid_g((NoFock, NoHartree), 2) is True
    for loops, sub, want in ((0, True, False), (1, True, False), (2, True, False), (0, False, False), (1, False, True), (2, False, True)):
        assert pq.is_valid_sigma((Girreducible,), loops, sub) is want
    assert pq.is_valid_sigma((NoFock,), 0, True) is False
    assert pq.is_valid_sigma((NoFock,), 1, True) is True        # a one-loop self-energy is Hartree or Fock ...
    assert pq.is_valid_sigma((NoFock, NoHartree), 1, True) is False  # ... and gone only if both are filtered
    assert pq.is_valid_sigma((NoFock, NoHartree), 2, True) is True
    assert pq.is_valid_sigma((NoFock,), 0, False) is False
    assert pq.is_valid_sigma((NoFock,), 1, False) is True
    assert pq.is_valid_sigma((NoFock, NoHartree), 1, False) is True
    assert pq.is_valid_sigma((NoFock,), 2, False) is True

    # test/front_end.jl:660-700
    def build_g(loops, flt):
        para = pq.DiagPara(type=pq.GreenDiag, hasTau=True, innerLoopNum=loops, isFermi=True, spin=2, filter=flt,
                           interaction=(pq.Interaction(ChargeCharge, Instant),))
        ext_k = [0.0] * para.totalLoopNum
        ext_k[0] = 1.0
        return pq.green(para, ext_k, (1, 2)) if pq.is_valid_g(para.filter, para.innerLoopNum) else None

    fd.uidreset()
    pq._ver4I.clear()
    assert isinstance(build_g(0, (NoHartree, Girreducible)), Graph)
    assert build_g(1, (NoHartree, Girreducible)) is None and build_g(2, (NoHartree, Girreducible)) is None
    assert isinstance(build_g(0, (NoHartree, NoFock)), Graph)
    assert build_g(1, (NoHartree, NoFock)) is None
    assert isinstance(build_g(2, (NoHartree, NoFock)), Graph)  # a higher-order sub-diagram is allowed


def test_rpa_chain_of_the_four_point_vertex():
    # test/front_end.jl:398-443: each bubble contributes 2, each dynamic interaction 2, two spin configurations
    loops = 3
    fd.uidreset()
    pq._ver4I.clear()
    para = pq.DiagPara(type=pq.Ver4Diag, hasTau=True, innerLoopNum=loops, interaction=(pq.Interaction(ChargeCharge, [Instant, Dynamic]),))
    k1, k2, k3 = (pq.get_k(para.totalLoopNum, i) for i in (1, 2, 3))
    ext_k = [k1, k2, k3, [a + c - b for a, b, c in zip(k1, k2, k3)]]
    weight = (2 ** loops) * (2 ** (loops + 1))

    def chain(chan):
        rows = []
        pq.rpa_chain(rows, para, ext_k, chan, 0, "RPA", -1.0)
        merged = pq.mergeby_df(rows, ["response"])
        by_resp = {r["response"]: _eval([r["diagram"]])[0] for r in merged}
        return by_resp.get(UpUp, 0.0), by_resp.get(UpDown, 0.0)

    upup, updown = chain(PHEr)
    assert upup == pytest.approx(-weight) and updown == pytest.approx(0.0)  # exchange: extra sign, no up-down
    upup, updown = chain(PHr)
    assert upup == pytest.approx(weight) and updown == pytest.approx(weight)


@pytest.mark.parametrize("loops", [1, 2, 3])
@pytest.mark.parametrize("chan", [(PHr,), (PHEr,), (PPr,), (PHr, PHEr, PPr)])
def test_vertex4_value_is_invariant_under_optimize(loops, chan):
    # test/front_end.jl:446-598 (`@assert w1 ≈ w1opt`), with content-keyed leaf values in place of the physical propagators
    fd.uidreset()
    pq._ver4I.clear()
    k0 = [0.0] * (loops + 2)
    kin_l, kin_r = list(k0), list(k0)
    kin_l[0] = 1.0
    kin_r[1] = 1.0
    leg_k = [kin_l, list(kin_l), kin_r, list(kin_r)]
    blocks = pq.ParquetBlocks(phi=[PHEr, PPr], ppi=[PHr, PHEr])
    para = pq.DiagPara(type=pq.Ver4Diag, isFermi=True, hasTau=True, innerLoopNum=loops, totalLoopNum=len(kin_l), totalTauNum=loops + 1,
                       spin=2, firstLoopIdx=3, firstTauIdx=1, filter=(NoHartree, Girreducible), transferLoop=tuple(0.0 for _ in kin_l),
                       interaction=(pq.Interaction(ChargeCharge, Instant),), extra=blocks)
    rows = pq.vertex4(para, leg_k, channels=chan, blocks=blocks)
    merged = [r["diagram"] for r in pq.mergeby_df(rows, ["response"])]
    assert len(merged) == 2
    w1 = _eval(merged, _leaf_by_content)
    ones1 = _eval(merged)
    optimised = list(opt.optimize(merged))
    w2 = _eval(optimised, _leaf_by_content)
    assert w2 == pytest.approx(w1, rel=1e-12)
    assert _eval(optimised) == pytest.approx(ones1, rel=0, abs=1e-9)
    n_before = len(fd.graph.post_order_unique(merged))
    assert len(fd.graph.post_order_unique(optimised)) <= n_before


def test_optimize_on_a_hand_built_graph():
    # test/computational_graph.jl:471-491
    fd.uidreset()
    g1 = Graph([])
    g2 = 2 * g1
    g3 = Graph([g2], subgraph_factors=[3], operator=Prod())
    g4 = Graph([g3], subgraph_factors=[5], operator=Prod())
    g5 = Graph([], factor=3.0)
    g6 = Graph([g5, g1], subgraph_factors=[1.0, 2.0], operator=Sum())
    g = Graph([g4, g6], operator=Sum())
    before = _eval([g])[0]
    (h,) = opt.optimize([g])
    assert _eval([h])[0] == before
    assert len([n for n in fd.graph.post_order_unique([h]) if not n.subgraphs]) <= 2


@pytest.mark.parametrize("order", [2, 3])
def test_optimize_level_one_merges_equivalent_nodes_and_keeps_the_values(order):
    """optimize!(level = 1) = remove_duplicated_nodes! (optimize.jl:16-36, :345-390) on the Parquet 4-point vertex: the
    values do not move, fewer nodes are left than at level 0's start, and the product's own merging of common
    sub-expressions (fdg_options.cse, arithmetic only: it ignores the diagram ids the reference's isequiv compares, and
    keeps operand order, which isequiv does not) lands in the same region -- order 3: 4659 statements, 3072 after the
    reference's pass over the level-0 graph, 2843 after the product's."""
    def build():
        fd.uidreset()
        pq._ver4I.clear()
        return [r["diagram"] for r in pq.vertex4(pq.DiagPara(type=pq.Ver4Diag, innerLoopNum=order))]

    g = build()
    want, ones = _eval(g, _leaf_by_content), _eval(g)
    n_start = len(fd.graph.post_order_unique(g))
    opt.optimize(g, level=1)
    assert _eval(g, _leaf_by_content) == pytest.approx(want, rel=1e-12) and _eval(g) == pytest.approx(ones, rel=0, abs=1e-9)
    assert len(fd.graph.post_order_unique(g)) < n_start / 3
    # the same pass over the graph optimize!(level = 0) leaves, beside the product's merging of that graph
    g0 = build()
    opt.optimize(g0, level=0)
    raw0, _ = fd.flatten(g0)
    n0 = O.Oracle(raw0).n_stmts
    merged_by_product = n0 - fd.compile_raw(raw0, cse=True).stats["cse_removed"]
    want0 = _eval(g0, _leaf_by_content)
    opt.remove_duplicated_nodes(Graph(list(g0)))
    assert _eval(g0, _leaf_by_content) == pytest.approx(want0, rel=1e-12)
    n1 = O.Oracle(fd.flatten(g0)[0]).n_stmts
    assert n1 < n0 and merged_by_product < n0
    assert abs(merged_by_product - n1) < 0.1 * n0
    if order == 3:
        assert (n0, n1, merged_by_product) == (4659, 3072, 2843)


def test_optimize_reaches_the_expected_graph_of_the_reference_test():
    """test/computational_graph.jl:471-491 and :676-695 with their own graph and expected result
    _h = 2 (-28 g1 + 3 g1') + 3 g1': the chains 2 * 3 * 5 are flattened into factors, the two leaves stay apart (the
    reference tells them apart by a user operator, here by their properties), equal sub-graphs are merged."""
    fd.uidreset()
    g1 = Graph([])
    g2 = 2 * g1
    g3 = Graph([g2], subgraph_factors=[3], operator=Prod())
    g4 = Graph([g3], subgraph_factors=[5], operator=Prod())
    g5 = Graph([], factor=3.0, properties=BareGreenId(k=[1.0], t=(1, 2)))
    h0 = Graph([g1, g4, g5], subgraph_factors=[2, -1, 1])
    h1 = Graph([h0], operator=Prod(), subgraph_factors=[2])
    h = Graph([h1, g5])
    g1p = g5.eldest()
    values = {g1.id: 0.7391, g1p.id: -1.1875}
    before = _eval([h], lambda leaf: values[leaf.id])[0]
    assert before == pytest.approx(2 * (-28 * values[g1.id] + 3 * values[g1p.id]) + 3 * values[g1p.id], rel=1e-15)
    (o,) = opt.optimize([h])
    assert isinstance(o.operator, Sum) and o.subgraph_factors == [2.0, 3.0] and len(o.subgraphs) == 2
    inner, leaf = o.subgraphs
    assert isinstance(inner.operator, Sum) and inner.subgraph_factors == [-28.0, 3.0]
    assert [s.id for s in inner.subgraphs] == [g1.id, g1p.id] and leaf is inner.subgraphs[1] and not leaf.subgraphs
    assert _eval([o], lambda leaf: values[leaf.id])[0] == pytest.approx(before, rel=1e-15)


def _getdiagram(spin=2.0, D=3, Nk=4, Nt=2):
    """test/front_end.jl:221-263 (the direct part of a two-bubble diagram)."""
    fd.uidreset()
    para = pq.DiagPara(type=pq.GreenDiag, innerLoopNum=0, totalLoopNum=Nk, hasTau=True, totalTauNum=Nt)
    gK = [[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0]]
    gT = [(1, 2), (2, 1)]
    g = [Graph([], properties=BareGreenId(k=gK[i], t=gT[i]), name="G") for i in range(2)]
    vdK = [[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 1.0, 0.0]]
    vd = [Graph([], properties=BareInteractionId(ChargeCharge, k=vdK[i]), name="Vd") for i in range(2)]
    veK = [[1, 0, -1, -1], [0, 1, 0, -1]]
    ve = [Graph([], properties=BareInteractionId(ChargeCharge, k=veK[i]), name="Ve") for i in range(2)]
    gid = GenericId(para)
    ggn = Graph([g[0], g[1]], properties=gid, operator=Prod())
    vdd = Graph([vd[0], vd[1]], properties=gid, operator=Prod(), factor=spin)
    vde = Graph([vd[0], ve[1]], properties=gid, operator=Prod(), factor=-1.0)
    ved = Graph([ve[0], vd[1]], properties=gid, operator=Prod(), factor=-1.0)
    vsum = Graph([vdd, vde, ved], properties=gid, operator=Sum())
    return Graph([vsum, ggn], properties=gid, operator=Prod(), factor=1 / (2 * math.pi) ** D, name="root")


def test_hand_built_parquet_like_graph_and_its_taylor_coefficients():
    # test/front_end.jl:284-310
    spin, D = 1.0, 3
    root = _getdiagram(spin, D)
    factor = 1 / (2 * math.pi) ** D
    rootval = _eval([root])[0]
    assert rootval == pytest.approx((-2 + spin) * factor)
    (root_o,) = opt.optimize([root])
    assert _eval([root_o])[0] == rootval
    # one derivative with respect to the fermionic ("x") and one to the bosonic ("y") propagators; every derivative = 1,
    # i.e. the coefficient leaf of order o carries 1 / o!
    root = _getdiagram(spin, D)
    dep = {}
    for node in fd.graph.post_order_unique([root]):
        if not node.subgraphs:
            dep[node.id] = [isinstance(node.properties, BareGreenId), isinstance(node.properties, BareInteractionId)]
    (series,), _ = taylor.taylorexpansion([root], dep, [1, 1])

    def coeff_leaf(leaf):
        orders = getattr(leaf, "orders", None) or [0, 0]
        return 1.0 / math.prod(math.factorial(int(o)) for o in orders)

    c = {o: _eval([g], coeff_leaf)[0] for o, g in series.coeffs.items()}
    assert c[(0, 0)] == pytest.approx((-2 + spin) * factor)
    assert c[(0, 1)] == pytest.approx((-2 + spin) * 2 * factor)
    assert c[(1, 0)] == pytest.approx((-2 + spin) * 2 * factor)


def test_committed_workloads_are_what_the_front_ends_build():
    """The headline workload file is the restated Parquet front end's output: rebuild the small orders here (the big ones
    take minutes) and compare with the committed arrays; the all-leaves-one checksums of every workload are in the
    manifest and are re-derived from the files by the oracle."""
    import json
    import os

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "workloads", "MANIFEST.json")) as fh:
        manifest = json.load(fh)
    for order in (2, 3):
        fd.uidreset()
        pq._ver4I.clear()
        graphs = [r["diagram"] for r in pq.vertex4(pq.DiagPara(type=pq.Ver4Diag, innerLoopNum=order))]
        opt.optimize(graphs)
        raw, _ = fd.flatten(graphs)
        ref = fd.RawGraph.load(os.path.join(here, "workloads", f"parquet_ver4_o{order}.npz"))
        assert all(np.array_equal(getattr(raw, k), getattr(ref, k)) for k in raw.__dataclass_fields__)
    for order, name, build in ((3, "parquet_ver3_o3", lambda o: pq.vertex3(pq.DiagPara(type=pq.Ver3Diag, innerLoopNum=o))),
                               (4, "parquet_polar_o4", lambda o: pq.polarization(pq.DiagPara(type=pq.PolarDiag, innerLoopNum=o)))):
        fd.uidreset()
        pq._ver4I.clear()
        graphs = [r["diagram"] for r in build(order)]
        opt.optimize(graphs)
        raw, _ = fd.flatten(graphs)
        ref = fd.RawGraph.load(os.path.join(here, "workloads", name + ".npz"))
        assert all(np.array_equal(getattr(raw, k), getattr(ref, k)) for k in raw.__dataclass_fields__)
    for name in ("parquet_ver4_o4", "gv_ver4_o4", "parquet_sigma_o3", "taylor_sigma_o3", "parquet_ver3_o4", "parquet_polar_o5"):
        raw = fd.RawGraph.load(os.path.join(here, "workloads", name + ".npz"))
        orc = O.Oracle(raw)
        ones = orc.eval(np.ones((orc.n_leaves, 1)))[:, 0]
        assert [float(x) for x in ones] == manifest[name]["all_leaves_one"]
        assert orc.n_leaves == manifest[name]["n_leaves"] and orc.n_roots == manifest[name]["n_roots"]
