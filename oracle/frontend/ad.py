"""Restated graph-level automatic differentiation (forwardAD_root!, build_derivative_graph, burn_from_targetleaves!; at the
end of the file the older forwardAD / node_derivative / backAD).
TEST / WORKLOAD INFRASTRUCTURE -- a producer of evaluator inputs (derivative graphs are compiled and evaluated like any
other graph), not part of the hot path.

Reference: src/computational_graph/operation.jl:331-338 (insert_dualDict! is not needed here), :354-450
           (forwardAD_root!), :453-461 (find_last_neighbor), :478-543 (build_derivative_graph),
           src/computational_graph/optimize.jl:405-456 (burn_from_targetleaves!),
           src/computational_graph/abstractgraph.jl:14 (decrement_power).

Pinned on the reference's own known answers (test/computational_graph.jl:887-1071) in tests/test_oracle_kat.py.

`dual` maps (node id, key) to the derivative graph of that node; key is a tuple of N booleans in forwardAD_root! (one
true entry: the variable differentiated) and a tuple of N orders in build_derivative_graph.  A node whose derivative is
needed before the traversal reaches it gets a placeholder named "UNDEFINED" that is filled in place later, exactly as in
the reference, so that graphs built earlier see the final content (object identity is what carries the sharing).
"""
from __future__ import annotations

import itertools
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

from fdgraph_b200.graph import Graph, Power, Prod, Sum, Unitary, constant_graph, linear_combination

UNDEFINED = "UNDEFINED"
BURNING = "BURNING"


def _pre_order_tree(graph: Graph):
    """AbstractTrees.PreOrderDFS: the tree expansion, parents before children, children in stored order.  The children
    list is read when the node is expanded (the traversal of a dual graph sees subgraphs filled in meanwhile)."""
    stack = [graph]
    while stack:
        node = stack.pop()
        yield node
        stack.extend(reversed(node.subgraphs))


def _decrement_power(op: Power):
    return Sum() if op.N == 2 else Power(op.N - 1)


def forwardAD_root(graphs: Sequence[Graph], idx: int = 1, dual: Optional[Dict[Tuple[int, Tuple[bool, ...]], Graph]] = None,
                   n_vars: int = 1) -> Dict[Tuple[int, Tuple[bool, ...]], Graph]:
    """operation.jl:354-450.  `idx` is 1-based like the reference's; `n_vars` is the N of the dictionary's key type."""
    if dual is None:
        dual = {}
    assert 1 <= idx <= n_vars, "the differential variable's index must be described by the key of dual."
    key2 = tuple(i == idx for i in range(1, n_vars + 1))
    done = set()  # objects expanded in this call: a second visit finds them defined and only re-walks their children

    def dual_of(sub: Graph) -> Graph:
        key = (sub.id, key2)
        if key not in dual:
            dual[key] = Graph([], name=UNDEFINED)
        return dual[key]

    for diag in graphs:
        stack = [diag]
        while stack:
            node = stack.pop()
            if id(node) in done:
                continue  # (the reference walks the sub-tree again and skips every node of it: all are defined by now)
            done.add(id(node))
            stack.extend(reversed(node.subgraphs))
            key_node = (node.id, key2)
            visited = key_node in dual
            if visited and dual[key_node].name != UNDEFINED:
                continue
            op = node.operator
            if isinstance(op, Sum):
                deriv, factors, new_op = [dual_of(s) for s in node.subgraphs], list(node.subgraph_factors), None
            elif isinstance(op, Prod):
                deriv = []
                for i, sub in enumerate(node.subgraphs):
                    d = dual_of(sub)
                    terms = [d if j == i else s for j, s in enumerate(node.subgraphs)]
                    deriv.append(Graph(terms, operator=Prod(), subgraph_factors=node.subgraph_factors))
                factors, new_op = [1.0] * len(deriv), None
            elif isinstance(op, Power):
                inner = Graph(node.subgraphs, subgraph_factors=[float(op.N)], operator=_decrement_power(op))
                deriv, factors, new_op = [dual_of(node.eldest()), inner], [1.0, node.subgraph_factors[0]], Prod()
            else:
                continue  # Unitary and user operators: no branch in the reference either
            if visited:
                d = dual[key_node]
                d.subgraphs, d.subgraph_factors, d.name = deriv, [float(f) for f in factors], node.name
                if new_op is not None:
                    d.operator = new_op  # (the reference writes `dual.operator = Prod`, a typo that would throw; the intent is this)
            else:
                dual[key_node] = Graph(deriv, subgraph_factors=factors, operator=new_op or Sum())
    return dual


def _find_last_neighbor(item: Tuple[int, ...]):
    loc = max((j for j, v in enumerate(item) if v > 0), default=None)
    if loc is None:
        return None
    return tuple(v - 1 if j == loc else v for j, v in enumerate(item))


def _leaves(graph: Graph):
    seen = set()
    for node in _pre_order_tree(graph):
        if not node.subgraphs and id(node) not in seen:
            seen.add(id(node))
            yield node


def build_derivative_graph(graphs, orders: Sequence[int], nodes_id: Optional[Iterable[int]] = None
                           ) -> Dict[Tuple[int, Tuple[int, ...]], Graph]:
    """operation.jl:478-543: derivative graphs of the roots (and of the leaves, or of `nodes_id`) up to `orders`."""
    if isinstance(graphs, Graph):
        graphs = [graphs]
    graphs = list(graphs)
    orders = tuple(int(o) for o in orders)
    n = len(orders)
    roots_id = list(dict.fromkeys(g.id for g in graphs))
    if nodes_id is None:
        nodes_id = list(dict.fromkeys(leaf.id for g in graphs for leaf in _leaves(g)))
    cumsum = list(itertools.accumulate(orders))

    def var_of(x: int) -> int:  # findfirst(val -> x <= val, cumsum_orders), 0-based
        return next(j for j, v in enumerate(cumsum) if x <= v)

    def onehot(j: int, kind=int):
        return tuple(kind(i == j) for i in range(n))

    def differs(a, b):
        return tuple(x != y for x, y in zip(a, b))

    idx0 = var_of(1)
    first_order = onehot(idx0)
    one: Dict[Tuple[int, Tuple[bool, ...]], Graph] = {}
    forwardAD_root(graphs, idx0 + 1, one, n)
    dual_graphs = [one[(g.id, onehot(idx0, bool))] for g in graphs]
    for x in range(2, sum(orders) + 1):
        idx = var_of(x)
        forwardAD_root(dual_graphs, idx + 1, one, n)
        dual_graphs = [one[(g.id, onehot(idx, bool))] for g in dual_graphs]
    dual: Dict[Tuple[int, Tuple[int, ...]], Graph] = {}
    zero = (0,) * n
    # Iterators.product runs its FIRST iterator fastest
    for node_id in nodes_id:
        for rev in itertools.product(*[range(o + 1) for o in reversed(orders)]):
            order = tuple(reversed(rev))
            if order == zero:
                continue
            prev = _find_last_neighbor(order)
            if prev == zero:
                dual[(node_id, order)] = one[(node_id, differs(prev, order))]
            else:
                dual[(node_id, order)] = one[(dual[(node_id, prev)].id, differs(prev, order))]
    cum0 = [0] + cumsum
    for root_id in roots_id:
        dual[(root_id, first_order)] = one[(root_id, onehot(idx0, bool))]
        prev = first_order
        for x in range(2, sum(orders) + 1):
            idx = var_of(x)
            order = tuple(x - cum0[idx] if j == idx else (orders[j] if j < idx else 0) for j in range(n))
            dual[(root_id, order)] = one[(dual[(root_id, prev)].id, differs(prev, order))]
            prev = order
    return dual


def burn_from_targetleaves(graphs: Sequence[Graph], targetleaves_id: Sequence[int]) -> Optional[int]:
    """optimize.jl:405-456: removes, in place, everything connected to the target leaves through Prod / Power nodes
    (those leaves are known to be zero).  Returns the id of a constant graph when a whole graph burnt, else None; such a
    graph becomes a Unitary leaf of weight zero carrying that id."""
    from fdgraph_b200.graph import post_order_unique

    graphs = list(graphs)
    targets = set(targetleaves_id)
    graphs_sum = linear_combination(graphs, [1.0] * len(graphs))
    for leaf in _leaves(graphs_sum):
        if leaf.id in targets:
            leaf.name = BURNING
    for node in post_order_unique([graphs_sum]):
        if any(s.name == BURNING for s in node.subgraphs):
            if isinstance(node.operator, (Prod, Power)):
                node.subgraphs, node.subgraph_factors, node.name = [], [], BURNING
            else:
                keep = [(s, f) for s, f in zip(node.subgraphs, node.subgraph_factors) if s.name != BURNING]
                node.subgraphs, node.subgraph_factors = [s for s, _ in keep], [f for _, f in keep]
                if not keep:
                    node.name = BURNING
    c1 = constant_graph(1.0)
    has_c0 = False
    for g in graphs:
        if g.name == BURNING:
            has_c0 = True
            g.id, g.operator, g.weight = c1.id, Unitary(), 0.0
    return c1.id if has_c0 else None


# ---------------------------------------------------------------------------------------------------------
# the older derivative builders: operation.jl:12-44 (linear_combination_number_with_graph), :53-122 (forwardAD),
# :130-147 (all_parent), :161-224 (node_derivative), :226-281 (recursive_backAD!, backAD).  A derivative is a Graph or
# a plain number (`Union{F, Graph}` in the reference); None stands for the reference's `nothing`.
# Known answers: test/computational_graph.jl:901-928.
# ---------------------------------------------------------------------------------------------------------


def _is_number(x) -> bool:
    return isinstance(x, (int, float))


def _times(a, b):
    """`*` of the reference between numbers and graphs (graph.jl:136-163, :403-415)."""
    if _is_number(a) and _is_number(b):
        return a * b
    if _is_number(b):
        return a * b          # Graph.__mul__(number)
    return a * b if not _is_number(a) else b * a  # graph * graph, or number * graph -> Graph.__mul__(number)


def linear_combination_number_with_graph(g: list, coeff: Optional[list] = None):
    if coeff is None:
        coeff = [1.0] * len(g)
    assert len(g) == len(coeff)
    subgraphs, subcoeff, subnumber = [], [], None
    for child, c in zip(g, coeff):
        if _is_number(child):
            subnumber = child * c if subnumber is None else subnumber + child * c
        else:
            assert isinstance(child, Graph), "The type of subgraphs in derivative is incorrect!"
            subgraphs.append(child)
            subcoeff.append(c)
    if subgraphs:
        if subnumber is not None:
            subgraphs.append(constant_graph(float(subnumber)))
            subcoeff.append(1.0)
        return linear_combination(subgraphs, subcoeff)
    return subnumber


def forwardAD(diag: Graph, ID: int):
    """d diag / d (the leaf with id ID), forward propagation (operation.jl:53-122)."""
    from fdgraph_b200.graph import post_order_unique

    dual: Dict[int, object] = {}
    rootid = -1
    for d in post_order_unique([diag]):
        rootid = d.id
        if d.id in dual:
            continue
        if not d.subgraphs:
            if d.id == ID:
                dual[d.id] = 1.0
            continue
        op = d.operator
        if isinstance(op, Sum):
            children, coeff = [], []
            for sub, f in zip(d.subgraphs, d.subgraph_factors):
                if sub.id in dual:
                    children.append(dual[sub.id])
                    coeff.append(f)
            dum = linear_combination_number_with_graph(children, coeff)
            if dum is not None:
                dual[d.id] = dum
        elif isinstance(op, Prod):
            factor, children = 1.0, []
            for si, sub in enumerate(d.subgraphs):
                if sub.id not in dual:
                    continue
                factor *= d.subgraph_factors[si]
                child = dual[sub.id]
                for sj, other in enumerate(d.subgraphs):
                    if si != sj:
                        child = _times(child, other)
                children.append(child)
            dum = linear_combination_number_with_graph(children)
            if dum is not None:
                dual[d.id] = _times(factor, dum)
        elif isinstance(op, Power):
            if d.eldest().id not in dual:
                continue
            children = [Graph(d.subgraphs, subgraph_factors=[float(op.N)], operator=_decrement_power(op))]
            child = dual[d.eldest().id]
            children.append(constant_graph(float(child)) if _is_number(child) else child)
            dual[d.id] = Graph(children, subgraph_factors=[d.subgraph_factors[0], 1.0], operator=Prod())
        else:
            raise NotImplementedError("not implemented!")
    if not dual:
        return 0.0
    return dual.get(rootid)  # (the reference indexes dual[rootid]: a KeyError there when the root does not depend on ID)


def node_derivative(g1: Graph, g2: Graph):
    """The LOCAL derivative d g1 / d g2 (only g1's own subgraphs are looked at), operation.jl:161-224."""
    import copy

    if not g1.subgraphs:
        return None
    op = g1.operator
    if isinstance(op, Sum):
        hits = [f for s, f in zip(g1.subgraphs, g1.subgraph_factors) if s.id == g2.id]
        return float(sum(hits)) if hits else None
    if isinstance(op, Prod):
        count, subgraphs, factors, factor = 0, [], [], None
        for s, f in zip(g1.subgraphs, g1.subgraph_factors):
            if s.id == g2.id:
                count += 1
                if count == 1:
                    factor = f          # the first g2 is the one removed
                    continue
            subgraphs.append(s)
            factors.append(f)
        if count == 0:
            return None
        if not subgraphs:
            return factor
        factors[0] *= count * factor
        g = copy.copy(g1)               # deepcopy in the reference; the subgraphs are replaced right away
        g.subgraphs, g.subgraph_factors = subgraphs, factors
        return g
    if isinstance(op, Power):
        if g1.eldest().id == g2.id:
            return Graph(g1.subgraphs, subgraph_factors=[f * op.N for f in g1.subgraph_factors], operator=_decrement_power(op))
        return None
    return None


def all_parent(diag: Graph) -> Dict[int, List[Graph]]:
    from fdgraph_b200.graph import post_order_unique

    nodes = post_order_unique([diag])
    result: Dict[int, List[Graph]] = {}
    for d in nodes:
        if d.id in result:
            continue
        parents, seen = [], set()
        for g in nodes:
            if g.id not in seen and any(s.id == d.id for s in g.subgraphs):
                parents.append(g)
                seen.add(g.id)
        result[d.id] = parents
    return result


def backAD(diag: Graph) -> Dict[Tuple[int, int], Graph]:
    """(root id, leaf id) -> d root / d leaf for every non-constant leaf, backward propagation (operation.jl:226-281)."""
    dual: Dict[int, object] = {}
    result: Dict[Tuple[int, int], Graph] = {}
    parents = all_parent(diag)

    def back(node: Graph):
        if node.id not in dual:
            if not parents[node.id]:
                dual[node.id] = 1.0
            else:
                terms = []
                for parent in parents[node.id]:
                    parent_ad = back(parent)
                    d_node = node_derivative(parent, node)
                    if d_node is not None and parent_ad is not None:
                        terms.append(_times(d_node, parent_ad))
                dual[node.id] = linear_combination_number_with_graph(terms)
        if not node.subgraphs:
            v = dual[node.id]
            result[(diag.id, node.id)] = constant_graph(float(v)) if _is_number(v) else v
        return dual[node.id]

    for leaf in _leaves(diag):
        if isinstance(leaf.operator, Unitary) or leaf.id in dual:
            continue
        back(leaf)
    return result
