"""Restated graph-level automatic differentiation (forwardAD_root!, build_derivative_graph, burn_from_targetleaves!).
TEST / WORKLOAD INFRASTRUCTURE -- a producer of evaluator inputs (derivative graphs are compiled and evaluated like any
other graph), not part of the hot path.

Reference: src/computational_graph/operation.jl:331-338 (insert_dualDict! is not needed here), :354-450
           (forwardAD_root!), :453-461 (find_last_neighbor), :478-543 (build_derivative_graph),
           src/computational_graph/optimize.jl:405-456 (burn_from_targetleaves!),
           src/computational_graph/abstractgraph.jl:14 (decrement_power).

Pinned on the reference's own known answers (test/computational_graph.jl:930-1071) in tests/test_oracle_kat.py.

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
