"""Host-side mirror of the reference's computational-graph node model.

This is the *input format* of the hot path: the evaluator consumes a DAG of
``Graph`` nodes.  Only what the back end reads is mirrored -- ``id``, ``operator``, ``subgraphs``,
``subgraph_factors`` -- plus the constructor / arithmetic conventions that decide which nodes and
factors exist (they define the rounding of the compiled function):

* node model and constructor, ``factor != 1`` wraps the node in a unary ``Prod``:
  reference ``src/computational_graph/graph.jl:28-74`` (``feynmangraph.jl:72-129`` for FeynmanGraph)
* operators ``Sum / Prod / Unitary / Power{N}``: ``src/computational_graph/abstractgraph.jl:3-12``
* ``constant_graph``: ``graph.jl:118-125``; scalar ``*``: ``graph.jl:136-163``
* ``linear_combination``: ``graph.jl:178-262``; ``multi_product``: ``graph.jl:304-401``; ``^``: ``graph.jl:413``
* global id counter ``uid()``: ``src/computational_graph/common.jl:1,15-22``

Nothing here evaluates anything: evaluation is the CUDA path (``compilers.py``).
"""
from __future__ import annotations

import math
from typing import Any, Iterable, List, Optional, Sequence

# ----------------------------------------------------------------------------------------------
# operators (abstractgraph.jl:3-12)
# ----------------------------------------------------------------------------------------------


class Operator:
    """Base of node operators.  Instances compare equal iff same type (abstractgraph.jl:15-16)."""

    code = -1

    def __eq__(self, other):
        return type(self) is type(other) and getattr(self, "N", None) == getattr(other, "N", None)

    def __hash__(self):
        return hash((type(self).__name__, getattr(self, "N", None)))

    def __repr__(self):
        return type(self).__name__


class Sum(Operator):
    code = 1


class Prod(Operator):
    code = 2


class Unitary(Operator):
    code = 0


class Power(Operator):
    code = 3

    def __init__(self, N: int):
        # abstractgraph.jl:8-9  "Power{N} makes no sense" for N in (0, 1)
        if int(N) in (0, 1):
            raise AssertionError(f"Power{{{N}}} makes no sense.")
        self.N = int(N)

    def __repr__(self):
        return f"Power{{{self.N}}}"


def _as_operator(op) -> Operator:
    if isinstance(op, Operator):
        return op
    if isinstance(op, type) and issubclass(op, Operator):
        return op()
    raise TypeError(f"not an operator: {op!r}")


def unary_istrivial(op: Operator) -> bool:
    """(+g) == g and (*g) == g  (abstractgraph.jl:36-37)."""
    return isinstance(op, (Sum, Prod))


# ----------------------------------------------------------------------------------------------
# uid counter (common.jl:1,15-22)
# ----------------------------------------------------------------------------------------------
_counter = [0]


def uid() -> int:
    _counter[0] += 1
    return _counter[0]


def uidreset() -> None:
    _counter[0] = 0


def _isapprox_one(x) -> bool:
    # Julia `x ≈ one(x)`: rtol = sqrt(eps)
    return abs(x - 1.0) <= math.sqrt(2.220446049250313e-16) * max(abs(x), 1.0)


# ----------------------------------------------------------------------------------------------
# Graph (graph.jl:28-74)
# ----------------------------------------------------------------------------------------------


class Graph:
    """Mirror of ``Graph{F,W}``: F = Python float factors, W decided at evaluation time."""

    __slots__ = ("id", "name", "orders", "subgraphs", "subgraph_factors", "operator", "weight", "properties")

    def __new__(cls, subgraphs: Sequence["Graph"] = (), *, factor=1.0, subgraph_factors=None, name="",
                operator=None, orders=None, weight=0.0, properties=None):
        op = _as_operator(operator) if operator is not None else Sum()
        subgraphs = list(subgraphs)
        if isinstance(op, Power):
            assert len(subgraphs) == 1, "Graph with Power operator must have one and only one subgraph."
        elif isinstance(op, Unitary):
            assert len(subgraphs) == 0, "Graph with Unitary operator must have no subgraphs."
        if subgraph_factors is None:
            subgraph_factors = [1.0] * len(subgraphs)
        g = object.__new__(cls)
        g.id = uid()
        g.name = str(name)
        g.orders = list(orders) if orders is not None else [0] * 16
        g.subgraphs = subgraphs
        g.subgraph_factors = [float(f) for f in subgraph_factors]
        assert len(g.subgraph_factors) == len(g.subgraphs)
        g.operator = op
        g.weight = weight
        g.properties = properties
        if _isapprox_one(float(factor)):
            return g
        # graph.jl:69-73: a non-unit `factor` becomes a unary Prod node above g
        w = object.__new__(cls)
        w.id = uid()
        w.name = g.name
        w.orders = g.orders
        w.subgraphs = [g]
        w.subgraph_factors = [float(factor)]
        w.operator = Prod()
        w.weight = weight * factor
        w.properties = properties
        return w

    def __init__(self, *a, **k):  # construction is done in __new__ (it may return a wrapper node)
        pass

    # -- getters mirrored from graph.jl:77-90 ----------------------------------------------------
    def isleaf(self) -> bool:
        return len(self.subgraphs) == 0

    def onechild(self) -> bool:
        return len(self.subgraphs) == 1

    def eldest(self) -> "Graph":
        assert self.subgraphs, "Graph has no children!"
        return self.subgraphs[0]

    def __repr__(self):
        return f"<{type(self).__name__} id={self.id} {self.operator!r} n={len(self.subgraphs)}>"

    # -- arithmetic (graph.jl:136-163, 264-272, 403-415) -------------------------------------------
    def _scale(self, c) -> "Graph":
        g = type(self)([self], subgraph_factors=[float(c)], operator=Prod(), orders=self.orders)
        if unary_istrivial(self.operator) and self.onechild():
            g.subgraph_factors[0] *= self.subgraph_factors[0]
            g.subgraphs = self.subgraphs  # aliasing, exactly like the reference (graph.jl:140)
        return g

    def __mul__(self, other):
        if isinstance(other, Graph):
            return multi_product(self, other)
        return self._scale(other)

    def __rmul__(self, other):
        return self._scale(other)

    def __add__(self, other):
        return linear_combination(self, other, 1.0, 1.0)

    def __sub__(self, other):
        return linear_combination(self, other, 1.0, -1.0)

    def __pow__(self, exponent: int):
        return type(self)([self], operator=Power(int(exponent)), orders=[o * exponent for o in self.orders])


class FeynmanGraph(Graph):
    """Mirror of ``FeynmanGraph{F,W}`` as far as the back end is concerned: same node fields, same
    factor wrapping (feynmangraph.jl:124-128); diagrammatic properties are carried opaquely."""

    __slots__ = ()


def constant_graph(factor=1.0) -> Graph:
    """graph.jl:118-125 -- a Unitary leaf (weight one); scaled through `*` when factor != 1."""
    g = Graph([], operator=Unitary(), weight=1.0)
    if _isapprox_one(float(factor)):
        return g
    return g * factor


def _pad_orders(a: Graph, b: Graph) -> None:
    la, lb = len(a.orders), len(b.orders)
    if la > lb:
        b.orders = list(b.orders) + [0] * (la - lb)
    else:
        a.orders = list(a.orders) + [0] * (lb - la)


def linear_combination(g1, g2=None, c1=1.0, c2=1.0, *, properties=None):
    """Pair form graph.jl:178-208 and vector form graph.jl:222-262 (vector form when g1 is a list)."""
    if isinstance(g1, (list, tuple)):
        return _linear_combination_vec(list(g1), g2, properties=properties)
    _pad_orders(g1, g2)
    assert g1.orders == g2.orders, "g1 and g2 have different orders."
    subgraphs = [g1, g2]
    factors = [float(c1), float(c2)]
    for i, g in enumerate((g1, g2)):
        if unary_istrivial(g.operator) and g.onechild():
            factors[i] *= g.subgraph_factors[0]
            subgraphs[i] = g.subgraphs[0]
    cls = type(g1)
    if subgraphs[0].id == subgraphs[1].id:
        return cls([subgraphs[0]], subgraph_factors=[factors[0] + factors[1]], operator=Sum(),
                   orders=g1.orders, properties=properties)
    return cls(subgraphs, subgraph_factors=factors, operator=Sum(), orders=g1.orders, properties=properties)


def _linear_combination_vec(graphs: List[Graph], constants=None, *, properties=None):
    if constants is None:
        constants = [1.0] * len(graphs)
    maxlen = max(len(g.orders) for g in graphs)
    for g in graphs:
        g.orders = list(g.orders) + [0] * (maxlen - len(g.orders))
    assert all(g.orders == graphs[0].orders for g in graphs), "Graphs do not all have the same order."
    subgraphs = list(graphs)
    factors = [float(c) for c in constants]
    for i, g in enumerate(graphs):
        if unary_istrivial(g.operator) and g.onechild():
            factors[i] *= g.subgraph_factors[0]
            subgraphs[i] = g.subgraphs[0]
    uniq: List[Graph] = []
    ufac: List[float] = []
    pos = {}
    for g, f in zip(subgraphs, factors):
        i = pos.get(g.id)
        if i is None:
            pos[g.id] = len(uniq)
            uniq.append(g)
            ufac.append(f)
        else:
            ufac[i] += f
    if not uniq:
        return None
    return type(graphs[0])(uniq, subgraph_factors=ufac, operator=Sum(), orders=graphs[0].orders, properties=properties)


def multi_product(g1, g2=None, c1=1.0, c2=1.0, *, properties=None):
    """Pair form graph.jl:304-331 and vector form graph.jl:345-401."""
    if isinstance(g1, (list, tuple)):
        return _multi_product_vec(list(g1), g2, properties=properties)
    subgraphs = [g1, g2]
    factors = [float(c1), float(c2)]
    for i, g in enumerate((g1, g2)):
        if unary_istrivial(g.operator) and g.onechild():
            factors[i] *= g.subgraph_factors[0]
            subgraphs[i] = g.subgraphs[0]
    cls = type(g1)
    if subgraphs[0].id == subgraphs[1].id:
        return cls([subgraphs[0]], subgraph_factors=[factors[0] * factors[1]], operator=Power(2),
                   orders=[2 * o for o in g1.orders], properties=properties)
    _pad_orders(g1, g2)
    return cls(subgraphs, subgraph_factors=factors, operator=Prod(),
               orders=[a + b for a, b in zip(g1.orders, g2.orders)], properties=properties)


def _multi_product_vec(graphs: List[Graph], constants=None, *, properties=None):
    if constants is None:
        constants = [1.0] * len(graphs)
    g1 = graphs[0]
    subgraphs = list(graphs)
    factors = [float(c) for c in constants]
    maxlen = max(len(g.orders) for g in graphs)
    g_orders = [0] * maxlen
    for i, g in enumerate(graphs):
        if unary_istrivial(g.operator) and g.onechild():
            factors[i] *= g.subgraph_factors[0]
            subgraphs[i] = g.subgraphs[0]
        g.orders = list(g.orders) + [0] * (maxlen - len(g.orders))
        g_orders = [a + b for a, b in zip(g_orders, g.orders)]
    uniq: List[Graph] = []
    ufac: List[float] = []
    counts: List[int] = []
    pos = {}
    for g, f in zip(subgraphs, factors):
        i = pos.get(g.id)
        if i is None:
            pos[g.id] = len(uniq)
            uniq.append(g)
            ufac.append(f)
            counts.append(1)
        else:
            ufac[i] *= f
            counts[i] += 1
    if not uniq:
        return None
    cls = type(g1)
    if len(ufac) == 1:
        return cls(uniq, subgraph_factors=ufac, operator=Power(counts[0]), orders=g_orders, properties=properties)
    subs = []
    for g, n in zip(uniq, counts):
        if n == 1:
            subs.append(g)
        else:
            subs.append(cls([g], operator=Power(n), orders=[o * n for o in g1.orders]))
    return cls(subs, subgraph_factors=ufac, operator=Prod(), orders=g_orders, properties=properties)


# ----------------------------------------------------------------------------------------------
# traversal helpers (AbstractTrees contract, tree_properties.jl:20-22): children = subgraphs, stored order
# ----------------------------------------------------------------------------------------------


def post_order_unique(roots: Iterable[Graph]):
    """Post-order DFS over the DAG, each *object* once (identity), children in stored order."""
    seen = set()
    out = []
    for r in roots:
        if id(r) in seen:
            continue
        stack = [(r, 0)]
        while stack:
            node, i = stack.pop()
            if i == 0 and id(node) in seen:
                continue
            if i < len(node.subgraphs):
                stack.append((node, i + 1))
                c = node.subgraphs[i]
                if id(c) not in seen:
                    stack.append((c, 0))
            else:
                if id(node) not in seen:
                    seen.add(id(node))
                    out.append(node)
    return out


def leaves(g: Graph):
    """AbstractTrees.Leaves order (tree expansion, duplicates kept, pruned by object for speed)."""
    return [n for n in post_order_unique([g]) if n.isleaf()]
