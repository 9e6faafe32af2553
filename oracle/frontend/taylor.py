"""Restated Taylor-mode AD front end (`taylorAD`).  TEST / WORKLOAD INFRASTRUCTURE.

Builds the graphs of BASELINE config 5 ("Taylor-mode AD renormalized self-energy"): every node of a diagram graph is
replaced by the truncated multivariate Taylor series of its value in the counter-term variables, each coefficient
being a new Graph.  Only the producer of a workload: the back end sees the result as an ordinary Sum/Prod/Power DAG.

Reference: src/utility.jl:11-13 (apply on TaylorSeries), :48-93 (taylorAD), :106-135 (taylorexpansion!, Graph),
           :243-252 (vector form); src/TaylorSeries/arithmetic.jl:10-16,27-33 (series x number), :44-56 (series +
           series), :170-191 (series x series, truncated at the maximal orders), :282-316 (^, power_by_squaring);
           src/TaylorSeries/constructors.jl:11-22; src/TaylorSeries/parameter.jl (set_variables / get_orders).

Known deviation (of the producer, not of the evaluator): the reference keeps coefficients in a Julia `Dict`, whose
iteration order is an implementation detail of the hash table; this restatement iterates in insertion order.  The set
of terms of every coefficient is the same, the operand order inside a coefficient's Sum can differ from a particular
Julia session (all-leaves-one values and the reference's own known answers, test/taylor.jl:42-56,96-112, are
order-independent and are what tests/test_oracle_kat.py pins).
"""
from __future__ import annotations

import itertools
from typing import Callable, Dict, List, Sequence, Tuple

from fdgraph_b200.graph import Graph, Power, Prod, Sum, post_order_unique

Orders = Tuple[int, ...]


class TaylorSeries:
    """constructors.jl:11-22 -- `coeffs` maps an order tuple to a coefficient (a Graph, or a number in the KATs)."""

    __slots__ = ("coeffs", "max_orders")

    def __init__(self, max_orders: Sequence[int], coeffs: Dict[Orders, object] = None):
        self.max_orders = tuple(int(o) for o in max_orders)
        self.coeffs: Dict[Orders, object] = dict(coeffs) if coeffs else {}

    # arithmetic.jl:10-16, 27-33
    def scale(self, c) -> "TaylorSeries":
        out = TaylorSeries(self.max_orders)
        for order, coeff in self.coeffs.items():
            out.coeffs[order] = c * coeff
        return out

    def __mul__(self, other):
        if isinstance(other, TaylorSeries):
            return self._mul_series(other)
        return self.scale(other)

    def __rmul__(self, other):
        return self.scale(other)

    # arithmetic.jl:44-56 and :68-100 (constant)
    def __add__(self, other):
        out = TaylorSeries(self.max_orders, self.coeffs)
        if isinstance(other, TaylorSeries):
            for order, coeff in other.coeffs.items():
                out.coeffs[order] = out.coeffs[order] + coeff if order in out.coeffs else coeff
            return out
        zero = tuple(0 for _ in self.max_orders)
        out.coeffs[zero] = out.coeffs[zero] + other if zero in out.coeffs else other
        return out

    __radd__ = __add__

    # arithmetic.jl:170-191
    def _mul_series(self, other: "TaylorSeries") -> "TaylorSeries":
        out = TaylorSeries(self.max_orders)
        for o1, c1 in self.coeffs.items():
            for o2, c2 in other.coeffs.items():
                order = tuple(a + b for a, b in zip(o1, o2))
                if all(a <= m for a, m in zip(order, self.max_orders)):
                    out.coeffs[order] = out.coeffs[order] + c1 * c2 if order in out.coeffs else c1 * c2
        return out

    # arithmetic.jl:282-316
    def __pow__(self, p: int) -> "TaylorSeries":
        p = int(p)
        if p < 0:
            raise ValueError("negative power of a Taylor series")
        if p == 1:
            return TaylorSeries(self.max_orders, self.coeffs)
        if p == 0:
            return TaylorSeries(self.max_orders, {tuple(0 for _ in self.max_orders): 1.0})
        if p == 2:
            return self * self
        x = self
        t = (p & -p).bit_length()  # trailing_zeros(p) + 1
        p >>= t
        t -= 1
        while t > 0:
            x = x * x
            t -= 1
        y = x
        while p > 0:
            t = (p & -p).bit_length()
            p >>= t
            t -= 1
            while t >= 0:
                x = x * x
                t -= 1
            y = y * x
        return y


def set_variables(orders: Sequence[int]) -> List[TaylorSeries]:
    """parameter.jl set_variables: the series equal to each variable (numeric coefficients; used by the KATs)."""
    n = len(orders)
    return [TaylorSeries(orders, {tuple(1 if j == i else 0 for j in range(n)): 1.0}) for i in range(n)]


def _apply(op, series: List[TaylorSeries], factors: List[float]) -> TaylorSeries:
    """utility.jl:11-13: sum / prod are left folds of (d * f)."""
    if isinstance(op, Power):
        return (series[0] ** op.N) * factors[0]
    terms = [d * f for d, f in zip(series, factors)]
    acc = terms[0]
    for t in terms[1:]:
        acc = acc + t if isinstance(op, Sum) else acc * t
    return acc


def taylorexpansion(graphs: Sequence[Graph], var_dependence: Dict[int, List[bool]], max_orders: Sequence[int],
                    to_coeff_map: Dict[int, TaylorSeries] = None):
    """utility.jl:106-135 for every graph of a vector (:243-252).  Iterative post-order (the reference recurses)."""
    if to_coeff_map is None:
        to_coeff_map = {}
    nvars = len(max_orders)
    result = []
    for root in graphs:
        for node in post_order_unique([root]):
            if node.id in to_coeff_map:
                continue
            if node.isleaf():
                var = var_dependence.get(node.id, [False] * nvars)
                ranges = [range(0, max_orders[i] + 1) if var[i] else range(0, 1) for i in range(nvars)]
                ts = TaylorSeries(max_orders)
                # Iterators.product varies the FIRST index fastest
                for rev in itertools.product(*reversed(ranges)):
                    o = tuple(reversed(rev))
                    if sum(o) == 0:
                        ts.coeffs[o] = node  # the zero-order coefficient of a leaf is the leaf itself
                    else:
                        ts.coeffs[o] = Graph([], operator=Sum(), properties=node.properties, orders=list(o))
                to_coeff_map[node.id] = ts
            else:
                ts = _apply(node.operator, [to_coeff_map[s.id] for s in node.subgraphs], node.subgraph_factors)
                for g in ts.coeffs.values():
                    g.properties = node.properties
                to_coeff_map[node.id] = ts
        result.append(to_coeff_map[root.id])
    return result, to_coeff_map


def taylorAD(graphs: Sequence[Graph], deriv_orders: Sequence[int], leaf_dep_funcs: Sequence[Callable]) -> Dict[Orders, List[Graph]]:
    """utility.jl:48-93: Dict(orders => [coefficient graph of every input graph that has that order])."""
    assert len(deriv_orders) == len(leaf_dep_funcs)
    var_dependence: Dict[int, List[bool]] = {}
    for root in graphs:
        for node in post_order_unique([root]):
            if node.isleaf() and node.id not in var_dependence:
                var_dependence[node.id] = [bool(f(node.properties)) for f in leaf_dep_funcs]
    series, _ = taylorexpansion(graphs, var_dependence, deriv_orders)
    out: Dict[Orders, List[Graph]] = {}
    for ts in series:
        for orders, g in ts.coeffs.items():
            out.setdefault(orders, []).append(g)
    return out
