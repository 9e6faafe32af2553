"""Restated IR passes of the reference (`optimize!`, level 0).  TEST / WORKLOAD INFRASTRUCTURE.

These passes are producers of the evaluator's input, not part of the hot path: they are restated only so
that the repository can build the *real* workload graphs named in BASELINE.json without Julia.

Reference:
  optimize!                         src/computational_graph/optimize.jl:16-36
  remove_duplicated_leaves!         optimize.jl:289-317   (unique_nodes! :255-277, isequiv abstractgraph.jl:307-350)
  flatten_all_chains! / flatten_chains!          optimize.jl:53-75, transform.jl:354-364
  merge_all_linear_combinations! / merge_linear_combination!   optimize.jl:117-139, transform.jl:472-497
  remove_all_zero_valued_subgraphs! / remove_zero_valued_subgraphs!   optimize.jl:85-107, transform.jl:388-448

The reference walks the tree expansion of the DAG (shared nodes are re-visited); every pass is idempotent per
node, so visiting each node object once in post-order gives the same result -- that is what is done here.
The O(n^2) `isequiv` scans are bucketed by a structural hash that is equal for equivalent nodes; candidates
inside a bucket are still confirmed with the restated `isequiv`, in the reference's scan order.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

from fdgraph_b200.graph import Graph, Power, Prod, Sum, post_order_unique, unary_istrivial


def _prop_key(p):
    if p is None:
        return None
    k = getattr(p, "key", None)
    return k() if callable(k) else p


def _approx(a, b) -> bool:
    return abs(a - b) <= 1.4901161193847656e-08 * max(abs(a), abs(b))


def isequiv(a: Graph, b: Graph, *ignore: str) -> bool:
    """abstractgraph.jl:307-350."""
    if type(a) is not type(b):
        return False
    if "weight" not in ignore and not (a.weight == b.weight or _approx(a.weight, b.weight)):
        return False
    if len(a.subgraph_factors) != len(b.subgraph_factors):
        return False
    if "id" not in ignore and a.id != b.id:
        return False
    if "name" not in ignore and a.name != b.name:
        return False
    if "orders" not in ignore and list(a.orders) != list(b.orders):
        return False
    if "operator" not in ignore and a.operator != b.operator:
        return False
    if "properties" not in ignore and _prop_key(a.properties) != _prop_key(b.properties):
        return False
    b_pairs = list(zip(b.subgraphs, b.subgraph_factors))
    for suba, fa in zip(a.subgraphs, a.subgraph_factors):
        for idx, (subb, fb) in enumerate(b_pairs):
            if fa == fb and (suba is subb or isequiv(suba, subb, *ignore)):
                del b_pairs[idx]
                break
        else:
            return False
    return True


def _node_hash(n: Graph, child_hash: Dict[int, int], with_name: bool) -> int:
    """Structural hash ignoring id (and optionally name): equal for isequiv-equivalent nodes."""
    kids = tuple(sorted((child_hash[id(c)], f) for c, f in zip(n.subgraphs, n.subgraph_factors)))
    return hash((n.name if with_name else None, tuple(n.orders), n.operator, _prop_key(n.properties), kids))


def unique_nodes(graphs: Sequence[Graph]) -> Dict[int, Graph]:
    """optimize.jl:255-277 with isequiv(e, g, :id, :name, :weight): id -> representative (first equivalent)."""
    mapping: Dict[int, Graph] = {}
    hv: Dict[int, int] = {}
    buckets: Dict[int, List[Graph]] = {}
    for g in graphs:
        for n in post_order_unique([g]):
            if id(n) not in hv:
                hv[id(n)] = _node_hash(n, hv, with_name=False)
        bucket = buckets.setdefault(hv[id(g)], [])
        for e in bucket:
            if isequiv(e, g, "id", "name", "weight"):
                mapping[g.id] = e
                break
        else:
            bucket.append(g)
            mapping[g.id] = g
    return mapping


def _leaves_tree_order(g: Graph) -> List[Graph]:
    return [n for n in post_order_unique([g]) if n.isleaf()]


def remove_duplicated_leaves(graphs: Sequence[Graph]) -> Sequence[Graph]:
    """optimize.jl:289-317."""
    leaves: List[Graph] = []
    for g in graphs:
        leaves.extend(_leaves_tree_order(g))
    leaves.sort(key=lambda x: x.id)
    uniq, seen = [], set()
    for l in leaves:
        if l.id not in seen:
            seen.add(l.id)
            uniq.append(l)
    mapping = unique_nodes(uniq)
    for n in post_order_unique(graphs):
        for si, sub in enumerate(n.subgraphs):
            if sub.isleaf():
                n.subgraphs[si] = mapping[sub.id]
    return graphs


def flatten_chains(g: Graph) -> Graph:
    """transform.jl:354-364."""
    for i, sub in enumerate(g.subgraphs):
        if unary_istrivial(sub.operator) and sub.onechild():
            flatten_chains(sub)
            g.subgraph_factors[i] = g.subgraph_factors[i] * sub.subgraph_factors[0]
            g.subgraphs[i] = sub.subgraphs[0]
    return g


def flatten_all_chains(graphs: Sequence[Graph]) -> Sequence[Graph]:
    for n in post_order_unique(graphs):
        flatten_chains(n)
    return graphs


def merge_linear_combination(g: Graph, hv: Dict[int, int] = None) -> Graph:
    """transform.jl:472-497: in a Sum, later subgraphs equivalent (isequiv(.., :id)) to an earlier one are merged."""
    if not isinstance(g.operator, Sum):
        return g
    if hv is None:
        hv = {}
        for n in post_order_unique(list(g.subgraphs)):
            hv[id(n)] = _node_hash(n, hv, with_name=True)
    merged: List[Graph] = []
    mfac: List[float] = []
    first_of: Dict[int, List[int]] = {}
    for s, f in zip(g.subgraphs, g.subgraph_factors):
        for k in first_of.get(hv[id(s)], ()):
            if merged[k] is s or isequiv(merged[k], s, "id"):
                mfac[k] += f
                break
        else:
            first_of.setdefault(hv[id(s)], []).append(len(merged))
            merged.append(s)
            mfac.append(f)
    g.subgraphs = merged
    g.subgraph_factors = mfac
    return g


def merge_all_linear_combinations(graphs: Sequence[Graph]) -> Sequence[Graph]:
    hv: Dict[int, int] = {}
    for n in post_order_unique(graphs):
        if isinstance(n.operator, Sum) and n.subgraphs:
            merge_linear_combination(n, hv)
        hv[id(n)] = _node_hash(n, hv, with_name=True)
    return graphs


def _has_zero_subfactors(g: Graph) -> bool:
    """tree_properties.jl:77-95."""
    op = g.operator
    if isinstance(op, Sum):
        return all(f == 0 for f in g.subgraph_factors)
    if isinstance(op, Prod):
        return any(f == 0 for f in g.subgraph_factors)
    if isinstance(op, Power):
        return g.subgraph_factors[0] == 0
    return False


def remove_zero_valued_subgraphs(g: Graph) -> Graph:
    """transform.jl:426-448 with mask_zero_subgraph_factors :388-416."""
    if g.isleaf() or (g.onechild() and g.subgraphs[0].isleaf()):
        return g
    subg = list(g.subgraphs)
    fac = list(g.subgraph_factors)
    for i, s in enumerate(subg):
        if s.isleaf():
            continue
        if _has_zero_subfactors(s):
            fac[i] = 0.0
    op = g.operator
    if isinstance(op, Sum):
        mask = [i for i, f in enumerate(fac) if f != 0] or [0]
    elif isinstance(op, Prod):
        z = next((i for i, f in enumerate(fac) if f == 0), None)
        mask = list(range(len(fac))) if z is None else [z]
    elif isinstance(op, Power):
        mask = [0]
    else:
        mask = list(range(len(fac)))
    g.subgraphs = [subg[i] for i in mask]
    g.subgraph_factors = [fac[i] for i in mask]
    return g


def remove_all_zero_valued_subgraphs(graphs: Sequence[Graph]) -> Sequence[Graph]:
    for n in post_order_unique(graphs):
        remove_zero_valued_subgraphs(n)
    return graphs


def remove_duplicated_nodes(root: Graph) -> Graph:
    """optimize.jl:345-390 (the method optimize!(level > 0) calls on Graph(graphs)): top-down, a node equivalent
    (isequiv modulo id / name / weight: same operator, orders, properties and the same MULTISET of (subgraph, factor)
    pairs) to one already kept is replaced by it in its parent; the search over the kept nodes is by structural hash here
    instead of the reference's scan of all of them."""
    unique: Dict[int, Graph] = {}
    buckets: Dict[int, List[Graph]] = {}
    hv: Dict[int, int] = {}
    for n in post_order_unique([root]):
        hv[id(n)] = _node_hash(n, hv, with_name=False)

    def process(node: Graph) -> Graph:
        # (explicit stack: the graphs are deeper than Python's recursion limit likes)
        result: Dict[int, Graph] = {}
        stack = [(node, 0)]
        while stack:
            n, i = stack.pop()
            if i == 0:
                if n.id in unique:
                    result[id(n)] = unique[n.id]
                    continue
                rep = next((g for g in buckets.get(hv[id(n)], ()) if isequiv(n, g, "id", "name", "weight")), None)
                if rep is not None:
                    result[id(n)] = rep
                    continue
            if i > 0:
                n.subgraphs[i - 1] = result[id(n.subgraphs[i - 1])]
            if i < len(n.subgraphs):
                stack.append((n, i + 1))
                stack.append((n.subgraphs[i], 0))
                continue
            unique[n.id] = n
            buckets.setdefault(hv[id(n)], []).append(n)
            result[id(n)] = n
        return result[id(node)]

    process(root)
    return root


def optimize(graphs: Sequence[Graph], level: int = 0) -> Sequence[Graph]:
    """`optimize!(graphs)` (optimize.jl:16-36).  Level 0 is what every example and front-end test of the reference uses;
    level > 0 merges equivalent inner nodes too (remove_duplicated_nodes! on a root above the graphs)."""
    if not graphs:
        return graphs
    if level > 0:
        remove_duplicated_nodes(Graph(list(graphs)))
    else:
        remove_duplicated_leaves(graphs)
    flatten_all_chains(graphs)
    merge_all_linear_combinations(graphs)
    remove_all_zero_valued_subgraphs(graphs)
    return graphs
