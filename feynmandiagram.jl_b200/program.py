"""Flattened ("raw") form of a Graph DAG: the arrays behind ``fdg_graph_desc`` (include/fdgraph.h).

The Julia / Python host only walks its node objects once through the public getters
``id / operator / subgraphs / subgraph_factors`` (the same four the reference emitter uses,
src/backend/static.jl:106-124) and hands plain arrays to the native library; statement order, leaf
numbering and everything else is decided natively (csrc/fdg_lower.cpp).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .graph import Graph, Power, Prod, Sum, Unitary, post_order_unique

OP_UNITARY, OP_SUM, OP_PROD, OP_POWER = 0, 1, 2, 3


@dataclass
class RawGraph:
    node_id: np.ndarray       # int64 [n]
    node_op: np.ndarray       # int32 [n]
    node_pow: np.ndarray      # int32 [n]
    child_ptr: np.ndarray     # int64 [n+1]
    child_node: np.ndarray    # int32 [e]
    child_factor: np.ndarray  # float64 [e]
    graphs: np.ndarray        # int32 [g]
    root_id: np.ndarray       # int64 [r]

    @property
    def n_nodes(self) -> int:
        return int(self.node_id.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.child_node.shape[0])

    def save(self, path: str) -> None:
        np.savez_compressed(path, **{k: getattr(self, k) for k in self.__dataclass_fields__})

    @staticmethod
    def load(path: str) -> "RawGraph":
        with np.load(path) as z:
            return RawGraph(**{k: np.ascontiguousarray(z[k]) for k in RawGraph.__dataclass_fields__})

    # the FDGRAPH file of include/fdgraph.h (fdg_graph_write / fdg_compile_file): what a Julia session writes
    _FDG_FIELDS = (("node_id", np.int64), ("node_op", np.int32), ("node_pow", np.int32), ("child_ptr", np.int64),
                   ("child_node", np.int32), ("child_factor", np.float64), ("graphs", np.int32), ("root_id", np.int64))

    def save_fdg(self, path: str) -> None:
        from . import _capi

        _capi.graph_write(self, path)

    @staticmethod
    def load_fdg(path: str) -> "RawGraph":
        with open(path, "rb") as fh:
            data = fh.read()
        if data[:8] != b"FDGRAPH\x01":
            raise ValueError(f"{path} is not an FDGRAPH file")
        n, e, g, r = (int(x) for x in np.frombuffer(data, "<i8", 4, 8))
        counts = {"node_id": n, "node_op": n, "node_pow": n, "child_ptr": n + 1, "child_node": e, "child_factor": e, "graphs": g, "root_id": r}
        off, out = 40, {}
        for name, dt in RawGraph._FDG_FIELDS:
            out[name] = np.frombuffer(data, np.dtype(dt).newbyteorder("<"), counts[name], off).astype(dt)
            off += counts[name] * np.dtype(dt).itemsize
        if off != len(data):
            raise ValueError(f"{path} is truncated or has trailing bytes")
        return RawGraph(**out)

    def to_graphs(self) -> List[Graph]:
        """Inverse of :func:`flatten`: Graph objects (ids preserved, shared nodes shared) for the entries of `graphs`."""
        n = self.n_nodes
        state = np.zeros(n, np.int8)
        order: List[int] = []
        for r in range(n):  # children before parents, whatever the order of the arrays
            if state[r]:
                continue
            stack = [(r, 0)]
            while stack:
                v, k = stack.pop()
                if k == 0:
                    if state[v]:
                        continue
                    state[v] = 1
                lo, hi = int(self.child_ptr[v]), int(self.child_ptr[v + 1])
                if lo + k < hi:
                    stack.append((v, k + 1))
                    c = int(self.child_node[lo + k])
                    if not state[c]:
                        stack.append((c, 0))
                else:
                    order.append(v)
        nodes: List[Optional[Graph]] = [None] * n
        ops = {OP_UNITARY: Unitary(), OP_SUM: Sum(), OP_PROD: Prod()}
        for i in order:
            lo, hi = int(self.child_ptr[i]), int(self.child_ptr[i + 1])
            subs = [nodes[int(c)] for c in self.child_node[lo:hi]]
            op = Power(int(self.node_pow[i])) if int(self.node_op[i]) == OP_POWER else ops[int(self.node_op[i])]
            g = Graph(subs, subgraph_factors=[float(f) for f in self.child_factor[lo:hi]], operator=op if subs else Sum())
            g.id = int(self.node_id[i])
            nodes[i] = g
        return [nodes[int(i)] for i in self.graphs]

    def validate_dtypes(self) -> "RawGraph":
        self.node_id = np.ascontiguousarray(self.node_id, dtype=np.int64)
        self.node_op = np.ascontiguousarray(self.node_op, dtype=np.int32)
        self.node_pow = np.ascontiguousarray(self.node_pow, dtype=np.int32)
        self.child_ptr = np.ascontiguousarray(self.child_ptr, dtype=np.int64)
        self.child_node = np.ascontiguousarray(self.child_node, dtype=np.int32)
        self.child_factor = np.ascontiguousarray(self.child_factor, dtype=np.float64)
        self.graphs = np.ascontiguousarray(self.graphs, dtype=np.int32)
        self.root_id = np.ascontiguousarray(self.root_id, dtype=np.int64)
        return self


def _opcode(op) -> int:
    if isinstance(op, Sum):
        return OP_SUM
    if isinstance(op, Prod):
        return OP_PROD
    if isinstance(op, Power):
        return OP_POWER
    if isinstance(op, Unitary):
        return OP_UNITARY
    # static.jl:6-11: "Static representation for computational graph nodes with operator X not yet implemented"
    raise NotImplementedError(
        f"Static representation for computational graph nodes with operator {op!r} not yet implemented!")


def flatten(graphs: Sequence[Graph], root: Optional[Sequence[int]] = None):
    """Returns (RawGraph, nodes) where ``nodes[i]`` is the Graph object behind node index i."""
    graphs = list(graphs)
    nodes: List[Graph] = post_order_unique(graphs)
    index = {id(n): i for i, n in enumerate(nodes)}
    n = len(nodes)
    node_id = np.empty(n, np.int64)
    node_op = np.empty(n, np.int32)
    node_pow = np.zeros(n, np.int32)
    child_ptr = np.zeros(n + 1, np.int64)
    child_node: List[int] = []
    child_factor: List[float] = []
    for i, g in enumerate(nodes):
        node_id[i] = g.id
        if g.subgraphs:
            node_op[i] = _opcode(g.operator)
        else:
            node_op[i] = OP_UNITARY if isinstance(g.operator, Unitary) else _opcode(g.operator)
        if isinstance(g.operator, Power):
            node_pow[i] = g.operator.N
        for c, f in zip(g.subgraphs, g.subgraph_factors):
            child_node.append(index[id(c)])
            child_factor.append(float(f))
        child_ptr[i + 1] = len(child_node)
    if root is None:
        root = [g.id for g in graphs]
    raw = RawGraph(node_id, node_op, node_pow, child_ptr,
                   np.asarray(child_node, np.int32), np.asarray(child_factor, np.float64),
                   np.asarray([index[id(g)] for g in graphs], np.int32),
                   np.asarray(list(root), np.int64))
    return raw.validate_dtypes(), nodes
