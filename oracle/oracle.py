"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see the header of fdg_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Four independent evaluators of the same function:

* ``Oracle.eval(..., mode="emitter")``  -- C, statement-for-statement the emitted function
  (reference src/backend/static.jl:13-46,98-133), the bit-exact target of the CUDA path;
* ``Oracle.eval(..., mode="interp")``   -- C, the interpreter's rounding (src/computational_graph/eval.jl:1-39);
* ``eval_interp_py``                    -- pure-Python restatement of ``eval!`` walking Graph objects
  (eval.jl:15-39): the evaluator every value test of the reference uses;
* ``eval_exact``                        -- exact rational arithmetic (fractions.Fraction), the yardstick
  for conditioning-aware error bounds.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from fractions import Fraction
from typing import Dict, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libfdg_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "fdg_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        proc = subprocess.run(["make", "-C", HERE, "-B" if force else "-s"], capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("building the oracle failed:\n" + proc.stdout + proc.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        L.oracle_lower.restype = vp
        L.oracle_lower.argtypes = [i64, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp]
        L.oracle_free.argtypes = [vp]
        L.oracle_num_leaves.restype = i64
        L.oracle_num_leaves.argtypes = [vp]
        L.oracle_num_stmts.restype = i64
        L.oracle_num_stmts.argtypes = [vp]
        L.oracle_last_root.restype = i32
        L.oracle_last_root.argtypes = [vp]
        L.oracle_leafmap.argtypes = [vp, vp]
        L.oracle_stmt.argtypes = [vp, i64] + [vp] * 6
        L.oracle_operand.argtypes = [vp, i64, vp, vp]
        L.oracle_eval_batch.restype = C.c_int
        L.oracle_eval_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, i64, vp, i64, i64, C.c_int]
        L.oracle_run_emitted.restype = C.c_int
        L.oracle_run_emitted.argtypes = [vp, vp, i64, vp, i64, i64, C.c_int]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


class Oracle:
    """Lowered (emitter-order) program over a RawGraph-like object (arrays as in include/fdgraph.h)."""

    def __init__(self, raw):
        L = lib()
        self.raw = raw
        a = [np.ascontiguousarray(raw.node_id, np.int64), np.ascontiguousarray(raw.node_op, np.int32),
             np.ascontiguousarray(raw.node_pow, np.int32), np.ascontiguousarray(raw.child_ptr, np.int64),
             np.ascontiguousarray(raw.child_node, np.int32), np.ascontiguousarray(raw.child_factor, np.float64),
             np.ascontiguousarray(raw.graphs, np.int32), np.ascontiguousarray(raw.root_id, np.int64)]
        self._keep = a
        self._p = L.oracle_lower(len(a[0]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data,
                                 a[4].ctypes.data, a[5].ctypes.data, len(a[6]), a[6].ctypes.data, len(a[7]),
                                 a[7].ctypes.data)
        if not self._p:
            raise ValueError("oracle: malformed graph (cycle, bad index, unknown operator or Power N < 2)")
        self.n_leaves = int(L.oracle_num_leaves(self._p))
        self.n_stmts = int(L.oracle_num_stmts(self._p))
        self.n_roots = len(a[7])
        self.last_root = int(L.oracle_last_root(self._p))
        lm = np.empty(max(self.n_leaves, 1), np.int32)
        L.oracle_leafmap(self._p, lm.ctypes.data)
        self.leaf_nodes = lm[: self.n_leaves]

    def __del__(self):
        try:
            if self._p:
                lib().oracle_free(self._p)
                self._p = None
        except Exception:
            pass

    def statements(self):
        """[(op, pow_n, [(val, factor), ...], root, leaf)] in emitter order."""
        L = lib()
        out = []
        op, pw, cnt, root, leaf = (C.c_int32() for _ in range(5))
        first = C.c_int64()
        val, f = C.c_int32(), C.c_double()
        for i in range(self.n_stmts):
            L.oracle_stmt(self._p, i, C.byref(op), C.byref(pw), C.byref(first), C.byref(cnt), C.byref(root), C.byref(leaf))
            ops = []
            if op.value >= 0:
                for e in range(first.value, first.value + cnt.value):
                    L.oracle_operand(self._p, e, C.byref(val), C.byref(f))
                    ops.append((val.value, f.value))
            out.append((op.value, pw.value, ops, root.value, leaf.value))
        return out

    def eval(self, leaf: np.ndarray, mode: str = "emitter", layout: str = "batch", nthreads: int = 1,
             root: Optional[np.ndarray] = None) -> np.ndarray:
        """layout "batch": leaf (L, B) -> root (R, B);  layout "sample": leaf (B, L) -> root (B, R)."""
        leaf = np.ascontiguousarray(leaf)
        cplx = leaf.dtype == np.complex128
        assert leaf.dtype in (np.float64, np.complex128)
        if layout == "batch":
            Lrows, B = leaf.shape
            assert Lrows >= self.n_leaves
            shape = (self.n_roots, B)
        else:
            B, Lcols = leaf.shape
            assert Lcols >= self.n_leaves
            shape = (B, self.n_roots)
        if root is None:
            root = np.zeros(shape, leaf.dtype)
        assert root.shape == shape and root.flags.c_contiguous and root.dtype == leaf.dtype
        ld_leaf = leaf.shape[1]
        ld_root = root.shape[1] if root.ndim == 2 else 1
        lib().oracle_eval_batch(self._p, int(cplx), 0 if mode == "emitter" else 1, 0 if layout == "batch" else 1,
                                leaf.ctypes.data, ld_leaf, root.ctypes.data, max(ld_root, 1), B, nthreads)
        return root


# ---------------------------------------------------------------------------------------------------
# pure-Python restatements (small cases)
# ---------------------------------------------------------------------------------------------------


def _post_order_tree(g):
    """AbstractTrees.PostOrderDFS over the *tree expansion* (shared nodes are re-visited), children in
    stored order -- what eval! iterates (eval.jl:20)."""
    stack = [(g, 0)]
    while stack:
        node, i = stack.pop()
        if i < len(node.subgraphs):
            stack.append((node, i + 1))
            stack.append((node.subgraphs[i], 0))
        else:
            yield node


def eval_interp_py(g, leafmap: Optional[Dict[int, int]] = None, leaf: Optional[Sequence] = None, one=1.0):
    """``eval!(g, leafmap, leaf)`` (eval.jl:15-39): leafmap is id -> 0-based index into ``leaf``; leaves default
    to 1.0 when no leafmap is given.  Stores ``node.weight`` like the reference and returns the root weight."""
    from fdgraph_b200.graph import Power, Prod, Sum  # the node model under test

    result = None
    for node in _post_order_tree(g):
        if not node.subgraphs:
            node.weight = one if not leafmap else leaf[leafmap[node.id]]
        else:
            op = node.operator
            terms = [d.weight * f for d, f in zip(node.subgraphs, node.subgraph_factors)]
            if isinstance(op, Sum):
                acc = terms[0]
                for t in terms[1:]:
                    acc = acc + t
            elif isinstance(op, Prod):
                acc = terms[0]
                for t in terms[1:]:
                    acc = acc * t
            elif isinstance(op, Power):
                acc = node.subgraphs[0].weight ** op.N * node.subgraph_factors[0]
            else:
                raise NotImplementedError(op)
            node.weight = acc
        result = node.weight
    return result


def eval_exact(oracle: Oracle, leaf_row: Sequence[float]):
    """Exact rational value of every root for ONE real sample: returns list of Fraction (None if unset)."""
    vals = []
    roots = [None] * oracle.n_roots
    for op, pw, ops, root, leaf in oracle.statements():
        if op < 0:
            v = Fraction(float(leaf_row[leaf]))
        elif op == 1:
            v = sum((vals[c] * Fraction(f) for c, f in ops), Fraction(0))
        elif op == 2:
            v = Fraction(1)
            for c, f in ops:
                v *= vals[c] * Fraction(f)
        else:
            v = vals[ops[0][0]] ** pw * Fraction(ops[0][1])
        vals.append(v)
        if root >= 0:
            roots[root] = v
    return roots


def eval_abs_bound(oracle: Oracle, leaf_row: Sequence[float]):
    """Value of every root with every quantity replaced by its absolute value: the natural scale for the
    rounding error of a cancelling sum (|computed - exact| <= k * eps * this)."""
    vals = []
    roots = [None] * oracle.n_roots
    for op, pw, ops, root, leaf in oracle.statements():
        if op < 0:
            v = abs(float(leaf_row[leaf]))
        elif op == 1:
            v = sum(vals[c] * abs(f) for c, f in ops)
        elif op == 2:
            v = 1.0
            for c, f in ops:
                v *= vals[c] * abs(f)
        else:
            v = vals[ops[0][0]] ** pw * abs(ops[0][1])
        vals.append(v)
        if root >= 0:
            roots[root] = v
    return roots
