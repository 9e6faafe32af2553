"""Host-side mirror of the reference's ``Compilers`` module for the graph-evaluation path.

    eval_graph, leafmap = compile(graphs; root=[id(g) for g in graphs])     # src/backend/static.jl:221-227
    eval_graph(root, leafVal)                                               # generated function, static.jl:100-131

Same two-value return, same leaf numbering, same in-place ``root`` semantics and return value (the
root written last).  What changes is *where* it runs: the callable evaluates a whole batch of
Monte-Carlo samples on a B200 through libfdgraph.so (include/fdgraph.h).

Layouts.  ``leafVal`` is (B, L) and ``root`` is (B, R) with the batch index unit-stride -- a Julia
``Matrix`` of that shape, ``np.asfortranarray`` or a torch tensor with strides (1, ld) -- the layout
of the reference's batched torch emitter (``leafVal[:, k]``, compiler_python.jl:23,28).  1-D arrays
are the single-sample call of the Julia / C emitters.  Host arrays go through ``fdg_eval_host``
(H2D, kernel, D2H); CUDA tensors are evaluated in place on the current torch stream.

There is NO CPU implementation behind this module: without libfdgraph.so or without a CUDA device
every evaluation raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from .emitters import (compile_C, compile_Julia, compile_Python, julia_to_C_typestr, to_Cstr, to_julia_str,  # noqa: F401
                       to_python_str, to_static)
from .graph import Graph
from .program import RawGraph, flatten

_DTYPES = {np.dtype(np.float64): _capi.FDG_F64, np.dtype(np.complex128): _capi.FDG_C128}


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _batch_major(a2: np.ndarray):
    """(B, K) array -> ((K, B) array with unit batch stride, leading dimension, is_view)."""
    t = a2.T
    isz = a2.itemsize
    ok = (t.shape[1] <= 1 or t.strides[1] == isz) and \
        (t.shape[0] <= 1 or (t.strides[0] % isz == 0 and t.strides[0] >= t.shape[1] * isz))
    view = ok
    if not ok:
        t = np.ascontiguousarray(t)
    ld = t.strides[0] // isz if t.shape[0] > 1 else max(t.shape[1], 1)
    return t, ld, view


class Evaluator:
    """The callable returned by :func:`compile` -- stands in for the generated ``eval_graph!``."""

    def __init__(self, raw: RawGraph, dtype=np.float64, max_slots: int = 0, prefetch: int = 0, schedule: int = 0,
                 backend: int = 0, jit_segment: int = 0, cse=None, fma: bool = False):
        self.dtype = np.dtype(dtype)
        if self.dtype not in _DTYPES:
            # static.jl:151  error("Unsupported type")
            raise TypeError(f"Unsupported type {self.dtype}: libfdgraph evaluates float64 or complex128 weights")
        self.raw = raw
        self._h = _capi.compile_raw(raw, _DTYPES[self.dtype], max_slots, prefetch, schedule, backend, jit_segment, cse, fma)
        self.stats = _capi.stats(self._h)
        self.n_leaves = self.stats["n_leaves"]
        self.n_roots = self.stats["n_roots"]
        self.leaf_nodes = _capi.leafmap(self._h, self.n_leaves)
        self.last_root = _capi.last_root(self._h)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            _capi.lib().fdg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def program_words(self) -> np.ndarray:
        return _capi.program_words(self._h)

    def jit_prepare(self, samples_per_thread: int = 2, accumulate: bool = False) -> dict:
        """Generate + assemble the specialised kernels now (host only)."""
        return _capi.jit_prepare(self._h, samples_per_thread, accumulate)

    def jit_last(self) -> dict:
        """Plan of the specialised kernels the last launch of this evaluator ran (which form, kernels, traffic)."""
        return _capi.jit_info(self._h, 0, False)

    def pipeline_prepare(self, accumulate: bool = True, n_sm: int = 148) -> dict:
        """Build the pipeline form of the specialised kernels for a device with ``n_sm`` SMs (host only); its plan."""
        return _capi.pipeline_prepare(self._h, accumulate, n_sm)

    def pipeline_stats(self, stream: int = 0, n_stages: int = 0) -> dict:
        """Clocks every stage of the last pipeline launch on ``stream`` spent alive / waiting (synchronises the stream)."""
        return _capi.pipeline_stats(self._h, stream, n_stages)

    def jit_ptx(self, samples_per_thread: int = 2, accumulate: bool = False, index: int = 0):
        return _capi.jit_ptx(self._h, samples_per_thread, accumulate, index)

    def set_launch(self, threads: int = 0, samples_per_thread: int = 0, blocks_per_sm: int = 0) -> None:
        _capi.check(_capi.lib().fdg_set_launch(self._h, threads, samples_per_thread, blocks_per_sm))

    @property
    def launches(self) -> int:
        return _capi.launch_count(self._h)

    # -- evaluation -----------------------------------------------------------------------------
    def __call__(self, root, leafVal):
        """``eval_graph!(root, leafVal)``: fills ``root`` in place, returns the last root written."""
        if _is_torch(leafVal) or _is_torch(root):
            return self._call_torch(root, leafVal)
        return self._call_numpy(root, leafVal)

    def _call_numpy(self, root, leafVal):
        single = np.ndim(leafVal) == 1
        lv = np.asarray(leafVal, dtype=self.dtype)
        if not isinstance(root, np.ndarray) or root.dtype != self.dtype:
            raise TypeError(f"root must be a numpy array of dtype {self.dtype} (it is written in place)")
        if single:
            if lv.shape[0] < self.n_leaves or root.shape[0] < self.n_roots or root.ndim != 1:
                raise IndexError("leafVal / root shorter than the number of leaves / roots")  # Julia BoundsError
            lv2 = np.ascontiguousarray(lv[: self.n_leaves]).reshape(self.n_leaves, 1)
            r2 = np.array(root[: self.n_roots]).reshape(self.n_roots, 1)
            self._eval_host(lv2, 1, r2, 1, 1)
            root[: self.n_roots] = r2[:, 0]
            return root[self.last_root] if self.last_root >= 0 else None
        if lv.ndim != 2 or root.ndim != 2 or lv.shape[0] != root.shape[0]:
            raise ValueError("expected leafVal of shape (B, L) and root of shape (B, R)")
        B = lv.shape[0]
        if lv.shape[1] < self.n_leaves or root.shape[1] < self.n_roots:
            raise IndexError("leafVal / root have fewer columns than leaves / roots")
        # batch-major storage: element (b, k) at k*ld + b.  (B, L) Fortran order has exactly that;
        # anything else costs one transposing copy.
        lvT, ld_leaf, _ = _batch_major(lv)
        out, ld_root, direct = _batch_major(root)
        self._eval_host(lvT, ld_leaf, out, ld_root, B)
        if not direct:
            root[:, : self.n_roots] = out.T[:, : self.n_roots]
        return root[:, self.last_root] if self.last_root >= 0 else None

    def _eval_host(self, leafT: np.ndarray, ld_leaf: int, rootT: np.ndarray, ld_root: int, batch: int) -> None:
        _capi.check(_capi.lib().fdg_eval_host(self._h, leafT.ctypes.data, ld_leaf, rootT.ctypes.data, ld_root, batch))

    def _call_torch(self, root, leafVal):
        import torch

        if not (leafVal.is_cuda and root.is_cuda):
            raise TypeError("torch tensors must both live on a CUDA device (pass numpy arrays for host data)")
        want = torch.float64 if self.dtype == np.float64 else torch.complex128
        if leafVal.dtype != want or root.dtype != want:
            raise TypeError(f"expected {want} tensors")
        if leafVal.dim() == 1:
            leafVal, root2 = leafVal.unsqueeze(0), root.unsqueeze(0)
        else:
            root2 = root
        B = leafVal.shape[0]
        if root2.shape[0] != B or leafVal.shape[1] < self.n_leaves or root2.shape[1] < self.n_roots:
            raise ValueError("expected leafVal of shape (B, L) and root of shape (B, R)")
        if B > 1 and (leafVal.stride(0) != 1 or root2.stride(0) != 1):
            raise ValueError("device tensors must be batch-major: shape (B, L) with strides (1, ld); "
                             "allocate as torch.empty(L, B).T")
        ld_leaf = leafVal.stride(1) if leafVal.shape[1] > 1 else max(B, 1)
        ld_root = root2.stride(1) if root2.shape[1] > 1 else max(B, 1)
        with torch.cuda.device(leafVal.device):
            stream = torch.cuda.current_stream().cuda_stream
            _capi.check(_capi.lib().fdg_eval(self._h, leafVal.data_ptr(), ld_leaf, root2.data_ptr(), ld_root, B, stream))
        return root2[:, self.last_root] if self.last_root >= 0 else None

    def eval_device(self, leaf_ptr: int, ld_leaf: int, root_ptr: int, ld_root: int, batch: int, stream: int = 0) -> None:
        """Raw device-pointer form of ``fdg_eval``."""
        _capi.check(_capi.lib().fdg_eval(self._h, leaf_ptr, ld_leaf, root_ptr, ld_root, batch, stream))

    def accumulate_device(self, leaf_ptr: int, ld_leaf: int, batch: int, acc_ptr: int, stream: int = 0) -> None:
        """``fdg_eval_accumulate``: acc[r] += sum over the batch of root r (deterministic, on device)."""
        _capi.check(_capi.lib().fdg_eval_accumulate(self._h, leaf_ptr, ld_leaf, batch, acc_ptr, stream))

    def accumulate(self, leafVal, acc) -> None:
        """torch form: ``leafVal`` (B, L) batch-major CUDA tensor, ``acc`` float64 CUDA tensor of R (2R) sums."""
        import torch

        B = leafVal.shape[0]
        ld_leaf = leafVal.stride(1) if leafVal.shape[1] > 1 else max(B, 1)
        if B > 1 and leafVal.stride(0) != 1:
            raise ValueError("device tensors must be batch-major: shape (B, L) with strides (1, ld)")
        with torch.cuda.device(leafVal.device):
            stream = torch.cuda.current_stream().cuda_stream
            self.accumulate_device(leafVal.data_ptr(), ld_leaf, B, acc.data_ptr(), stream)


def compile(graphs: Sequence[Graph], root: Optional[Sequence[int]] = None, *, dtype=np.float64,
            max_slots: int = 0, prefetch: int = 0, schedule: int = 0, backend: int = 0,
            jit_segment: int = 0, cse=None, fma: bool = False) -> Tuple[Evaluator, Dict[int, Graph]]:
    """``Compilers.compile(graphs; root)`` (static.jl:221-227) -> ``(eval_graph, leafmap)``.

    ``leafmap[k]`` is the leaf Graph whose value is read from column ``k`` of ``leafVal`` (0-based
    here; the reference's Dict is 1-based, static.jl:117-119).
    """
    raw, nodes = flatten(graphs, root)
    ev = Evaluator(raw, dtype=dtype, max_slots=max_slots, prefetch=prefetch, schedule=schedule, backend=backend,
                   jit_segment=jit_segment, cse=cse, fma=fma)
    leafmap = {k: nodes[int(i)] for k, i in enumerate(ev.leaf_nodes)}
    return ev, leafmap


def compile_raw(raw: RawGraph, *, dtype=np.float64, max_slots: int = 0, prefetch: int = 0, schedule: int = 0,
                backend: int = 0, jit_segment: int = 0, cse=None, fma: bool = False) -> Evaluator:
    """Compile an already flattened graph (e.g. a workload file written by another host).  ``fma=True`` (opt-in) lets
    the specialised kernels fuse multiplies into adds: faster where FP64 issue is the limit, NOT bit-identical."""
    return Evaluator(raw, dtype=dtype, max_slots=max_slots, prefetch=prefetch, schedule=schedule, backend=backend,
                     jit_segment=jit_segment, cse=cse, fma=fma)


def compile_file(path: str, *, dtype=np.float64, backend: int = 0, jit_segment: int = 0, cse=None) -> Evaluator:
    """Evaluator of a graph stored as an FDGRAPH file (fdg_compile_file): a graph flattened by a Julia session elsewhere
    (`FDGraphB200.save_graph`) or by `RawGraph.save_fdg`."""
    ev = Evaluator.__new__(Evaluator)
    ev.dtype = np.dtype(dtype)
    if ev.dtype not in _DTYPES:
        raise TypeError(f"Unsupported type {ev.dtype}: libfdgraph evaluates float64 or complex128 weights")
    ev._h = _capi.compile_file(path, _DTYPES[ev.dtype], backend, jit_segment, cse)  # the library reads the file itself
    ev.raw = RawGraph.load_fdg(path)
    ev.stats = _capi.stats(ev._h)
    ev.n_leaves, ev.n_roots = ev.stats["n_leaves"], ev.stats["n_roots"]
    ev.leaf_nodes = _capi.leafmap(ev._h, ev.n_leaves)
    ev.last_root = _capi.last_root(ev._h)
    return ev


class LeafGenerator:
    """Leaf values computed on the device from the Monte-Carlo variables (include/fdgraph.h, fdg_leafgen_*): what the
    integrand of example/benchmark.jl:44-81 does for every sample with the metadata of `leafstates`
    (src/frontend/frontends.jl:175-232).  `meta` holds leaf_type, leaf_order (L, 2), tau_in, tau_out, loop_index (0-based)
    and loop_basis (n_basis, n_loops)."""

    def __init__(self, meta, dim: int = 3, kF: float = 1.919, beta: float = 3.0, lam: float = 1.2, n_tau: int = 0):
        self._keep = {k: np.ascontiguousarray(meta[k], np.float64 if k == "loop_basis" else np.int32)
                      for k in ("leaf_type", "leaf_order", "tau_in", "tau_out", "loop_index", "loop_basis")}
        k = self._keep
        self.n_leaves = int(k["leaf_type"].shape[0])
        self.n_basis, self.n_loops = (int(x) for x in k["loop_basis"].shape) if k["loop_basis"].ndim == 2 else (0, 0)
        self.dim = int(dim)
        self.n_tau = int(n_tau) if n_tau else (int(max(k["tau_in"].max(initial=-1), k["tau_out"].max(initial=-1))) + 1)
        d = _capi.LeafGenDesc()
        d.n_leaves = self.n_leaves
        i32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        d.leaf_type, d.leaf_order = k["leaf_type"].ctypes.data_as(i32p), k["leaf_order"].ctypes.data_as(i32p)
        d.tau_in, d.tau_out = k["tau_in"].ctypes.data_as(i32p), k["tau_out"].ctypes.data_as(i32p)
        d.loop_index, d.loop_basis = k["loop_index"].ctypes.data_as(i32p), k["loop_basis"].ctypes.data_as(f64p)
        d.n_basis, d.n_loops, d.dim, d.n_tau = self.n_basis, self.n_loops, self.dim, self.n_tau
        d.kF, d.beta, d.lam = float(kF), float(beta), float(lam)
        self.kF, self.beta, self.lam = float(kF), float(beta), float(lam)
        self._g = C.c_void_p()
        _capi.check(_capi.lib().fdg_leafgen_create(C.byref(d), C.byref(self._g)))

    def close(self) -> None:
        if getattr(self, "_g", None):
            _capi.lib().fdg_leafgen_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def var_rows(self) -> int:
        """rows of the (K, T) input per sample: dim * n_loops + n_tau"""
        return self.dim * self.n_loops + self.n_tau

    def jit_prepare(self, wide: bool = False, index: int = -1):
        """The kernels specialised for this graph's leaves, built and assembled now (host only): counters [, PTX of one]."""
        return _capi.leafgen_jit_prepare(self._g, wide, index)

    def fill_device(self, K_ptr: int, T_ptr: int, ld_var: int, batch: int, leaf_ptr: int, ld_leaf: int, stream: int = 0) -> None:
        """K[(j * dim + c) * ld_var + b], T[t * ld_var + b] -> leaf[l * ld_leaf + b] (device pointers)."""
        _capi.check(_capi.lib().fdg_leafgen_fill(self._g, K_ptr, T_ptr, ld_var, batch, leaf_ptr, ld_leaf, stream))

    def accumulate_device(self, ev: Evaluator, K_ptr: int, T_ptr: int, ld_var: int, batch: int, acc_ptr: int, stream: int = 0) -> None:
        """acc[r] += sum over the batch; the leaf matrix only ever exists one sub-batch at a time."""
        _capi.check(_capi.lib().fdg_eval_generated_accumulate(ev._h, self._g, K_ptr, T_ptr, ld_var, batch, acc_ptr, stream))

    def accumulate_host(self, ev: Evaluator, K: np.ndarray, T: np.ndarray) -> np.ndarray:
        """K (dim * n_loops, B) and T (n_tau, B) host arrays, batch unit-stride -> the R per-root sums."""
        K = np.ascontiguousarray(K, np.float64)
        T = np.ascontiguousarray(T, np.float64)
        B = K.shape[1]
        assert K.shape[0] == self.dim * self.n_loops and T.shape == (self.n_tau, B)
        acc = np.zeros(max(ev.n_roots, 1))
        _capi.check(_capi.lib().fdg_eval_generated_host(ev._h, self._g, K.ctypes.data, T.ctypes.data, B, B, acc.ctypes.data))
        return acc[: ev.n_roots]
