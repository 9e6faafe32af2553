"""ctypes binding of libfdgraph.so -- the same stub a Julia `ccall` wrapper makes."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _build

FDG_OK = 0
ERR_NAMES = {1: "BAD_ARG", 2: "BAD_GRAPH", 3: "UNSUPPORTED", 4: "CUDA", 5: "NCCL", 6: "NO_DEVICE", 7: "CAPACITY"}
FDG_F64, FDG_C128 = 0, 1


class FdgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfdgraph: {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class GraphDesc(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int64), ("n_edges", C.c_int64),
        ("node_id", C.POINTER(C.c_int64)), ("node_op", C.POINTER(C.c_int32)), ("node_pow", C.POINTER(C.c_int32)),
        ("child_ptr", C.POINTER(C.c_int64)), ("child_node", C.POINTER(C.c_int32)),
        ("child_factor", C.POINTER(C.c_double)),
        ("n_graphs", C.c_int64), ("graphs", C.POINTER(C.c_int32)),
        ("n_roots", C.c_int64), ("root_id", C.POINTER(C.c_int64)),
    ]


class Options(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("max_slots", C.c_int32), ("prefetch", C.c_int32), ("schedule", C.c_int32),
                ("backend", C.c_int32), ("jit_segment", C.c_int32), ("cse", C.c_int32), ("fma", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n_leaves", "n_inner", "n_roots", "n_operands", "n_packets", "n_slots", "n_scratch", "leaf_loads",
        "flops_add", "flops_mul", "bytes_in", "bytes_out", "max_depth", "cse_removed")] + [("reserved", C.c_int64 * 2)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "reserved"}


class LeafGenDesc(C.Structure):
    _fields_ = [
        ("n_leaves", C.c_int64), ("leaf_type", C.POINTER(C.c_int32)), ("leaf_order", C.POINTER(C.c_int32)),
        ("tau_in", C.POINTER(C.c_int32)), ("tau_out", C.POINTER(C.c_int32)), ("loop_index", C.POINTER(C.c_int32)),
        ("n_basis", C.c_int64), ("n_loops", C.c_int64), ("dim", C.c_int64), ("n_tau", C.c_int64),
        ("loop_basis", C.POINTER(C.c_double)), ("kF", C.c_double), ("beta", C.c_double), ("lam", C.c_double),
    ]


# every symbol include/fdgraph.h declares; tests check that the library exports each one
EXPORTS = [
    "fdg_abi_version", "fdg_last_error", "fdg_compile", "fdg_destroy", "fdg_stats", "fdg_leafmap", "fdg_last_root",
    "fdg_program_words", "fdg_eval", "fdg_eval_accumulate", "fdg_eval_host", "fdg_set_launch", "fdg_launch_count",
    "fdg_comm_unique_id", "fdg_comm_init", "fdg_comm_destroy", "fdg_allreduce", "fdg_jit_prepare", "fdg_jit_ptx",
    "fdg_jit_info", "fdg_leafgen_create", "fdg_leafgen_destroy", "fdg_leafgen_fill", "fdg_eval_generated_accumulate",
    "fdg_eval_generated_host", "fdg_graph_write", "fdg_compile_file", "fdg_pipeline_prepare", "fdg_pipeline_stats",
    "fdg_probe_fp64", "fdg_leafgen_jit_prepare",
]
BACKEND_AUTO, BACKEND_VM, BACKEND_JIT = 0, 1, 2

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Loads libfdgraph.so (building it in-tree if a compiler is there).  Fails loudly otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path):
        path = _build.build()
    L = C.CDLL(path)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    L.fdg_abi_version.restype = C.c_int
    L.fdg_last_error.restype = C.c_char_p
    L.fdg_compile.argtypes = [C.POINTER(GraphDesc), C.POINTER(Options), C.POINTER(vp)]
    L.fdg_destroy.argtypes = [vp]
    L.fdg_stats.argtypes = [vp, C.POINTER(Stats)]
    L.fdg_leafmap.argtypes = [vp, C.POINTER(i32)]
    L.fdg_last_root.argtypes = [vp, C.POINTER(i32)]
    L.fdg_program_words.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(i64)]
    L.fdg_eval.argtypes = [vp, vp, i64, vp, i64, i64, vp]
    L.fdg_eval_accumulate.argtypes = [vp, vp, i64, i64, vp, vp]
    L.fdg_eval_host.argtypes = [vp, vp, i64, vp, i64, i64]
    L.fdg_set_launch.argtypes = [vp, i32, i32, i32]
    L.fdg_launch_count.argtypes = [vp, C.POINTER(i64)]
    L.fdg_comm_unique_id.argtypes = [vp]
    L.fdg_comm_init.argtypes = [C.POINTER(vp), i32, i32, vp]
    L.fdg_comm_destroy.argtypes = [vp]
    L.fdg_allreduce.argtypes = [vp, vp, i64, vp]
    L.fdg_jit_prepare.argtypes = [vp, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    L.fdg_jit_ptx.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
    L.fdg_jit_info.argtypes = [vp, i32, i32, C.POINTER(i64), i32]
    L.fdg_pipeline_prepare.argtypes = [vp, i32, i32, i32, C.POINTER(i64), i32]
    L.fdg_pipeline_stats.argtypes = [vp, vp, C.POINTER(i64), i32]
    L.fdg_probe_fp64.argtypes = [i32, i64, vp, vp, C.POINTER(i64)]
    L.fdg_graph_write.argtypes = [C.POINTER(GraphDesc), C.c_char_p]
    L.fdg_compile_file.argtypes = [C.c_char_p, C.POINTER(Options), C.POINTER(vp)]
    L.fdg_leafgen_create.argtypes = [C.POINTER(LeafGenDesc), C.POINTER(vp)]
    L.fdg_leafgen_destroy.argtypes = [vp]
    L.fdg_leafgen_jit_prepare.argtypes = [vp, i32, i32, C.POINTER(i64), i32, C.POINTER(C.c_char_p)]
    L.fdg_leafgen_fill.argtypes = [vp, vp, vp, i64, i64, vp, i64, vp]
    L.fdg_eval_generated_accumulate.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp]
    L.fdg_eval_generated_host.argtypes = [vp, vp, vp, vp, i64, i64, vp]
    for name in EXPORTS:
        if name not in ("fdg_abi_version", "fdg_last_error"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != FDG_OK:
        raise FdgError(rc, lib().fdg_last_error().decode("utf-8", "replace"))


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _desc(raw) -> GraphDesc:
    raw.validate_dtypes()
    d = GraphDesc()
    d.n_nodes, d.n_edges = raw.n_nodes, raw.n_edges
    d.node_id = _ptr(raw.node_id, C.c_int64)
    d.node_op = _ptr(raw.node_op, C.c_int32)
    d.node_pow = _ptr(raw.node_pow, C.c_int32)
    d.child_ptr = _ptr(raw.child_ptr, C.c_int64)
    d.child_node = _ptr(raw.child_node, C.c_int32)
    d.child_factor = _ptr(raw.child_factor, C.c_double)
    d.n_graphs, d.graphs = int(raw.graphs.shape[0]), _ptr(raw.graphs, C.c_int32)
    d.n_roots, d.root_id = int(raw.root_id.shape[0]), _ptr(raw.root_id, C.c_int64)
    return d


def graph_write(raw, path: str) -> None:
    """fdg_graph_write: the FDGRAPH file of a RawGraph (include/fdgraph.h)."""
    d = _desc(raw)
    check(lib().fdg_graph_write(C.byref(d), os.fsencode(path)))


def _cse_mode(cse) -> int:
    """None -> 0 (automatic), True -> 1 (always), False -> -1 (never)"""
    return 0 if cse is None else (1 if cse else -1)


def compile_file(path: str, dtype: int = FDG_F64, backend: int = 0, jit_segment: int = 0, cse=None) -> C.c_void_p:
    o = Options()
    o.dtype, o.backend, o.jit_segment, o.cse = int(dtype), int(backend), int(jit_segment), _cse_mode(cse)
    h = C.c_void_p()
    check(lib().fdg_compile_file(os.fsencode(path), C.byref(o), C.byref(h)))
    return h


def compile_raw(raw, dtype: int = FDG_F64, max_slots: int = 0, prefetch: int = 0, schedule: int = 0, backend: int = 0,
                jit_segment: int = 0, cse=None, fma: bool = False) -> C.c_void_p:
    """fdg_compile on a RawGraph; returns the opaque handle."""
    L = lib()
    raw.validate_dtypes()
    d = GraphDesc()
    d.n_nodes, d.n_edges = raw.n_nodes, raw.n_edges
    d.node_id = _ptr(raw.node_id, C.c_int64)
    d.node_op = _ptr(raw.node_op, C.c_int32)
    d.node_pow = _ptr(raw.node_pow, C.c_int32)
    d.child_ptr = _ptr(raw.child_ptr, C.c_int64)
    d.child_node = _ptr(raw.child_node, C.c_int32)
    d.child_factor = _ptr(raw.child_factor, C.c_double)
    d.n_graphs, d.graphs = int(raw.graphs.shape[0]), _ptr(raw.graphs, C.c_int32)
    d.n_roots, d.root_id = int(raw.root_id.shape[0]), _ptr(raw.root_id, C.c_int64)
    o = Options()
    o.dtype, o.max_slots, o.prefetch, o.schedule = int(dtype), int(max_slots), int(prefetch), int(schedule)
    o.backend, o.jit_segment, o.cse, o.fma = int(backend), int(jit_segment), _cse_mode(cse), int(bool(fma))
    h = C.c_void_p()
    check(L.fdg_compile(C.byref(d), C.byref(o), C.byref(h)))
    return h


def stats(h) -> dict:
    s = Stats()
    check(lib().fdg_stats(h, C.byref(s)))
    return s.as_dict()


def leafmap(h, n_leaves: int) -> np.ndarray:
    out = np.empty(max(n_leaves, 1), np.int32)
    check(lib().fdg_leafmap(h, _ptr(out, C.c_int32)))
    return out[:n_leaves]


def last_root(h) -> int:
    v = C.c_int32()
    check(lib().fdg_last_root(h, C.byref(v)))
    return int(v.value)


def program_words(h) -> np.ndarray:
    p = C.POINTER(C.c_uint32)()
    n = C.c_int64()
    check(lib().fdg_program_words(h, C.byref(p), C.byref(n)))
    return np.ctypeslib.as_array(p, shape=(int(n.value),)).copy()


def launch_count(h) -> int:
    v = C.c_int64()
    check(lib().fdg_launch_count(h, C.byref(v)))
    return int(v.value)


def jit_prepare(h, samples_per_thread: int = 2, accumulate: bool = False) -> dict:
    nk, nc, nb = C.c_int32(), C.c_int32(), C.c_int64()
    check(lib().fdg_jit_prepare(h, samples_per_thread, int(accumulate), C.byref(nk), C.byref(nc), C.byref(nb)))
    info = jit_info(h, samples_per_thread, accumulate)
    info.update(kernels=int(nk.value), cross_rows=int(nc.value), cubin_bytes=int(nb.value))
    return info


def jit_info(h, samples_per_thread: int = 0, accumulate: bool = False) -> dict:
    """Counters of a prepared variant; samples_per_thread = 0: of the variant the last launch ran."""
    out = (C.c_int64 * 15)()
    check(lib().fdg_jit_info(h, samples_per_thread, int(accumulate), out, 15))
    return {"kernels": int(out[0]), "cross_rows": int(out[1]), "cross_values": int(out[2]),
            "leaf_loads": int(out[3]), "cross_loads": int(out[4]), "cross_stores": int(out[5]), "operations": int(out[6]),
            "grid_stride": bool(out[7]), "max_code_bytes": int(out[8]), "cse": bool(out[9]), "fp64_instr": int(out[10]),
            "model_ns": out[11] / 1000.0, "bulk": bool(out[12]), "bulk_smem": int(out[13]), "refetch_loads": int(out[14])}


def pipeline_prepare(h, accumulate: bool = True, n_sm: int = 148) -> dict:
    """Builds the pipeline kernel for a device with n_sm SMs (host only) and returns its plan."""
    out = (C.c_int64 * 12)()
    check(lib().fdg_pipeline_prepare(h, int(accumulate), n_sm, 0, out, 12))
    keys = ("stages", "cross_rows", "cross_values", "leaf_loads", "cross_loads", "cross_stores", "operations", "max_code_bytes",
            "linked_bytes", "ring_bytes", "passes", "boundary_rows")
    info = {k: int(out[i]) for i, k in enumerate(keys)}
    n = info["stages"]
    buf = (C.c_int64 * n)()
    check(lib().fdg_pipeline_prepare(h, int(accumulate), n_sm, 1, buf, n))
    info["stage_blocks"] = [int(x) for x in buf]
    check(lib().fdg_pipeline_prepare(h, int(accumulate), n_sm, 2, buf, n))
    info["stage_cost"] = [int(x) for x in buf]
    buf2 = (C.c_int64 * (n + 1))()
    check(lib().fdg_pipeline_prepare(h, int(accumulate), n_sm, 3, buf2, n + 1))
    info["stage_start"] = [int(x) for x in buf2]
    return info


def pipeline_stats(h, stream: int, n_stages: int) -> dict:
    out = (C.c_int64 * (1 + 2 * n_stages))()
    check(lib().fdg_pipeline_stats(h, stream, out, 1 + 2 * n_stages))
    return {"stalled": bool(out[0]), "busy": [int(out[1 + 2 * k]) for k in range(n_stages)],
            "waiting": [int(out[2 + 2 * k]) for k in range(n_stages)]}


def jit_ptx(h, samples_per_thread: int, accumulate: bool, index: int):
    p, log = C.c_char_p(), C.c_char_p()
    check(lib().fdg_jit_ptx(h, samples_per_thread, int(accumulate), index, C.byref(p), C.byref(log)))
    return p.value.decode(), (log.value or b"").decode()


def leafgen_jit_prepare(g, wide: bool = False, index: int = -1):
    """Builds + assembles the specialised leaf-generation kernels (host only); counters and, for index >= 0, the PTX."""
    out = (C.c_int64 * 5)()
    ptx = C.c_char_p()
    check(lib().fdg_leafgen_jit_prepare(g, int(wide), max(index, 0), out, 5, C.byref(ptx) if index >= 0 else None))
    info = {"kernels": int(out[0]), "code_bytes": int(out[1]), "instructions": int(out[2]), "leaves_covered": int(out[3]),
            "max_code_bytes": int(out[4])}
    return (info, ptx.value.decode()) if index >= 0 else info
