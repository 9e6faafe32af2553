"""CPU BASELINE / CHECKER -- TEST INFRASTRUCTURE ONLY (same rules as oracle.py).

The reference's own CPU design point for this path is `Compilers.compile_C` + a C compiler (src/backend/static.jl:155-197,
:269-279; README.md:100-116): one straight-line C function per graph set, one call per sample.  This module produces
exactly that -- the text of `to_Cstr` (fdgraph_b200.emitters, same text as the reference) compiled with
`gcc -O2 -ffp-contract=off` (no FMA contraction: the bits of the Julia function) -- and runs it over a batch with
OpenMP over samples (oracle_run_emitted in fdg_oracle.c).  It is

  * an independent check of the oracle and of the GPU (three implementations, one set of bits), and
  * the CPU baseline bench.py times beside the GPU: ~20x faster per core than the array-walking oracle, i.e. the
    honest comparator ("what a user of the reference gets from compile_C on the same box").
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from typing import Optional

import numpy as np

from . import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE = os.path.join(HERE, "_build")


# The emitted text writes Power{N} as `pow(g, N)` (static.jl:37-38).  libm's pow is not what the Julia function
# computes (x^2 -> x*x, x^3 -> x*x*x by Base.literal_pow; N >= 4 -> base/math.jl pow_body; complex:
# Base.power_by_squaring), and for `complex double` it would not even keep the imaginary part.  The harness therefore
# compiles the UNCHANGED emitted text behind a prelude that gives `pow` the Julia meaning.
_PRELUDE_F64 = r"""#include <math.h>
static inline double fdg_pow(double x, long n) {
    if (n == 2) return x * x;
    if (n == 3) return x * x * x;
    double y = 1.0, xnlo = 0.0, ynlo = 0.0;
    while (n > 1) {
        if (n & 1) { double err = fma(y, xnlo, x * ynlo); double pr = x * y; ynlo = fma(x, y, -pr) + err; y = pr; }
        double err = x * 2 * xnlo; double pr = x * x; xnlo = fma(x, x, -pr) + err; x = pr; n >>= 1;
    }
    double err = fma(y, xnlo, x * ynlo);
    return (isfinite(x) && isfinite(err)) ? fma(x, y, err) : x * y;
}
#define pow(x, n) fdg_pow((x), (n))
"""
_PRELUDE_C128 = r"""#include <math.h>
#include <complex.h>
static inline complex double fdg_cpow(complex double x, int p) {
    if (p == 2) return x * x;
    if (p == 3) return x * x * x;
    int t = __builtin_ctz((unsigned)p) + 1;
    p >>= t;
    while (--t > 0) x = x * x;
    complex double y = x;
    while (p > 0) {
        t = __builtin_ctz((unsigned)p) + 1;
        p >>= t;
        while (--t >= 0) x = x * x;
        y = y * x;
    }
    return y;
}
#define pow(x, n) fdg_cpow((x), (n))
"""


def graphs_from_raw(raw):
    """Graph objects (ids preserved) behind the flattened arrays."""
    return raw.to_graphs()


def _key(raw, dtype: str) -> str:
    h = hashlib.sha1()
    for k in ("node_id", "node_op", "node_pow", "child_ptr", "child_node", "child_factor", "graphs", "root_id"):
        h.update(np.ascontiguousarray(getattr(raw, k)).tobytes())
    h.update(dtype.encode())
    return h.hexdigest()[:16]


def emitted_path(raw, dtype: str = "f64") -> str:
    return os.path.join(CACHE, f"emitted_{_key(raw, dtype)}.so")


def build_emitted(raw, dtype: str = "f64", timeout: Optional[float] = None) -> str:
    """to_Cstr text -> shared object (cached under oracle/_build by content hash).  Raises on failure / timeout."""
    import fdgraph_b200 as fd

    path = emitted_path(raw, dtype)
    if os.path.exists(path):
        return path
    os.makedirs(CACHE, exist_ok=True)
    graphs = graphs_from_raw(raw)
    text, _ = fd.Compilers.to_Cstr(graphs, root=[int(r) for r in raw.root_id],
                                   datatype="Float64" if dtype == "f64" else "ComplexF64")
    src = path[:-3] + ".c"
    with open(src, "w") as fh:
        fh.write(_PRELUDE_C128 if dtype != "f64" else _PRELUDE_F64)
        fh.write(text + "\n")
    # -fcx-limited-range: complex multiply is the four-multiply formula of Julia's *(::Complex, ::Complex), no C99 NaN recovery
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fcx-limited-range", "-fPIC", "-shared", "-o", path + ".tmp", src]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if proc.returncode != 0:
        raise RuntimeError("gcc failed on the emitted function:\n" + proc.stderr[-2000:])
    os.replace(path + ".tmp", path)
    return path


class Emitted:
    """The compiled `eval_graph(root, leafVal)` of one graph set, called once per sample."""

    def __init__(self, raw, dtype: str = "f64", timeout: Optional[float] = None):
        self.dtype = dtype
        self.orc = O.Oracle(raw)
        self.so = C.CDLL(build_emitted(raw, dtype, timeout))
        self.fn = C.cast(self.so.eval_graph, C.c_void_p)
        L = O.lib()
        L.oracle_run_emitted_c128.restype = C.c_int
        L.oracle_run_emitted_c128.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int]

    def eval(self, leaf: np.ndarray, root: Optional[np.ndarray] = None, nthreads: int = 1) -> np.ndarray:
        """leaf (B, L) sample-major -> root (B, R)."""
        npdt = np.float64 if self.dtype == "f64" else np.complex128
        leaf = np.ascontiguousarray(leaf, npdt)
        B = leaf.shape[0]
        if root is None:
            root = np.zeros((B, max(self.orc.n_roots, 1)), npdt)
        assert root.flags.c_contiguous and root.dtype == npdt
        run = O.lib().oracle_run_emitted if self.dtype == "f64" else O.lib().oracle_run_emitted_c128
        run(self.fn, leaf.ctypes.data, leaf.shape[1], root.ctypes.data, root.shape[1], B, int(nthreads))
        return root
