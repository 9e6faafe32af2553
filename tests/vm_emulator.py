"""Reference emulator of the packet ISA (feynmandiagram.jl_b200/csrc/fdg_isa.h), numpy, vectorised over
samples.  Test infrastructure: it lets the CPU-only suite check the *lowering* (statement order, fold order,
slot allocation, spills, prefetch hoisting, wait_group accounting) against the oracle without a GPU; the GPU
suite then only has to show that the CUDA kernel implements the same ISA.

The emulator is deliberately hostile about the asynchronous parts:
  * an LDL overwrites its slot immediately (worst case for write-after-read hazards), and
  * reading a slot whose cp.async group has not been covered by a WAIT raises (read-before-arrival).
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

NREG = 4
FIRST_REG_OP = 8
OP_END, OP_LDL, OP_WAIT, OP_SPILL, OP_FILL = 0, 1, 2, 3, 4
(R_MOV, R_MUL, R_ADD, R_MOVF, R_MULF, R_ADDF, R_SCALE, R_RADDF, R_RMULF, R_XADDF, R_XMULF, R_POW, R_ST,
 R_ROOT) = range(14)
NAMES = ["MOV", "MUL", "ADD", "MOVF", "MULF", "ADDF", "SCALE", "RADDF", "RMULF", "XADDF", "XMULF", "POW", "ST", "ROOT"]
SLOT_BITS = 12


def _fma(x, y, z):
    return float(Fraction(x) * Fraction(y) + Fraction(z))


def _pow_body(x: float, n: int) -> float:
    y, xnlo, ynlo = 1.0, 0.0, 0.0
    while n > 1:
        if n & 1:
            err = _fma(y, xnlo, x * ynlo)
            p = x * y
            ynlo = _fma(x, y, -p) + err
            y = p
        err = x * 2 * xnlo
        p = x * x
        xnlo = _fma(x, x, -p) + err
        x = p
        n >>= 1
    err = _fma(y, xnlo, x * ynlo)
    if np.isfinite(x) and np.isfinite(err):
        return _fma(x, y, err)
    return x * y


def _mk(re, im):
    out = np.empty(np.shape(re), np.complex128)
    out.real = re
    out.imag = im
    return out


def _cmul(a, b):
    # Julia *(z::Complex, w::Complex): no fused operations, no inf/nan fix-ups
    return _mk(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


def _cscale(a, f):
    return _mk(a.real * f, a.imag * f)


def _vpow(a: np.ndarray, n: int, cplx: bool):
    mul = _cmul if cplx else (lambda x, y: x * y)
    if n == 2:
        return mul(a, a)
    if n == 3:
        return mul(mul(a, a), a)
    if not cplx:
        return np.array([_pow_body(float(x), n) for x in a], dtype=np.float64)
    x = a
    t = (n & -n).bit_length()
    n >>= t
    t -= 1
    while t > 0:
        x = mul(x, x)
        t -= 1
    y = x
    while n > 0:
        t = (n & -n).bit_length()
        n >>= t
        while t > 0:
            x = mul(x, x)
            t -= 1
        y = mul(y, x)
    return y


def disassemble(words: np.ndarray):
    out = []
    for i, (w0, w1, w2, w3) in enumerate(np.asarray(words, np.uint32).reshape(-1, 4).tolist()):
        op, n, arg = w0 & 0xFF, (w0 >> 8) & 3, w0 >> 10
        f = np.array([w2, w3], np.uint32).view(np.float64)[0]
        if op == OP_END:
            out.append(f"{i:5d} END")
        elif op == OP_LDL:
            parts = [f"v[{w & 0xFFF}]<-leaf{w >> SLOT_BITS}" for w in (w1, w2, w3)[:n]]
            out.append(f"{i:5d} LDL " + ", ".join(parts))
        elif op == OP_WAIT:
            out.append(f"{i:5d} WAIT {arg}")
        elif op == OP_SPILL:
            out.append(f"{i:5d} SPILL scratch[{arg}] <- v[{w1}]")
        elif op == OP_FILL:
            out.append(f"{i:5d} FILL v[{w1}] <- scratch[{arg}]")
        else:
            base, d = divmod(op - FIRST_REG_OP, NREG)
            name = NAMES[base] if base < len(NAMES) else f"?{base}"
            if base in (R_MOV, R_MUL, R_ADD):
                out.append(f"{i:5d} {name}{n} a{d} " + " ".join(f"v[{w}]" for w in (w1, w2, w3)[:n]))
            elif base in (R_MOVF, R_MULF, R_ADDF, R_XADDF, R_XMULF):
                out.append(f"{i:5d} {name} a{d} v[{w1}] f={f!r}")
            elif base in (R_SCALE, R_RADDF, R_RMULF):
                out.append(f"{i:5d} {name} a{d} f={f!r}")
            else:
                out.append(f"{i:5d} {name} a{d} arg={arg}")
    return "\n".join(out)


def run(words: np.ndarray, leaf: np.ndarray, n_roots: int, strict: bool = True):
    """leaf: (L, B) float64 or complex128.  Returns (root (R, B), set-mask (R,), counters)."""
    w = np.asarray(words, np.uint32).reshape(-1, 4)
    cplx = leaf.dtype == np.complex128
    B = leaf.shape[1]
    mul = _cmul if cplx else (lambda x, y: x * y)
    scale = _cscale if cplx else (lambda x, f: x * f)
    root = np.zeros((n_roots, B), leaf.dtype)
    root_set = np.zeros(n_roots, bool)
    slots = {}
    slot_group = {}
    scratch = {}
    acc = [np.zeros(B, leaf.dtype) for _ in range(NREG)]
    committed = completed = 0
    cnt = {"packets": 0, "ldl": 0, "wait": 0, "slot_reads": 0, "max_slot": -1}

    def rd(s):
        cnt["slot_reads"] += 1
        if strict:
            g = slot_group.get(s, -1)
            if g >= completed:
                raise AssertionError(f"slot {s} read before its cp.async group {g} was waited for (completed={completed})")
            if s not in slots:
                raise AssertionError(f"slot {s} read before being written")
        return slots[s]

    ended = False
    np.seterr(all="ignore")  # overflow / nan are legitimate values here; they must match bit for bit too
    for pc in range(w.shape[0]):
        w0, w1, w2, w3 = (int(x) for x in w[pc])
        op, n, arg = w0 & 0xFF, (w0 >> 8) & 3, w0 >> 10
        f = float(np.array([w2, w3], np.uint32).view(np.float64)[0])
        cnt["packets"] += 1
        if op == OP_END:
            ended = True
            break
        if op == OP_LDL:
            assert 1 <= n <= 3
            for ww in (w1, w2, w3)[:n]:
                s, l = ww & ((1 << SLOT_BITS) - 1), ww >> SLOT_BITS
                slots[s] = leaf[l].copy()
                slot_group[s] = committed
                cnt["ldl"] += 1
                cnt["max_slot"] = max(cnt["max_slot"], s)
            committed += 1
            continue
        if op == OP_WAIT:
            assert arg <= 7
            completed = max(completed, committed - arg)
            cnt["wait"] += 1
            continue
        if op == OP_SPILL:
            scratch[arg] = rd(w1).copy()
            continue
        if op == OP_FILL:
            slots[w1] = scratch[arg].copy()
            slot_group[w1] = -1
            cnt["max_slot"] = max(cnt["max_slot"], w1)
            continue
        assert op >= FIRST_REG_OP, f"bad opcode {op}"
        base, d = divmod(op - FIRST_REG_OP, NREG)
        A = acc[d]
        if base == R_MOV:
            v = rd(w1)
            if n >= 2:
                v = mul(v, rd(w2))
            if n >= 3:
                v = mul(v, rd(w3))
            acc[d] = v.copy()
        elif base == R_MUL:
            for ww in (w1, w2, w3)[:n]:
                A = mul(A, rd(ww))
            acc[d] = A
        elif base == R_ADD:
            for ww in (w1, w2, w3)[:n]:
                A = A + rd(ww)
            acc[d] = A
        elif base == R_MOVF:
            acc[d] = scale(rd(w1), f)
        elif base == R_MULF:
            acc[d] = scale(mul(A, rd(w1)), f)
        elif base == R_ADDF:
            acc[d] = A + scale(rd(w1), f)
        elif base == R_SCALE:
            acc[d] = scale(A, f)
        elif base == R_RADDF:
            assert d >= 1
            acc[d - 1] = acc[d - 1] + scale(A, f)
        elif base == R_RMULF:
            assert d >= 1
            acc[d - 1] = scale(mul(acc[d - 1], A), f)
        elif base == R_XADDF:
            acc[d] = rd(w1) + scale(A, f)
        elif base == R_XMULF:
            acc[d] = scale(mul(rd(w1), A), f)
        elif base == R_POW:
            acc[d] = _vpow(A, arg, cplx)
        elif base == R_ST:
            slots[arg] = A.copy()
            slot_group[arg] = -1
            cnt["max_slot"] = max(cnt["max_slot"], arg)
        elif base == R_ROOT:
            root[arg] = A
            root_set[arg] = True
        else:
            raise AssertionError(f"bad register op {base}")
    assert ended, "program has no END packet"
    return root, root_set, cnt
