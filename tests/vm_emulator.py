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

CHUNK = 32
TERM_MAX = 11
(OP_END, OP_NOP, OP_LDL, OP_SPILL, OP_FILL, OP_TERM, OP_MOV, OP_MUL, OP_ADD, OP_MULF, OP_SCALE, OP_RADDF, OP_RMULF,
 OP_XADDF, OP_XMULF, OP_POW, OP_ST, OP_ROOT) = range(18)
NAMES = ["END", "NOP", "LDL", "SPILL", "FILL", "TERM", "MOV", "MUL", "ADD", "MULF", "SCALE", "RADDF", "RMULF", "XADDF",
         "XMULF", "POW", "ST", "ROOT"]
SLOT_BITS = 12


def _hdr(w0):
    return dict(op=w0 & 63, k=(w0 >> 6) & 15, wait=(w0 >> 10) & 7, push=(w0 >> 13) & 1, first=(w0 >> 14) & 1, slot0=w0 >> 20)


def _term_records(w, pc, k, n):
    """records of a TERM block whose header is packet pc -> ([(slots, f)], packets consumed after the header)"""
    rec = 2 if k > 4 else 1
    out = []
    for t in range(n):
        r = w[pc + 1 + t * rec]
        e = w[pc + 2 + t * rec] if rec == 2 else [0, 0, 0, 0]
        words = [r[0], r[1], e[0], e[1], e[2], e[3]]
        slots = [(int(words[q >> 1]) >> (16 * (q & 1))) & 0xFFFF for q in range(k)]
        f = float(np.array([r[2], r[3]], np.uint32).view(np.float64)[0])
        out.append((slots, f))
    return out, n * rec


def _ldl_words(w, pc, k):
    ws = list(w[pc][1:4])
    if k > 3:
        ws += list(w[pc + 1])
    return [int(x) for x in ws[:k]], (1 if k > 3 else 0)


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
    w = np.asarray(words, np.uint32).reshape(-1, 4).tolist()
    out = []
    pc = 0
    while pc < len(w):
        w0, w1, w2, w3 = w[pc]
        h = _hdr(w0)
        op, k = h["op"], h["k"]
        f = np.array([w2, w3], np.uint32).view(np.float64)[0]
        pre = f"{pc:5d} " + (f"[wait {h['wait'] - 1}] " if h["wait"] else "") + ("[push] " if h["push"] else "")
        name = NAMES[op] if op < len(NAMES) else f"?{op}"
        if op == OP_LDL:
            ws, extra = _ldl_words(w, pc, k)
            out.append(pre + "LDL " + ", ".join(f"v[{x & 0xFFF}]<-leaf{x >> SLOT_BITS}" for x in ws))
            pc += extra
        elif op in (OP_SPILL, OP_FILL):
            out.append(pre + f"{name} v[{w1}] scratch[{w2}]")
        elif op == OP_TERM:
            recs, extra = _term_records(w, pc, k, w1)
            for t, (sl, ff) in enumerate(recs):
                lead = pre if t == 0 else " " * len(pre)
                out.append(lead + ("A  = " if h["first"] and t == 0 else "A += ") + " * ".join(f"v[{x}]" for x in sl) + f" * {ff!r}")
            pc += extra
        elif op in (OP_MOV, OP_MUL, OP_ADD):
            out.append(pre + f"{name}{k} " + " ".join(f"v[{x}]" for x in (w1, w2, w3)[:k]))
        elif op in (OP_MULF, OP_XADDF, OP_XMULF):
            out.append(pre + f"{name} v[{w1}] f={float(f)!r}")
        elif op in (OP_SCALE, OP_RADDF, OP_RMULF):
            out.append(pre + f"{name} f={float(f)!r}")
        elif op in (OP_POW, OP_ST, OP_ROOT):
            out.append(pre + f"{name} {w1}")
        else:
            out.append(pre + name)
        if op == OP_END:
            break
        pc += 1
    return "\n".join(out)


def run(words: np.ndarray, leaf: np.ndarray, n_roots: int, strict: bool = True):
    """leaf: (L, B) float64 or complex128.  Returns (root (R, B), set-mask (R,), counters)."""
    w = np.asarray(words, np.uint32).reshape(-1, 4).tolist()
    cplx = leaf.dtype == np.complex128
    B = leaf.shape[1]
    mul = _cmul if cplx else (lambda x, y: x * y)
    scale = _cscale if cplx else (lambda x, f: x * f)
    root = np.zeros((n_roots, B), leaf.dtype)
    root_set = np.zeros(n_roots, bool)
    slots = {}
    slot_group = {}
    scratch = {}
    zero = np.zeros(B, leaf.dtype)
    A, R1, R2, R3 = zero, zero, zero, zero
    depth = 0
    committed = completed = 0
    cnt = {"packets": 0, "ldl": 0, "wait": 0, "slot_reads": 0, "max_slot": -1, "max_depth": 0}

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
    pc = 0
    while pc < len(w):
        w0, w1, w2, w3 = w[pc]
        h = _hdr(w0)
        op, k = h["op"], h["k"]
        f = float(np.array([w2, w3], np.uint32).view(np.float64)[0])
        cnt["packets"] += 1
        if h["wait"]:
            assert h["wait"] - 1 <= 6
            completed = max(completed, committed - (h["wait"] - 1))
            cnt["wait"] += 1
        if h["push"]:
            assert op == OP_MOV or (op == OP_TERM and h["first"]), "push on a packet that does not start a fold"
            R3, R2, R1 = R2, R1, A
            depth += 1
            assert depth <= 3, "accumulator stack overflow"
            cnt["max_depth"] = max(cnt["max_depth"], depth)
        if op == OP_END:
            ended = True
            break
        elif op == OP_NOP:
            pass
        elif op == OP_LDL:
            assert 1 <= k <= 7
            ws, extra = _ldl_words(w, pc, k)
            assert pc // CHUNK == (pc + extra) // CHUNK, "LDL extension packet crosses a chunk boundary"
            pc += extra
            cnt["packets"] += extra
            for ww in ws:
                s, l = ww & ((1 << SLOT_BITS) - 1), ww >> SLOT_BITS
                slots[s] = leaf[l].copy()
                slot_group[s] = committed
                cnt["ldl"] += 1
                cnt["max_slot"] = max(cnt["max_slot"], s)
            committed += 1
        elif op == OP_SPILL:
            scratch[w2] = rd(w1).copy()
        elif op == OP_FILL:
            slots[w1] = scratch[w2].copy()
            slot_group[w1] = -1
            cnt["max_slot"] = max(cnt["max_slot"], w1)
        elif op == OP_TERM:
            assert 1 <= k <= TERM_MAX and w1 >= 1
            recs, extra = _term_records(w, pc, k, w1)
            assert pc // CHUNK == (pc + extra) // CHUNK, "TERM block crosses a chunk boundary"
            pc += extra
            cnt["packets"] += extra
            cnt["terms"] = cnt.get("terms", 0) + len(recs)
            for t_i, (sl, ff) in enumerate(recs):
                t = rd(sl[0])
                for x in sl[1:]:
                    t = mul(t, rd(x))
                t = scale(t, ff)
                A = t.copy() if (h["first"] and t_i == 0) else A + t
        elif op == OP_MOV:
            v = rd(w1)
            if k >= 2:
                v = mul(v, rd(w2))
            if k >= 3:
                v = mul(v, rd(w3))
            A = v.copy()
        elif op == OP_MUL:
            for ww in (w1, w2, w3)[:k]:
                A = mul(A, rd(ww))
        elif op == OP_ADD:
            for ww in (w1, w2, w3)[:k]:
                A = A + rd(ww)
        elif op == OP_MULF:
            A = scale(mul(A, rd(w1)), f)
        elif op == OP_SCALE:
            A = scale(A, f)
        elif op == OP_RADDF:
            assert depth >= 1, "pop from an empty accumulator stack"
            A = R1 + scale(A, f)
            R1, R2 = R2, R3
            depth -= 1
        elif op == OP_RMULF:
            assert depth >= 1, "pop from an empty accumulator stack"
            A = scale(mul(R1, A), f)
            R1, R2 = R2, R3
            depth -= 1
        elif op == OP_XADDF:
            A = rd(w1) + scale(A, f)
        elif op == OP_XMULF:
            A = scale(mul(rd(w1), A), f)
        elif op == OP_POW:
            A = _vpow(A, w1, cplx)
        elif op == OP_ST:
            slots[w1] = A.copy()
            slot_group[w1] = -1
            cnt["max_slot"] = max(cnt["max_slot"], w1)
        elif op == OP_ROOT:
            root[w1] = A
            root_set[w1] = True
        else:
            raise AssertionError(f"bad opcode {op}")
        pc += 1
    assert ended, "program has no END packet"
    return root, root_set, cnt
