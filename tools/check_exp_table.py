"""Checks the table-based exp of the leaf generator (csrc/fdg_capi.cu, lg_exp_tab) against 60-digit arithmetic: the same
operations (fused multiply-adds emulated exactly), 20 000 arguments in [-746, 0].  Prints the constants the kernel uses."""
import math
import random
import struct

import mpmath as mp

mp.mp.prec = 200


def fma(a, b, c):
    return float(mp.mpf(a) * mp.mpf(b) + mp.mpf(c))


def trunc_bits(v, bits):
    m, e = mp.frexp(v)
    return float(mp.ldexp(mp.floor(mp.ldexp(m, bits)), e - bits))


LN2_32 = mp.log(2) / 32
HI = trunc_bits(LN2_32, 30)
LO = float(LN2_32 - mp.mpf(HI))
INV = float(32 / mp.log(2))
TAB = [float(mp.power(2, mp.mpf(j) / 32)) for j in range(32)]
MAGIC = 6755399441055744.0
C = [1.0, 0.5, 1 / 6, 1 / 24, 1 / 120, 1 / 720]


def exp_tab(x):
    t = fma(x, INV, MAGIC)
    n = struct.unpack("<i", struct.pack("<d", t)[:4])[0]
    nd = t - MAGIC
    r = fma(nd, -HI, x)
    r = fma(nd, -LO, r)
    q = C[5]
    for c in (C[4], C[3], C[2], C[1], C[0]):
        q = fma(q, r, c)
    q = q * r
    j, m = n & 31, n >> 5
    return math.ldexp(fma(TAB[j], q, TAB[j]), m)


if __name__ == "__main__":
    print("ln2/32 hi", repr(HI), "lo", repr(LO), "32/ln2", repr(INV))
    print("table", ", ".join(v.hex() for v in TAB[:4]), "...")
    random.seed(1)
    worst = 0.0
    for _ in range(20000):
        x = -random.random() * random.choice([1e-3, 1, 10, 100, 700])
        want = mp.exp(mp.mpf(x))
        if want > 1e-300:
            worst = max(worst, float(abs((mp.mpf(exp_tab(x)) - want) / want)))
    print(f"worst relative error {worst:.3e} = {worst / 2 ** -52:.2f} ulp")
    assert worst < 2 ** -52
