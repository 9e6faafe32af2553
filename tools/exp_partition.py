"""Development experiment (CPU only): how far can the leaf re-reads of a kernel plan drop if the units of the emitted
function (a multi-use value with the single-use values folded into it, long folds chopped into chunks) may be assigned
to kernels freely -- subject to the dependences -- instead of being cut out of ONE linear order?

    FDG_JIT_DUMP_IR=/tmp/ir.bin python -c "...jit_prepare..."   # the fold-order IR
    python tools/exp_partition.py /tmp/ir.bin --kernels 19
"""
import argparse
import random
import sys
from collections import defaultdict

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ir")
    ap.add_argument("--kernels", type=int, default=19)
    ap.add_argument("--chunk", type=int, default=120)
    ap.add_argument("--slack", type=float, default=1.15)
    ap.add_argument("--iters", type=int, default=400000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--temp", type=float, default=0.3)
    ap.add_argument("--free", action="store_true")
    a = ap.parse_args()
    ir = np.fromfile(a.ir, dtype=np.int32).reshape(-1, 4)
    kind, A, B, _ = ir.T
    N = len(ir)
    binary = kind <= 1
    cost1 = np.where(kind == 5, 0, 1)
    # uses and the (last) consumer of every value
    uses = np.zeros(N, np.int64)
    cons = np.full(N, -1, np.int64)
    for i in range(N):
        if A[i] >= 0:
            uses[A[i]] += 1
            cons[A[i]] = i
        if binary[i] and B[i] >= 0:
            uses[B[i]] += 1
            cons[B[i]] = i
    # units: head = value with uses != 1 or a ROOT op; a single-use value belongs to the unit of its consumer
    unit = np.arange(N)
    for i in range(N - 1, -1, -1):
        if kind[i] != 4 and uses[i] == 1:
            unit[i] = unit[cons[i]]
    # chop big units into chunks of <= chunk ops (in IR order)
    members = defaultdict(list)
    for i in range(N):
        members[unit[i]].append(i)
    chunk_of = np.zeros(N, np.int64)
    chunks = []
    for h, ops in members.items():
        for k in range(0, len(ops), a.chunk):
            for i in ops[k:k + a.chunk]:
                chunk_of[i] = len(chunks)
            chunks.append(ops[k:k + a.chunk])
    C = len(chunks)
    ccost = np.array([int(cost1[c].sum()) for c in chunks])
    print(f"ops={N} units={len(members)} chunks={C} biggest unit={max(map(len, members.values()))}")
    # leaves of a chunk, predecessor chunks (values read from another chunk)
    cleaves = []
    preds = [set() for _ in range(C)]
    succs = [set() for _ in range(C)]
    outvals = [set() for _ in range(C)]  # values of chunk c read by other chunks
    readers = defaultdict(set)            # value -> chunks reading it (other than its own)
    for c, ops in enumerate(chunks):
        ls = set()
        for i in ops:
            for x in ((A[i],) if not binary[i] else (A[i], B[i])):
                if x < 0:
                    ls.add(-x - 1)
                elif chunk_of[x] != c:
                    preds[c].add(chunk_of[x])
                    succs[chunk_of[x]].add(c)
                    outvals[chunk_of[x]].add(x)
                    readers[x].add(c)
        cleaves.append(ls)
    L = 1 + max(max(s) for s in cleaves if s)
    K = a.kernels
    total = int(ccost.sum())
    cap = int(total / K * a.slack)
    # initial assignment: the linear plan (equal-cost cuts of the IR order), chunk placed by its last op
    cum = np.cumsum(cost1)
    seg = np.minimum((cum[[c[-1] for c in chunks]] - 1) * K // total, K - 1).astype(np.int64)
    load = np.zeros(K, np.int64)
    for c in range(C):
        load[seg[c]] += ccost[c]
    cnt = np.zeros((K, L), np.int64)
    for c in range(C):
        for l in cleaves[c]:
            cnt[seg[c], l] += 1

    def cross_cost_of_value(x):
        s0 = seg[chunk_of[x]]
        later = {seg[r] for r in readers[x]} - {s0}
        return (1 + len(later)) if later else 0

    def total_cost():
        leaf = int((cnt > 0).sum())
        cross = sum(cross_cost_of_value(x) for x in readers)
        return leaf, cross

    lf, cr = total_cost()
    print(f"linear plan, K={K}: leaf loads {lf}, cross rows moved {cr}, total {lf + cr}; max load {load.max()} cap {cap}")
    rng = random.Random(a.seed)
    cur = lf + cr
    best = cur
    T0 = a.temp
    for it in range(a.iters):
        T = T0 * (1 - it / a.iters) + 0.01
        c = rng.randrange(C)
        lo = 0 if a.free else max((seg[p] for p in preds[c]), default=0)
        hi = K - 1 if a.free else min((seg[s] for s in succs[c]), default=K - 1)
        if lo >= hi and not (lo == hi and seg[c] != lo):
            continue
        s_old = seg[c]
        s_new = rng.randint(lo, hi)
        if s_new == s_old or load[s_new] + ccost[c] > cap:
            continue
        # delta
        d = 0
        for l in cleaves[c]:
            if cnt[s_old, l] == 1:
                d -= 1
            if cnt[s_new, l] == 0:
                d += 1
        vals = set(outvals[c])
        for i in chunks[c]:
            for x in ((A[i],) if not binary[i] else (A[i], B[i])):
                if x >= 0 and chunk_of[x] != c:
                    vals.add(x)
        before = sum(cross_cost_of_value(x) for x in vals)
        seg[c] = s_new
        after = sum(cross_cost_of_value(x) for x in vals)
        d += 0 if a.free else after - before
        if d <= 0 or rng.random() < np.exp(-d / T):
            for l in cleaves[c]:
                cnt[s_old, l] -= 1
                cnt[s_new, l] += 1
            load[s_old] -= ccost[c]
            load[s_new] += ccost[c]
            cur += d
            best = min(best, cur)
        else:
            seg[c] = s_old
        if it % 50000 == 0:
            print(f"  it {it}: cost {cur} (best {best}) T={T:.2f}", flush=True)
    lf, cr = total_cost()
    print(f"after search: leaf loads {lf}, cross rows moved {cr}, total {lf + cr}; loads {load.min()}..{load.max()}")


if __name__ == "__main__":
    sys.exit(main())
