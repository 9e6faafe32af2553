"""Kernel-only sweep of launch / lowering parameters on one GPU (development tool, not a bench line)."""
import argparse
import itertools
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fdgraph_b200 as fd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="gv_ver4_o4")
    ap.add_argument("--gb", type=float, default=8.0)
    ap.add_argument("--slots", default="32,48,64,96")
    ap.add_argument("--threads", default="64,128,256")
    ap.add_argument("--spt", default="2,4")
    ap.add_argument("--prefetch", default="-1,8,24,64")
    ap.add_argument("--mode", default="acc")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--backend", type=int, default=1, help="1 = VM, 2 = JIT")
    ap.add_argument("--jit-segment", default="0")
    a = ap.parse_args()
    raw = fd.RawGraph.load(os.path.join(ROOT, "workloads", a.workload + ".npz"))
    base = fd.compile_raw(raw)
    L, R = base.n_leaves, base.n_roots
    B = 1 << int(math.floor(math.log2(a.gb * 2 ** 30 / (8 * L))))
    leaf = torch.rand(L, B, dtype=torch.float64, device="cuda") + 0.5
    root = torch.empty(R, B, dtype=torch.float64, device="cuda") if a.mode == "eval" else None
    acc = torch.zeros(R, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    print(f"# {a.workload}: L={L} R={R} B={B} stats={base.stats}")
    rows = []
    if a.backend == 2:
        import time

        for seg, spt in itertools.product([int(x) for x in a.jit_segment.split(",")], [int(x) for x in a.spt.split(",")]):
            f = fd.compile_raw(raw, backend=2, jit_segment=seg, cse=os.environ.get("FDG_CSE") is not None)
            f.set_launch(0, spt, 0)
            t0 = time.time()
            info = f.jit_prepare(spt, a.mode != "eval")
            tc = time.time() - t0
            best = 1e30
            for r in range(a.reps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if a.mode == "eval":
                    f.eval_device(leaf.data_ptr(), B, root.data_ptr(), B, B, stream)
                else:
                    f.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), stream)
                e1.record()
                torch.cuda.synchronize()
                if r:
                    best = min(best, e0.elapsed_time(e1))
            st = f.stats
            print(json.dumps(dict(backend="jit", seg=seg, spt=spt, compile_s=round(tc, 1), ms=round(best, 3), Msamples_s=round(B / best / 1e3, 2),
                                  GBs=round(8 * L * B / best / 1e6, 1), gflops=round((st["flops_add"] + st["flops_mul"]) * B / best / 1e6, 1),
                                  **info)), flush=True)
        return
    for ms_, pf in itertools.product([int(x) for x in a.slots.split(",")], [int(x) for x in a.prefetch.split(",")]):
        f = fd.compile_raw(raw, max_slots=ms_, prefetch=pf, backend=1)
        for T, spt in itertools.product([int(x) for x in a.threads.split(",")], [int(x) for x in a.spt.split(",")]):
            try:
                f.set_launch(T, spt, 0)
                best = 1e30
                for r in range(a.reps + 1):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    if a.mode == "eval":
                        f.eval_device(leaf.data_ptr(), B, root.data_ptr(), B, B, stream)
                    else:
                        f.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), stream)
                    e1.record()
                    torch.cuda.synchronize()
                    if r:
                        best = min(best, e0.elapsed_time(e1))
                st = f.stats
                row = dict(slots=ms_, used=st["n_slots"], pf=pf, T=T, spt=spt, ms=round(best, 3), Msamples_s=round(B / best / 1e3, 2),
                           GBs=round(8 * L * B / best / 1e6, 1), gflops=round((st["flops_add"] + st["flops_mul"]) * B / best / 1e6, 1),
                           scratch=st["n_scratch"], loads=st["leaf_loads"], packets=st["n_packets"])
            except Exception as e:  # noqa: BLE001
                row = dict(slots=ms_, pf=pf, T=T, spt=spt, error=str(e)[:80])
            rows.append(row)
            print(json.dumps(row), flush=True)
    ok = [r for r in rows if "ms" in r]
    if ok:
        print("# best:", json.dumps(min(ok, key=lambda r: r["ms"])))


if __name__ == "__main__":
    main()
