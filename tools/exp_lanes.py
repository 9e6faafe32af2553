"""Development tool: the headline graph's launch sequences on ONE stream (every kernel owns the device) against TWO
streams whose persistent kernels take half of the SMs each (FDG_JIT_BULK_GRID), so that a DRAM-bound kernel of one
sequence runs beside an FP64-bound kernel of the other.  Kernel-only timings, accumulate mode; the sums of every
variant are compared with the first one's.

    python tools/exp_lanes.py --workload parquet_ver4_o4 --gb 8 "lanes=1" "lanes=2,grid=74,parts=4" "lanes=2,parts=4"
"""
import argparse
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
    ap.add_argument("--workload", default="parquet_ver4_o4")
    ap.add_argument("--gb", type=float, default=8.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    raw = fd.RawGraph.load(os.path.join(ROOT, "workloads", a.workload + ".npz"))
    f = fd.compile_raw(raw, dtype=np.float64)
    L, R = f.n_leaves, f.n_roots
    B = 1 << int(math.floor(math.log2(a.gb * 2 ** 30 / (8 * L))))
    leaf = torch.rand(L, B, dtype=torch.float64, device="cuda") + 0.5
    lanes_s = [torch.cuda.Stream(), torch.cuda.Stream()]
    ref = None
    print(f"# {a.workload}: L={L} R={R} B={B}", flush=True)
    for var in a.variants:
        kv = dict(x.split("=") for x in var.split(",") if x)
        lanes, parts, grid, skew = int(kv.get("lanes", 1)), int(kv.get("parts", 1)), kv.get("grid"), float(kv.get("skew", 0))
        os.environ.pop("FDG_JIT_BULK_GRID", None)
        if grid:
            os.environ["FDG_JIT_BULK_GRID"] = grid
        # parts of whole tiles; with a skew the first part of lane 1 is shorter, which sets the two sequences out of phase
        tile = 256
        bounds = [B * i // parts // tile * tile for i in range(parts)] + [B]
        if skew and parts >= 2:
            bounds[2 if parts > 2 else 1] -= int((bounds[2 if parts > 2 else 1] - bounds[1 if parts > 2 else 0]) * skew) // tile * tile
        accs = [torch.zeros(R, dtype=torch.float64, device="cuda") for _ in range(lanes)]
        best = 1e30
        for r in range(a.reps + 1):
            for x in accs:
                x.zero_()
            main_s = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if lanes == 1:
                for p in range(parts):
                    f.accumulate_device(leaf.data_ptr() + 8 * bounds[p], B, bounds[p + 1] - bounds[p], accs[0].data_ptr(), main_s.cuda_stream)
            else:
                for s in lanes_s:
                    s.wait_stream(main_s)
                for p in range(parts):
                    s = lanes_s[p % 2]
                    f.accumulate_device(leaf.data_ptr() + 8 * bounds[p], B, bounds[p + 1] - bounds[p], accs[p % 2].data_ptr(), s.cuda_stream)
                for s in lanes_s:
                    main_s.wait_stream(s)
            e1.record()
            torch.cuda.synchronize()
            if r:
                best = min(best, e0.elapsed_time(e1))
        total = sum(accs).cpu().numpy()
        if ref is None:
            ref = total
        err = float(np.max(np.abs(total - ref) / np.maximum(np.abs(ref), 1e-300)))
        last = f.jit_last()
        print(f"{var:40s} {B / best * 1e3 / 1e6:9.2f} Msamples/s {best:9.3f} ms  kernels={last['kernels']} {'bulk' if last.get('bulk') else 'ring'}  "
              f"max rel. difference of the sums to the first variant {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
