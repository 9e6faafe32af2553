"""Runs one configuration a few times (for ncu captures)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import fdgraph_b200 as fd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="parquet_ver4_o4")
ap.add_argument("--samples", type=int, default=1 << 17)
ap.add_argument("--slots", type=int, default=0)
ap.add_argument("--prefetch", type=int, default=0)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--spt", type=int, default=0)
ap.add_argument("--mode", default="acc")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--backend", type=int, default=0)
ap.add_argument("--jit-segment", type=int, default=0)
a = ap.parse_args()
raw = fd.RawGraph.load(os.path.join(ROOT, "workloads", a.workload + ".npz"))
f = fd.compile_raw(raw, max_slots=a.slots, prefetch=a.prefetch, backend=a.backend, jit_segment=a.jit_segment)
f.set_launch(a.threads, a.spt, 0)
L, R, B = f.n_leaves, f.n_roots, a.samples
leaf = torch.rand(L, B, dtype=torch.float64, device="cuda") + 0.5
root = torch.empty(R, B, dtype=torch.float64, device="cuda")
acc = torch.zeros(R, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if a.mode == "eval":
        f.eval_device(leaf.data_ptr(), B, root.data_ptr(), B, B, s)
    else:
        f.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    print(f"{a.workload} B={B} {e0.elapsed_time(e1):.3f} ms  {B / e0.elapsed_time(e1) / 1e3:.2f} Msamples/s launches={f.launches}")
