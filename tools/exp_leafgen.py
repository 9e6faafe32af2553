"""Development tool: time the leaf-generation kernel (fdg_leafgen_fill) and the generate-and-evaluate path on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fdgraph_b200 as fd  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "parquet_ver4_o4"
B = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 21)
raw = fd.RawGraph.load(os.path.join(ROOT, "workloads", name + ".npz"))
meta = dict(np.load(os.path.join(ROOT, "workloads", name + ".leaves.npz")))
gen = fd.LeafGenerator(meta)
ev = fd.compile_raw(raw)
rows = gen.var_rows
g = torch.Generator(device="cuda").manual_seed(3)
var = torch.rand(rows, B, dtype=torch.float64, device="cuda", generator=g) * 2 - 0.7
var[gen.dim * gen.n_loops:] = torch.rand(gen.n_tau, B, dtype=torch.float64, device="cuda", generator=g) * gen.beta
leaf = torch.empty(gen.n_leaves, B, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
K, T = var.data_ptr(), var[gen.dim * gen.n_loops:].data_ptr()
best = 1e30
for r in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gen.fill_device(K, T, B, B, leaf.data_ptr(), B, s)
    e1.record()
    torch.cuda.synchronize()
    if r:
        best = min(best, e0.elapsed_time(e1))
print(f"{name}: L={gen.n_leaves} B={B}  leafgen {B / best * 1e3 / 1e6:.1f} Msamples/s ({best:.3f} ms, {gen.n_leaves * 8 * B / best / 1e6:.0f} GB/s written)")
acc = torch.zeros(ev.n_roots, dtype=torch.float64, device="cuda")
best = 1e30
for r in range(4):
    acc.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gen.accumulate_device(ev, K, T, B, B, acc.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    if r:
        best = min(best, e0.elapsed_time(e1))
print(f"   generate + evaluate (device-resident K, T): {B / best * 1e3 / 1e6:.1f} Msamples/s")
assert torch.isfinite(leaf).all()
