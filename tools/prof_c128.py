import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import fdgraph_b200 as fd
raw = fd.RawGraph.load("/root/repo/workloads/taylor_sigma_o3.npz")
f = fd.compile_raw(raw, dtype=np.complex128, backend=2)
f.set_launch(0, 1, 0)
B = 1 << 22
leaf = torch.empty(f.n_leaves, B, dtype=torch.complex128, device="cuda")
torch.view_as_real(leaf).copy_(torch.rand(f.n_leaves, B, 2, dtype=torch.float64, device="cuda") + 0.5)
acc = torch.zeros(2 * f.n_roots, dtype=torch.float64, device="cuda")
f.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(f.jit_last())
