import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
os.environ["FDG_JIT_BULK"] = "1"
import numpy as np, torch
import fdgraph_b200 as fd
import graphgen
from oracle import oracle as O
raw = fd.RawGraph.load("/root/repo/workloads/parquet_ver4_o3.npz")
for acc in (False, True):
    ev = fd.compile_raw(raw, backend=2, jit_segment=700)
    ev.set_launch(0, 1, 0)
    B = 3001
    leaf = graphgen.leaf_values(3, ev.n_leaves, B, signed=True, ld=B + 1)
    d = torch.from_numpy(leaf).cuda()
    s = torch.cuda.current_stream().cuda_stream
    if acc:
        a = torch.zeros(ev.n_roots, dtype=torch.float64, device="cuda")
        ev.accumulate_device(d.data_ptr(), B + 1, B, a.data_ptr(), s)
    else:
        r = torch.zeros(ev.n_roots, B, dtype=torch.float64, device="cuda")
        ev.eval_device(d.data_ptr(), B + 1, r.data_ptr(), B, B, s)
    torch.cuda.synchronize()
    assert ev.jit_last()["bulk"]
    if not acc:
        assert r.cpu().numpy().tobytes() == O.Oracle(raw).eval(np.ascontiguousarray(leaf[:, :B])).tobytes()
print("sanitizer run ok")
