"""Development tool: time variants of the specialised back end on one GPU and check them bit-for-bit against the
first variant (kernel-only timings, not bench lines).

    python tools/exp_jit.py --workload parquet_ver4_o4 --gb 8 "seg=3000,FDG_JIT_ROOT_ORDER=0" "seg=6000" ...
"""
import argparse
import math
import os
import sys
import time

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
    ap.add_argument("--check", type=int, default=32768, help="samples of the eval-mode equality check")
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    raw = fd.RawGraph.load(os.path.join(ROOT, "workloads", a.workload + ".npz"))
    npdt, tdt, es = (np.float64, torch.float64, 8) if a.dtype == "f64" else (np.complex128, torch.complex128, 16)
    base = fd.compile_raw(raw, dtype=npdt)
    L, R = base.n_leaves, base.n_roots
    B = 1 << int(math.floor(math.log2(a.gb * 2 ** 30 / (es * L))))
    leaf = torch.empty(L, B, dtype=tdt, device="cuda")
    torch.view_as_real(leaf).copy_(torch.rand(L, B, 2, dtype=torch.float64, device="cuda") + 0.5) if a.dtype != "f64" else leaf.copy_(
        torch.rand(L, B, dtype=torch.float64, device="cuda") + 0.5)
    acc = torch.zeros(R * (es // 8), dtype=torch.float64, device="cuda")
    nchk = min(a.check, B)
    stream = torch.cuda.current_stream().cuda_stream
    print(f"# {a.workload}: L={L} R={R} B={B} flops/sample={base.stats['flops_add'] + base.stats['flops_mul']}", flush=True)
    ref = None
    for var in a.variants:
        kv = dict(x.split("=") for x in var.split(",") if x)
        for k in [k for k in os.environ if k.startswith("FDG_")]:
            os.environ.pop(k, None)
        seg, spt, fma = int(kv.pop("seg", 0)), int(kv.pop("spt", 1)), int(kv.pop("fma", 0))
        cse = kv.pop("cse", "0")
        cse = None if cse == "auto" else bool(int(cse))  # auto: the planner decides (FDG_CSE_MODE = 0 / 1 / 2 forces plain / merged / scoped)
        os.environ.update(kv)
        os.environ.setdefault("FDG_JIT_NO_REFIT", "1")  # experiments time the budget they ask for
        try:
            f = fd.compile_raw(raw, dtype=npdt, backend=2, jit_segment=seg, cse=cse, fma=bool(fma))
            f.set_launch(0, spt, 0)
            t0 = time.time()
            info = f.jit_prepare(spt, True)
            tc = time.time() - t0
            root = torch.zeros(R, nchk, dtype=tdt, device="cuda")
            f.eval_device(leaf.data_ptr(), B, root.data_ptr(), nchk, nchk, stream)
            torch.cuda.synchronize()
            same = None
            if ref is None:
                ref = root.clone()
            else:
                same = bool(torch.equal(torch.view_as_real(root) if a.dtype != "f64" else root,
                                        torch.view_as_real(ref) if a.dtype != "f64" else ref))
            best = 1e30
            for r in range(a.reps + 1):
                acc.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                f.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), stream)
                e1.record()
                torch.cuda.synchronize()
                if r:
                    best = min(best, e0.elapsed_time(e1))
            try:
                last = f.jit_last()  # the plan of the variant the timed launches ran (the bulk form has its own)
                info = {**info, **last, "cubin_bytes": last["max_code_bytes"] * last["kernels"]}
            except Exception:  # noqa: BLE001
                pass
            model = (info["leaf_loads"] + info["cross_loads"] + info["cross_stores"]) * es
            print(f"{var:60s} {B / best * 1e3 / 1e6:9.2f} Msamples/s  {best:9.3f} ms  kernels={info['kernels']:3d} rows={info['cross_rows']:5d} "
                  f"model={model / 1e3:6.1f} KB/sample fp64={info['fp64_instr']} code<={info['max_code_bytes'] / 1e3:6.1f} KB/kernel {'bulk' if info.get('bulk') else 'ring'}  compile={tc:5.1f}s  "
                  f"bit-equal-to-first={same}", flush=True)
            del f
        except Exception as ex:  # noqa: BLE001
            print(f"{var:60s} FAILED: {ex}", flush=True)


if __name__ == "__main__":
    main()
