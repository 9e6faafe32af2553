"""Development tool: the pipeline form of the specialised back end against the classic launch sequence on one GPU --
results (eval mode bit for bit, accumulators to rounding), kernel-only timings, per-stage busy / waiting clocks.

    python tools/exp_pipe.py --workload parquet_ver4_o4 --gb 8 "window=1776" "window=2368,FDG_PIPE_STAGES=24" ...
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


def clear_env():
    for k in [k for k in os.environ if k.startswith("FDG_")]:
        os.environ.pop(k, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="parquet_ver4_o4")
    ap.add_argument("--gb", type=float, default=8.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=1 << 16)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--no-classic", action="store_true", help="skip the classic timing (profiling runs)")
    ap.add_argument("--tune", type=int, default=0, help="rounds of profile-guided re-cutting")
    ap.add_argument("variants", nargs="*")
    a = ap.parse_args()
    raw = fd.RawGraph.load(os.path.join(ROOT, "workloads", a.workload + ".npz"))
    npdt, tdt, es = (np.float64, torch.float64, 8) if a.dtype == "f64" else (np.complex128, torch.complex128, 16)
    clear_env()
    base = fd.compile_raw(raw, dtype=npdt, backend=2)
    L, R = base.n_leaves, base.n_roots
    W = es // 8
    B = 1 << int(math.floor(math.log2(a.gb * 2 ** 30 / (es * L))))
    leaf = torch.empty(L, B, dtype=tdt, device="cuda")
    if a.dtype == "f64":
        leaf.copy_(torch.rand(L, B, dtype=torch.float64, device="cuda") + 0.5)
    else:
        torch.view_as_real(leaf).copy_(torch.rand(L, B, 2, dtype=torch.float64, device="cuda") + 0.5)
    stream = torch.cuda.current_stream().cuda_stream
    nchk = min(a.check, B)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    print(f"# {a.workload}: L={L} R={R} B={B} SMs={sms}", flush=True)

    def run(f, label, n_stages=0):
        acc = torch.zeros(R * W, dtype=torch.float64, device="cuda")
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
        print(f"{label:60s} {B / best * 1e3 / 1e6:9.2f} Msamples/s  {best:9.3f} ms", flush=True)
        return acc.clone(), best

    # classic launch sequence: the reference for everything below
    root_ref = torch.zeros(R, nchk, dtype=tdt, device="cuda")
    base.eval_device(leaf.data_ptr(), B, root_ref.data_ptr(), nchk, nchk, stream)
    torch.cuda.synchronize()
    acc_ref = None
    if not a.no_classic:
        acc_ref, _ = run(base, "classic")

    for var in a.variants or ["window=0"]:
        kv = dict(x.split("=") for x in var.split(",") if x)
        clear_env()
        window, cse = int(kv.pop("window", 0)), int(kv.pop("cse", 0))
        os.environ.update(kv)
        os.environ["FDG_JIT_PIPE"] = "1"
        os.environ["FDG_PIPE_MIN_BATCH"] = "1"
        if window:
            os.environ["FDG_PIPE_WINDOW"] = str(window)
        try:
            f = fd.compile_raw(raw, dtype=npdt, backend=2, cse=bool(cse))
            t0 = time.time()
            info = f.pipeline_prepare(True, sms)
            tc = time.time() - t0
            S = info["stages"]
            print(f"## {var}: stages={S} blocks={info['stage_blocks']} cross_rows={info['cross_rows']} "
                  f"loads={info['leaf_loads']}+{info['cross_loads']} stores={info['cross_stores']} code={info['max_code_bytes']} "
                  f"smem={info['ring_bytes']} compile={tc:.1f}s", flush=True)
            # eval mode, bit for bit against the classic kernels
            root = torch.zeros(R, nchk, dtype=tdt, device="cuda")
            f.eval_device(leaf.data_ptr(), B, root.data_ptr(), nchk, nchk, stream)
            torch.cuda.synchronize()
            st = f.pipeline_stats(stream, info_eval_stages(f, sms))
            same = bool(torch.equal(torch.view_as_real(root) if a.dtype != "f64" else root,
                                    torch.view_as_real(root_ref) if a.dtype != "f64" else root_ref))
            print(f"   eval bit-equal-to-classic={same} stalled={st['stalled']}", flush=True)
            acc, ms = run(f, "   pipeline " + var)
            st = f.pipeline_stats(stream, S)
            err = float((acc - acc_ref).abs().max()) / float(acc_ref.abs().max()) if acc_ref is not None else float("nan")
            print(f"   accumulate max|diff|/scale={err:.2e} stalled={st['stalled']}", flush=True)
            alive = np.array(st["busy"], float)
            wait = np.array(st["waiting"], float)
            blocks = np.array(info["stage_blocks"], float)
            frac = wait / np.maximum(alive, 1)
            print("   waiting fraction per stage:", " ".join(f"{x:.2f}" for x in frac), flush=True)
            # clocks of work per tile in each stage (alive - waiting, per warp-tile), relative to the estimate
            work = (alive - wait)
            est = np.array(info["stage_cost"], float)
            rel = work / work.sum() / (est / est.sum())
            print("   measured / estimated cost per stage:", " ".join(f"{x:.2f}" for x in rel), flush=True)
            n_tiles = (B + 31) // 32
            print("   busy clocks per tile:", " ".join(f"{x / n_tiles:.0f}" for x in work), flush=True)
            print("   estimated issue cycles per tile:", " ".join(f"{x:.0f}" for x in est), flush=True)
            # profile-guided re-cut: the measured busy time per estimated cost of every stage stretches the cost axis
            for rnd in range(a.tune):
                os.environ["FDG_PIPE_PREV"] = ",".join(str(x) for x in info["stage_start"])
                os.environ["FDG_PIPE_MEASURED"] = ",".join(f"{x:.0f}" for x in work)
                del f
                f = fd.compile_raw(raw, dtype=npdt, backend=2, cse=bool(cse))
                info2 = f.pipeline_prepare(True, sms)
                if info2["stages"] != S:
                    print("   (stage count changed; stop tuning)")
                    break
                acc, ms = run(f, f"   pipeline {var} tuned x{rnd + 1}")
                st = f.pipeline_stats(stream, S)
                alive = np.array(st["busy"], float)
                wait = np.array(st["waiting"], float)
                work = alive - wait
                # the estimate the next weights refer to is the unweighted one of the new cuts
                est = np.array(info2["stage_cost"], float)
                info = info2
                print("   waiting fraction per stage:", " ".join(f"{x:.2f}" for x in wait / np.maximum(alive, 1)), flush=True)
                print("   busy clocks per tile:", " ".join(f"{x / n_tiles:.0f}" for x in work), flush=True)
            del f
        except Exception as ex:  # noqa: BLE001
            print(f"{var:60s} FAILED: {ex}", flush=True)


def info_eval_stages(f, sms):
    return f.pipeline_prepare(False, sms)["stages"]


if __name__ == "__main__":
    main()
