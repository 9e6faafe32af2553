#!/usr/bin/env python
"""bench.py -- MC-sample graph-evals/sec of the graph-evaluation hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--samples S]

One "step" = the compiled graph set evaluated at S Monte-Carlo sample points per GPU with the per-root sums
accumulated on device (fdg_eval_accumulate) and, for N > 1, one NCCL all-reduce of the R accumulators
(fdg_allreduce).  S defaults to 2^26 (BASELINE.json).  The leaf matrix of S samples does not fit HBM for the
order-4 graphs (2^26 x 1514 x 8 B = 813 GB), so a resident batch (<= --resident-gb of HBM, >> L2) is generated
on device before timing and a step is ceil(S / resident) passes over it: every pass streams the whole resident
batch from HBM (inputs larger than L2, no cache flush needed).

`value`   graph-evals/s = samples/s x R roots, whole job over all ranks, inputs resident in HBM.
`e2e`     same metric through the reference-facing call eval_graph(root, leafVal) with HOST buffers (pinned):
          H2D of the leaves and D2H of the roots inside the timed region.
`--impl reference` times the reference's own CPU design point for the path -- the emitted C function (to_Cstr text)
compiled with gcc, one call per sample (oracle/emit_c.py; "port": no julia binary exists in this image) -- on all host
cores, on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="parquet_ver4_o4")
    ap.add_argument("--backend", type=int, default=0, help="0 auto (specialised kernels, VM fallback), 1 VM, 2 JIT")
    ap.add_argument("--jit-segment", type=int, default=0)
    ap.add_argument("--samples", type=int, default=1 << 26, help="samples per step per GPU")
    ap.add_argument("--resident-gb", type=float, default=64.0, help="device memory for the resident leaf matrix of the timed batch (GiB)")
    ap.add_argument("--e2e-samples", type=int, default=0, help="samples per e2e step (0 = about 2 GiB of leaves)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline leg")
    ap.add_argument("--dtype", default="f64", choices=["f64", "c128"])
    ap.add_argument("--max-slots", type=int, default=0)
    ap.add_argument("--prefetch", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--spt", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-torch-emitter", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the compact lines of the other BASELINE configurations")
    ap.add_argument("--fma", action="store_true", help="opt-in contraction (fdg_options.fma): NOT the bit-exact default, never the headline")
    return ap.parse_args()


def st_small(f):
    return f.stats["n_operands"] <= 400


def load_workload(name):
    import fdgraph_b200 as fd

    path = os.path.join(ROOT, "workloads", name + ".npz")
    if not os.path.exists(path):
        raise SystemExit(f"unknown workload {name!r}: {path} not found")
    return fd.RawGraph.load(path)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa(local):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's local_cpulist): the pinned
    staging buffers of the host-buffer path are then first-touched on that node and the copies do not cross sockets.
    Returns a short description for the bench line; silently does nothing where sysfs says nothing."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        cpus = open(path + "/local_cpulist").read().strip()
        node = open(path + "/numa_node").read().strip()
        want = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            want.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(want & allowed)
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            return f"process bound to the {len(use)} CPUs of NUMA node {node} (GPU {local}, {path})"
        return f"NUMA node {node}: all {len(allowed)} allowed CPUs are local"
    except Exception as ex:  # noqa: BLE001
        return f"not bound ({type(ex).__name__})"


_M64 = (1 << 64) - 1


def _i64(c):
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def fill_leaves(torch, leaf, dtype, first_sample, seed):
    """leaf[l, b] = 0.5 + u, u in [0, 1) from a counter-based generator (splitmix64 of seed, leaf, GLOBAL sample index,
    component): the value of a sample does not depend on which rank holds it or on the batch it is generated in."""
    L, B = leaf.shape
    comp = 1 if dtype == "f64" else 2
    view = leaf if comp == 1 else torch.view_as_real(leaf)
    b = torch.arange(first_sample, first_sample + B, dtype=torch.int64, device=leaf.device)
    rows = max(1, (1 << 25) // max(B, 1))

    def lsr(z, k):  # logical shift right of an int64 tensor
        return (z >> k) & _i64((1 << (64 - k)) - 1)

    for l0 in range(0, L, rows):
        l1 = min(L, l0 + rows)
        li = torch.arange(l0, l1, dtype=torch.int64, device=leaf.device)[:, None]
        for c in range(comp):
            z = (li * _i64(0x9E3779B97F4A7C15) + b[None, :] * _i64(0xD1B54A32D192ED03) + _i64(seed * 0x632BE59BD9B4E019 + c * 0x94D049BB133111EB))
            z = (z ^ lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
            z = (z ^ lsr(z, 27)) * _i64(0x94D049BB133111EB)
            z = z ^ lsr(z, 31)
            u = lsr(z, 11).to(torch.float64) * (1.0 / (1 << 53))
            if comp == 1:
                view[l0:l1] = u + 0.5
            else:
                view[l0:l1, :, c] = u + 0.5
            del z, u


def probe_fp64(torch, _capi, stream):
    """The FP64 rate of this GPU, measured in this run (fdg_probe_fp64): chains of DMUL + DADD (what the bit-exact kernels
    may issue) and of DFMA, 8 independent chains per thread, timed with CUDA events after a warm-up launch."""
    import ctypes

    sink = torch.zeros(8, dtype=torch.float64, device="cuda")
    res = {}
    for fma, key in ((0, "dmul_dadd_gops"), (1, "dfma_gops")):
        ops = ctypes.c_int64()
        best = 0.0
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _capi.check(_capi.lib().fdg_probe_fp64(fma, 20000, stream, sink.data_ptr(), ctypes.byref(ops)))
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        res[key] = best
    res["how"] = ("measured in this run: fdg_probe_fp64, 8 independent chains per thread, 8 x 256 threads per SM, ~50 ms per launch; "
                  "G instruction-lanes/s (a DFMA counts once; x2 for its flops)")
    return res


def probe_pcie(torch):
    """Pinned-memory copy bandwidth of this GPU's PCIe link, both directions, 1 GiB, best of 3 (beside the e2e number)."""
    n = 1 << 27
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    out = {}
    for key, (dst, src) in (("h2d_gbs", (d, h)), ("d2h_gbs", (h, d))):
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = max(best, n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[key] = best
    out["how"] = "torch copy_ of 1 GiB between pinned host memory and the device, CUDA events, best of 3"
    del h, d
    return out


def parity_check(torch, w, stream, n=2048):
    """The checker beside the CPU baseline: `n` strided samples of the batch the GPU was timed on, evaluated in eval mode on
    the device and by the CPU arm's oracle; the bytes must be equal (the fma opt-in: within 1e-12)."""
    from oracle import oracle as O

    f, leaf, res = w["f"], w["leaf"], w["res"]
    if leaf is None:
        return "skipped"
    stride = max(1, res // n)
    idx = torch.arange(0, res, stride, device="cuda")[:n]
    sub = leaf[:, idx].contiguous()
    nb = sub.shape[1]
    root = torch.zeros(max(w["R"], 1), nb, dtype=w["tdt"], device="cuda")
    f.eval_device(sub.data_ptr(), nb, root.data_ptr(), nb, nb, stream)
    torch.cuda.synchronize()
    want = O.Oracle(w["raw"]).eval(np.ascontiguousarray(sub.cpu().numpy()))
    got = root.cpu().numpy()[: w["R"]]
    if got.tobytes() == want.tobytes():
        return f"{nb} strided samples of the timed batch: device bytes == oracle bytes"
    err = float(np.max(np.abs(got - want) / (np.abs(want).max(axis=1, keepdims=True) + 1e-300)))
    if err <= 1e-12:
        return f"{nb} strided samples of the timed batch: max relative difference {err:.2e} (not bit-equal)"
    raise SystemExit(f"parity check failed: device result differs from the oracle by {err}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_evaluator(raw, dtype, compile_timeout):
    """The CPU arm: the reference's own design point for this path, `Compilers.compile_C` + a C compiler -- the emitted
    straight-line function (same text as to_Cstr, static.jl:155-197) built with gcc -O2 -ffp-contract=off and called
    once per sample, OpenMP over samples (oracle/emit_c.py).  Prebuilt by __graft_entry__.build() for the default
    workloads; compiled here otherwise (bounded by `compile_timeout`).  If that cannot be done in time the array-walking
    oracle (oracle/fdg_oracle.c, ~20x slower per core) runs instead, and the line says so."""
    from oracle import emit_c
    from oracle import oracle as O

    try:
        em = emit_c.Emitted(raw, dtype, timeout=compile_timeout)

        def run(leaf, root, cores):
            em.eval(leaf, root=root, nthreads=cores)

        return run, em.orc, "to_Cstr text compiled with gcc -O2 -ffp-contract=off (the reference's compile_C path), one call per sample"
    except Exception as ex:  # noqa: BLE001
        orc = O.Oracle(raw)

        def run(leaf, root, cores):
            orc.eval(leaf, mode="emitter", layout="sample", nthreads=cores, root=root)

        return run, orc, f"array-walking C port of the emitted function (gcc on the emitted text unavailable: {type(ex).__name__})"


def cpu_reference(raw, dtype, target_seconds, nthreads=0, compile_timeout=240.0):
    """Times the CPU arm on a bounded sample.  Returns (samples/s, threads, samples, seconds, description)."""
    run_fn, orc, what = cpu_evaluator(raw, dtype, compile_timeout)
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: ask the scheduler, not OpenMP)
    cores = len(os.sched_getaffinity(0)) if nthreads <= 0 else nthreads
    rng = np.random.default_rng(1234)
    npdt = np.float64 if dtype == "f64" else np.complex128

    def run(n, reps=1):
        leaf = (0.5 + rng.random((n, max(orc.n_leaves, 1)))).astype(npdt)
        root = np.zeros((n, max(orc.n_roots, 1)), npdt)
        t = time.perf_counter()
        for _ in range(reps):
            run_fn(leaf, root, cores)
        return time.perf_counter() - t

    n = max(cores * 16, 256)
    run(n)  # warm
    t = run(n)
    # about `target_seconds` of CPU work: as many samples as 2 GiB of leaves hold, evaluated `reps` times over
    n2 = int(min(max(n, n * target_seconds / max(t, 1e-6)), 1 << 24, (2 << 30) // max(orc.n_leaves * (8 if dtype == "f64" else 16), 1)))
    n2 = max(cores, (n2 // cores) * cores)
    reps = max(1, int(round(target_seconds / max(t * n2 / n, 1e-6))))
    t2 = run(n2, reps)
    return n2 * reps / t2, cores, n2 * reps, t2, what


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    raw = load_workload(a.workload)

    if a.impl == "reference":
        if rank != 0:
            return
        run_fn, orc, what = cpu_evaluator(raw, a.dtype, 240.0)
        R = orc.n_roots
        cores = len(os.sched_getaffinity(0))
        per_step = a.cpu_seconds / max(a.steps + a.warmup, 1)
        npdt = np.float64 if a.dtype == "f64" else np.complex128
        rng = np.random.default_rng(1234)
        probe = max(cores * 16, 256)
        leaf = (0.5 + rng.random((probe, max(orc.n_leaves, 1)))).astype(npdt)
        root = np.zeros((probe, max(R, 1)), npdt)
        run_fn(leaf, root, cores)
        t = time.perf_counter()
        run_fn(leaf, root, cores)
        rate0 = probe / max(time.perf_counter() - t, 1e-6)
        n = int(max(cores, min(rate0 * per_step, (2 << 30) // max(orc.n_leaves * leaf.itemsize, 1))))
        reps = max(1, int(round(rate0 * per_step / n)))  # a step = `reps` passes over the n resident samples
        leaf = (0.5 + rng.random((n, max(orc.n_leaves, 1)))).astype(npdt)
        root = np.zeros((n, max(R, 1)), npdt)
        for _ in range(a.warmup * reps):
            run_fn(leaf, root, cores)
        t = time.perf_counter()
        for _ in range(a.steps * reps):
            run_fn(leaf, root, cores)
        dt = time.perf_counter() - t
        val = n * reps * a.steps * R / dt
        sample = (f"{n * reps} samples/step ({reps} passes over {n} resident samples) x {a.steps} steps of {a.workload}, "
                  f"sample-major leaves; {what}; OpenMP over samples")
        print(json.dumps({
            "impl": "reference", "metric": "MC-sample graph-evals/sec", "value": val, "unit": "graph-evals/s",
            "samples_per_s": val / R, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.dtype, "data": "synthetic",
            "config": {"workload": a.workload, "samples_per_step": n * reps, "roots": R, "leaves": orc.n_leaves},
            "cpu_baseline": {"value": val, "unit": "graph-evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "graph-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch

    import fdgraph_b200 as fd
    from fdgraph_b200 import _capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)  # before any pinned allocation: first touch then lands on the GPU's own NUMA node
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # our own communicator for the C-ABI all-reduce; torch.distributed is only the rendezvous
        import ctypes

        ident = [None]
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            _capi.check(_capi.lib().fdg_comm_unique_id(buf))
            ident = [bytes(buf)]
        dist.broadcast_object_list(ident, src=0)
        comm = ctypes.c_void_p()
        idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(ident[0])
        _capi.check(_capi.lib().fdg_comm_init(ctypes.byref(comm), world, rank, idbuf))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    peak, peak_src = peaks()
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    stream = torch.cuda.current_stream().cuda_stream
    fp64 = probe_fp64(torch, _capi, stream) if rank == 0 else None
    pcie = probe_pcie(torch) if rank == 0 and not a.no_e2e else None

    def run_workload(name, dtype, samples, steps, warmup, headline):
        """One workload: compile, resident batch (counter-based values keyed by the GLOBAL sample index, so that the ranks of
        an N-GPU run hold the shards [rank * S, (rank + 1) * S) of one sample stream), timed accumulate steps, roofline."""
        raw_w = load_workload(name)
        npdt = np.float64 if dtype == "f64" else np.complex128
        tdt = torch.float64 if dtype == "f64" else torch.complex128
        es = 8 if dtype == "f64" else 16
        t_compile = time.perf_counter()
        f = fd.compile_raw(raw_w, dtype=npdt, max_slots=a.max_slots, prefetch=a.prefetch, backend=a.backend, jit_segment=a.jit_segment, fma=a.fma)
        jit_info = None
        if a.backend != 1:
            try:
                jit_info = f.jit_prepare(2 if (st_small(f) and dtype == "f64") else 1, True)
            except Exception:  # noqa: BLE001  (AUTO falls back to the VM inside the library)
                jit_info = None
        t_compile = time.perf_counter() - t_compile  # lowering + planning + PTX assembly (what Compilers.compile costs once)
        f.set_launch(a.threads, a.spt, 0)
        st = f.stats
        L, R, W = st["n_leaves"], st["n_roots"], (1 if dtype == "f64" else 2)
        budget = int((a.resident_gb if headline else min(a.resident_gb, 16.0)) * 2 ** 30)
        res = min(samples, max(1024, budget // max(L * es, 1)))
        res = 1 << int(math.floor(math.log2(res)))
        passes = max(1, -(-samples // res))
        samples_step = passes * res
        lo, hi = fd.shard_range(res * world, world, rank)  # this rank's shard of the global resident sample set
        assert hi - lo == res
        leaf = torch.empty(max(L, 1), res, dtype=tdt, device="cuda")
        fill_leaves(torch, leaf, dtype, first_sample=lo, seed=1234)
        acc = torch.zeros(max(R * W, 1), dtype=torch.float64, device="cuda")

        def step(events=None):
            for _ in range(passes):
                if events is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                f.accumulate_device(leaf.data_ptr(), res, res, acc.data_ptr(), stream)
                if events is not None:
                    e1.record()
                    events.append((e0, e1))
            if comm is not None:
                _capi.check(_capi.lib().fdg_allreduce(comm, acc.data_ptr(), R * W, stream))

        for _ in range(max(warmup, 0)):
            acc.zero_()
            step()
        barrier()
        sampler = ClockSampler(local) if headline else None
        if sampler and rank == 0:
            sampler.start()
        launches0 = f.launches
        events = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record()
        for _ in range(steps):
            step(events)
        t1.record()
        barrier()
        clocks = sampler.stop() if (sampler and rank == 0) else None
        ms = max_over_ranks(t0.elapsed_time(t1))
        launches = f.launches - launches0
        kern_ms = [e0.elapsed_time(e1) for e0, e1 in events]
        sps = samples_step * steps * world / (ms * 1e-3)

        # the one collective of the path, checked once outside the timed region: every rank's own sums, added up by
        # fdg_allreduce (our communicator) and by torch.distributed, must agree to rounding
        allreduce_check = None
        if comm is not None:
            acc.zero_()
            f.accumulate_device(leaf.data_ptr(), res, res, acc.data_ptr(), stream)
            torch.cuda.synchronize()
            mine = acc.clone()
            _capi.check(_capi.lib().fdg_allreduce(comm, acc.data_ptr(), R * W, stream))
            torch.cuda.synchronize()
            ref = mine.clone()
            dist.all_reduce(ref)
            mag = mine.abs().clone()
            dist.all_reduce(mag)
            rel = float(((acc - ref).abs() / (mag + 1e-300)).max().item())
            allreduce_check = {"max_rel_diff_vs_torch_distributed": rel, "ok": rel <= 64 * 2.3e-16, "n": R * W}
            if not allreduce_check["ok"]:
                raise SystemExit(f"fdg_allreduce disagrees with torch.distributed: {rel}")

        if jit_info is not None:
            # the plan of the variant the timed launches actually ran (the bulk form is a different plan from the ring form
            # prepared above: other cuts, other kernel count)
            try:
                last = f.jit_last()
                jit_info = {**jit_info, **last, "form": "bulk (persistent warp-specialised, cp.async.bulk + mbarrier ring)" if last["bulk"]
                            else ("grid-stride" if last["grid_stride"] else "ring (cp.async per thread)")}
                jit_info.pop("cubin_bytes", None)
            except Exception:  # noqa: BLE001
                pass
        avg_launch_s = 1e-3 * float(np.mean(kern_ms)) if kern_ms else float("nan")
        bytes_launch = st["bytes_in"] * res  # accumulate mode: sizeof(W) * L per sample
        achieved = bytes_launch / avg_launch_s / 1e9
        flops_launch = (st["flops_add"] + st["flops_mul"]) * res
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and jit_info is not None:
            tj = json.load(open(tp)).get(name + ("" if dtype == "f64" else "_c128"))
            if tj and tj.get("resident_samples"):
                traffic = tj["dram_bytes_per_launch"] * res / tj["resident_samples"]
                traffic_src = "profiles/traffic.json: " + tj.get("source", "ncu dram bytes per sample of an earlier run") + ", scaled to this launch"
        kernel = "fdg_vm_kernel (packet VM)" if jit_info is None else \
            f"fdg_seg0..{jit_info['kernels'] - 1} (specialised PTX kernels, {jit_info['kernels']} per pass, timed as a group)"
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": kernel,
                    "avg_launch_ms": 1e3 * avg_launch_s, "algorithmic_bytes_per_sample": st["bytes_in"],
                    "fp64_gflops_achieved": flops_launch / avg_launch_s / 1e9, "flops_per_sample": st["flops_add"] + st["flops_mul"],
                    "flop_per_byte": (st["flops_add"] + st["flops_mul"]) / max(st["bytes_in"], 1)}
        if fp64:
            # FP64 ceiling of this path, MEASURED in this run: one DMUL or DADD per lane and issue (nothing is contracted)
            roofline["fp64_peak_gflops_no_fma"] = fp64["dmul_dadd_gops"]
            roofline["fp64_peak_gflops_fma"] = fp64["dfma_gops"] * 2
            roofline["fp64_peak_source"] = fp64["how"]
            roofline["fp64_frac_algorithmic"] = roofline["fp64_gflops_achieved"] / fp64["dmul_dadd_gops"]
            if jit_info is not None:
                roofline["fp64_instr_executed_per_sample"] = jit_info["fp64_instr"]
                roofline["fp64_frac_executed"] = jit_info["fp64_instr"] * res / avg_launch_s / 1e9 / fp64["dmul_dadd_gops"]
                roofline["fp64_note"] = ("fp64_frac_algorithmic counts the operations of the REFERENCE function (flops_per_sample); the kernels "
                                         "execute fewer (x * -1 folded into the reader, equal sub-expressions merged within a kernel): "
                                         "fp64_frac_executed is the fraction of the measured DMUL+DADD rate the hardware actually runs at")
        if traffic:
            roofline["traffic_gbs"] = traffic / avg_launch_s / 1e9
            roofline["traffic_frac_of_peak"] = traffic / avg_launch_s / 1e9 / peak
        if jit_info is not None:
            roofline["planned_bytes_per_sample"] = (jit_info["leaf_loads"] + jit_info["cross_loads"] + jit_info["cross_stores"]) * es
            roofline["planned_frac_of_peak"] = roofline["planned_bytes_per_sample"] * res / avg_launch_s / 1e9 / peak
        return {"f": f, "raw": raw_w, "st": st, "jit": jit_info, "leaf": leaf, "res": res, "passes": passes, "samples_step": samples_step,
                "sps": sps, "ms": ms, "launches": launches, "clocks": clocks, "roofline": roofline, "t_compile": t_compile,
                "allreduce_check": allreduce_check, "L": L, "R": R, "W": W, "es": es, "tdt": tdt, "first_sample": lo}

    def e2e_host(w, be_cap):
        """eval_graph(root, leafVal) on pinned host arrays: H2D of the leaves, kernels, D2H of the roots, synchronous."""
        f, L, R, es, tdt = w["f"], w["L"], w["R"], w["es"], w["tdt"]
        be = a.e2e_samples or max(1024, min(a.samples, be_cap // max(L * es, 1)))
        be = 1 << int(math.floor(math.log2(be)))
        hl = torch.empty(max(L, 1), be, dtype=tdt).pin_memory()
        hr = torch.empty(max(R, 1), be, dtype=tdt).pin_memory()
        hl_np, hr_np = hl.numpy(), hr.numpy()
        hl_np[...] = 0.5 + np.random.default_rng(99 + rank).random(hl_np.shape)
        for _ in range(2):
            f(hr_np.T, hl_np.T)
        barrier()
        k = max(2, min(a.steps, 5))
        t = time.perf_counter()
        for _ in range(k):
            f(hr_np.T, hl_np.T)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t)
        out = {"value": be * k * world * R / dt, "unit": "graph-evals/s", "samples_per_s": be * k * world / dt,
               "h2d_bytes_per_step": L * es * be, "d2h_bytes_per_step": R * es * be, "samples_per_step_per_gpu": be, "steps": k,
               "host_gbs_moved": (L + R) * es * be * k * world / dt / 1e9,
               "api": "Compilers.compile(graphs) -> eval_graph(root, leafVal) on pinned host arrays (fdg_eval_host)"}
        del hl, hr
        return out

    # ================================ the headline workload ================================
    w = run_workload(a.workload, a.dtype, a.samples, a.steps, a.warmup, headline=True)
    f, st, jit_info, leaf, res, R, L, W, es = w["f"], w["st"], w["jit"], w["leaf"], w["res"], w["R"], w["L"], w["W"], w["es"]
    raw = w["raw"]
    out = {
        "metric": "MC-sample graph-evals/sec", "value": w["sps"] * R, "unit": "graph-evals/s", "samples_per_s": w["sps"],
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": w["ms"] / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": a.workload, "mode": "accumulate (per-root sums on device" + (", NCCL all-reduce)" if world > 1 else ")"),
                   "samples_per_step_per_gpu": w["samples_step"], "resident_samples": res, "passes_per_step": w["passes"],
                   "leaves": L, "statements": st["n_inner"], "roots": R,
                   "backend": "vm" if jit_info is None else "jit", "jit": jit_info,
                   "arithmetic": "fma opt-in (not bit-identical)" if a.fma else "bit-exact (no contraction)", "compile_seconds": round(w["t_compile"], 2),
                   "vm_packets": st["n_packets"], "vm_slots": st["n_slots"],
                   "l2": f"resident inputs {L * es * res / 2 ** 30:.1f} GiB per GPU >> 126 MB L2, no flush needed; a step is "
                         f"{w['passes']} passes over the SAME resident samples (2^26 distinct samples of this graph are 528 GB)",
                   "leaf_values": "0.5 + U[0,1) from a counter-based generator keyed (seed 1234, leaf, GLOBAL sample index): rank r of N holds "
                                  "samples [r S, (r + 1) S) of one stream (sharding.shard_range), so rank 0's shard is the 1-GPU set",
                   "numa": numa},
        "roofline": w["roofline"], "gpu_launches": int(w["launches"]), "clocks": w["clocks"],
    }
    if w["allreduce_check"]:
        out["allreduce_check"] = w["allreduce_check"]

    if not a.no_e2e:
        out["e2e"] = e2e_host(w, 2 << 30)
        if pcie:
            out["e2e"]["pcie_pinned_h2d_gbs"], out["e2e"]["pcie_pinned_d2h_gbs"] = pcie["h2d_gbs"], pcie["d2h_gbs"]
            out["e2e"]["pcie_probe"] = pcie["how"]

    # ---- e2e with the leaves generated on the device: host (K, tau) in, R sums out ----------------
    side = os.path.join(ROOT, "workloads", a.workload + ".leaves.npz")
    if not a.no_e2e and a.dtype == "f64" and os.path.exists(side):
        meta = dict(np.load(side))
        gen = fd.LeafGenerator(meta)  # kF, beta, lambda of example/benchmark.jl:11-18
        bg = 1 << 22
        rows = gen.var_rows
        hv = torch.empty(rows, bg, dtype=torch.float64).pin_memory()
        hv_np = hv.numpy()
        rng = np.random.default_rng(7 + rank)
        hv_np[...] = rng.random(hv_np.shape) * 2 - 0.7
        hv_np[gen.dim * gen.n_loops:] = rng.random((gen.n_tau, bg)) * gen.beta
        hK, hT = hv_np[: gen.dim * gen.n_loops], hv_np[gen.dim * gen.n_loops:]
        for _ in range(2):
            gen.accumulate_host(f, hK, hT)
        barrier()
        k_gen = max(2, min(a.steps, 5))
        t = time.perf_counter()
        for _ in range(k_gen):
            gen.accumulate_host(f, hK, hT)  # H2D of (K, tau), leaf generation, graph kernels, D2H of the R sums; synchronous
        dt = time.perf_counter() - t
        # the generation kernel alone, on device-resident variables
        dv = hv.cuda()
        nfill = min(bg, res)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gen.fill_device(dv.data_ptr(), dv[gen.dim * gen.n_loops:].data_ptr(), bg, nfill, leaf.data_ptr(), res, stream)
        e0.record()
        gen.fill_device(dv.data_ptr(), dv[gen.dim * gen.n_loops:].data_ptr(), bg, nfill, leaf.data_ptr(), res, stream)
        e1.record()
        torch.cuda.synchronize()
        dt = max_over_ranks(dt)
        out["e2e_generated"] = {
            "value": bg * k_gen * world * R / dt, "unit": "graph-evals/s", "samples_per_s": bg * k_gen * world / dt,
            "h2d_bytes_per_step": rows * 8 * bg, "d2h_bytes_per_step": R * 8, "samples_per_step_per_gpu": bg, "steps": k_gen,
            "leafgen_kernel_samples_per_s": nfill / (e0.elapsed_time(e1) * 1e-3),
            "api": "fdg_eval_generated_host: host (K, tau) of example/benchmark.jl:44-53 -> leaves on device (fdg_leafgen) -> "
                   "graph kernels -> R per-root sums; leaf values overwrite the synthetic resident batch AFTER the timed region above"}
        del hv, dv
        fill_leaves(torch, leaf, a.dtype, first_sample=w["first_sample"], seed=1234)  # the synthetic values again (parity check below)

    # ---- the reference's own batched design point on this GPU: the emitted torch function (compiler_python.jl) ------
    if rank == 0 and world == 1 and not a.no_torch_emitter and a.dtype == "f64":
        try:
            import fdgraph_b200.emitters as em

            graphs = raw.to_graphs()
            text, _ = em.to_python_str(graphs, root=[int(r) for r in raw.root_id])
            ns = {}
            exec(compile(text, "<to_python_str>", "exec"), ns)  # one elementwise torch op per node, every g<ID> kept alive
            n_stmt = st["n_inner"] + L
            bt = max(256, min(1 << 20, int(6 * 2 ** 30 // (8 * max(n_stmt, 1)))))  # about 6 GiB of live intermediates
            bt = 1 << int(math.floor(math.log2(bt)))
            lt = (torch.rand(bt, max(L, 1), dtype=torch.float64, device="cuda") + 0.5).T.contiguous().T  # leafVal[:, k] unit-stride
            ns["eval_graph"](lt)
            torch.cuda.synchronize()
            reps, t = 0, time.perf_counter()
            while reps < 3 or (time.perf_counter() - t < 2.0 and reps < 50):
                ns["eval_graph"](lt)
                reps += 1
            torch.cuda.synchronize()
            dt = time.perf_counter() - t
            out["reference_torch_emitter"] = {
                "value": bt * reps * R / dt, "unit": "graph-evals/s", "samples_per_s": bt * reps / dt, "batch": bt, "calls": reps,
                "what": "the text of Compilers.to_python_str (compiler_python.jl:9-52, fdgraph_b200.emitters) executed by torch on this "
                        "GPU: one kernel launch per node -- what the reference offers on a GPU today"}
            del lt, ns
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            out["reference_torch_emitter"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    # ---- CPU baseline beside it (rank 0, N=1 only), and the checker on a strided subset of the timed batch ---------------
    if rank == 0 and world == 1 and not a.no_cpu:
        rate, cores, n, secs, what = cpu_reference(raw, a.dtype, a.cpu_seconds)
        out["cpu_baseline"] = {"value": rate * R, "unit": "graph-evals/s", "cores": cores, "kind": "port",
                               "sample": f"{n} samples of {a.workload} in {secs:.1f} s; {what}; sample-major leaves, OpenMP over samples"}
        out["cpu_baseline"]["parity_check"] = parity_check(torch, w, stream)

    # ---- every other BASELINE configuration, compact (each in its own CUDA-event region; the headline above is cfg4) ------
    if not a.no_configs and a.workload == "parquet_ver4_o4" and a.dtype == "f64":
        del leaf
        w["leaf"] = None
        torch.cuda.empty_cache()
        cfgs = {}
        for key, name, dtype, samples in (("cfg1", "parquet_sigma_o2", "f64", 1 << 16), ("cfg2", "parquet_sigma_o3", "f64", 1 << 24),
                                         ("cfg3", "gv_ver4_o4", "f64", 1 << 26), ("cfg5", "taylor_sigma_o3", "c128", 1 << 24)):
            wc = run_workload(name, dtype, samples, steps=3, warmup=3, headline=False)
            rl = wc["roofline"]
            c = {"workload": name, "dtype": dtype, "samples_per_step_per_gpu": wc["samples_step"], "samples_per_s": wc["sps"],
                 "graph_evals_per_s": wc["sps"] * wc["R"], "ms_per_step": wc["ms"] / 3, "kernels_per_pass": (wc["jit"] or {}).get("kernels"),
                 "frac_of_hbm_peak_algorithmic": rl["frac"], "frac_of_hbm_peak_planned_traffic": rl.get("planned_frac_of_peak"),
                 "planned_bytes_per_sample": rl.get("planned_bytes_per_sample"), "algorithmic_bytes_per_sample": rl["algorithmic_bytes_per_sample"],
                 "fp64_frac_executed": rl.get("fp64_frac_executed"), "cse": (wc["jit"] or {}).get("cse"), "form": (wc["jit"] or {}).get("form"),
                 "gpu_launches": int(wc["launches"])}
            if not a.no_e2e:
                e = e2e_host(wc, 1 << 29)
                c["e2e_samples_per_s"], c["e2e_graph_evals_per_s"] = e["samples_per_s"], e["value"]
            if rank == 0 and world == 1 and not a.no_cpu:
                try:
                    rate, cores, n, secs, what = cpu_reference(wc["raw"], dtype, 3.0, compile_timeout=5.0)
                    c["cpu_samples_per_s"], c["cpu_cores"] = rate, cores
                    c["cpu_kind"] = "emitted C" if what.startswith("to_Cstr") else "array-walking port"
                    c["parity_check"] = parity_check(torch, wc, stream)
                except Exception as ex:  # noqa: BLE001
                    c["cpu_samples_per_s"] = None
                    c["cpu_note"] = f"{type(ex).__name__}"
            cfgs[key] = c
            del wc
            torch.cuda.empty_cache()
        cfgs["cfg4"] = "the headline of this line"
        out["configs"] = cfgs

    if rank == 0:
        print(json.dumps(out))
    if comm is not None:
        _capi.lib().fdg_comm_destroy(comm)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
