"""Microbenchmark: which SMs of a B200 share an instruction cache, and how much straight-line code it holds.

The pipeline form of the specialised kernels lets different SMs run different code; whether that
works depends on who shares the instruction cache.  Here every SM runs one of several straight-line FP64 functions
(`body_kb` KB of machine code each, looped), chosen by SM id from a table, and reports its clocks:

  * pairs:  SM i runs function A alone, then together with SM j running function B, for every j -- the SMs j that slow
    SM i down share a cache level with it that 2 x body_kb does not fit;
  * groups: all SMs busy, SM s runs function (s // g) % n_func for g = 1, 2, 4, ... against "all the same function".

    python tools/icache_probe.py --kb 96 --out gpurun_out/icache.json
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np


def gen_ptx(n_func: int, n_instr: int) -> str:
    out = [".version 8.7", ".target sm_100a", ".address_size 64", ""]
    for f in range(n_func):
        out.append(f".func body{f}(.param .b64 a_iters, .param .b64 a_out)")
        out.append("{")
        out.append("\t.reg .f64 %fd<20>;\n\t.reg .b64 %rd<6>;\n\t.reg .pred %p<2>;")
        out.append("\tld.param.u64 %rd0, [a_iters];\n\tld.param.u64 %rd1, [a_out];")
        for i in range(8):
            out.append(f"\tmov.f64 %fd{i}, 0d3FF00000000{f:01X}{i:01X}000;")
        out.append(f"\tmov.f64 %fd8, 0d3FEFFFFF0000{f:01X}000;\n\tmov.f64 %fd9, 0d3F50624DD2F1A9FC;\n\tmov.f64 %fd10, 0d3FEFFFFE0000{f:01X}000;")
        out.append(f"L{f}:")
        for k in range(n_instr):
            i = k % 8
            out.append(f"\tfma.rn.f64 %fd{i}, %fd{i}, %fd{8 if (k // 8 + f) % 3 else 10}, %fd9;")
        out.append("\tsub.u64 %rd0, %rd0, 1;\n\tsetp.ne.u64 %p0, %rd0, 0;\n\t@%p0 bra L" + str(f) + ";")
        out.append("\tadd.rn.f64 %fd0, %fd0, %fd1;\n\tadd.rn.f64 %fd2, %fd2, %fd3;\n\tadd.rn.f64 %fd4, %fd4, %fd5;\n\tadd.rn.f64 %fd6, %fd6, %fd7;")
        out.append("\tadd.rn.f64 %fd0, %fd0, %fd2;\n\tadd.rn.f64 %fd4, %fd4, %fd6;\n\tadd.rn.f64 %fd0, %fd0, %fd4;")
        out.append("\tst.global.f64 [%rd1], %fd0;\n\tret;\n}")
    out.append(".visible .entry probe(.param .u64 p_tab, .param .u64 p_clk, .param .u64 p_sink, .param .u64 p_iters)")
    out.append(".maxntid 256, 1, 1\n{")
    out.append("\t.reg .b64 %rd<12>;\n\t.reg .b32 %r<6>;\n\t.reg .pred %p<3>;")
    out.append("\tld.param.u64 %rd0, [p_tab];\n\tcvta.to.global.u64 %rd0, %rd0;\n\tld.param.u64 %rd1, [p_clk];\n\tcvta.to.global.u64 %rd1, %rd1;")
    out.append("\tld.param.u64 %rd2, [p_sink];\n\tcvta.to.global.u64 %rd2, %rd2;\n\tld.param.u64 %rd3, [p_iters];")
    out.append("\tmov.u32 %r0, %smid;\n\tmul.wide.u32 %rd4, %r0, 4;\n\tadd.u64 %rd5, %rd0, %rd4;\n\tld.global.u32 %r1, [%rd5];")
    out.append("\tsetp.eq.u32 %p0, %r1, 0;\n\t@%p0 bra DONE;")
    out.append("\tmov.u32 %r2, %tid.x;\n\tmul.wide.u32 %rd6, %r2, 8;\n\tmul.wide.u32 %rd7, %r0, 2048;\n\tadd.u64 %rd6, %rd6, %rd7;\n\tadd.u64 %rd6, %rd2, %rd6;")
    out.append("\tbar.sync 0;\n\tmov.u64 %rd8, %clock64;")
    for f in range(n_func):
        out.append(f"\tsetp.eq.u32 %p1, %r1, {f + 1};\n\t@%p1 bra C{f};")
    out.append("\tbra DONE;")
    for f in range(n_func):
        out.append(f"C{f}:\n\t{{\n\t.param .b64 q0;\n\t.param .b64 q1;\n\tst.param.b64 [q0], %rd3;\n\tst.param.b64 [q1], %rd6;\n\tcall.uni body{f}, (q0, q1);\n\t}}\n\tbra FIN;")
    out.append("FIN:\n\tbar.sync 0;\n\tmov.u64 %rd9, %clock64;\n\tsub.u64 %rd9, %rd9, %rd8;")
    out.append("\tsetp.eq.u32 %p2, %r2, 0;\n\tmul.wide.u32 %rd4, %r0, 8;\n\tadd.u64 %rd10, %rd1, %rd4;\n\t@%p2 st.global.u64 [%rd10], %rd9;")
    out.append("DONE:\n\tret;\n}")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kb", type=int, default=96, help="machine code per function, KB")
    ap.add_argument("--funcs", type=int, default=8)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--out", default="")
    ap.add_argument("--pairs-for", type=int, default=3, help="how many probe SMs to run the pair test for")
    a = ap.parse_args()
    n_instr = a.kb * 1024 // 16
    ptx = gen_ptx(a.funcs, n_instr)
    # the cubin is built where there is time for it (ptxas takes a while on straight-line code) and travels with the repo
    cub_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"icache_probe_{a.kb}kb_{a.funcs}.cubin")
    if not os.path.exists(cub_path):
        with tempfile.TemporaryDirectory() as td:
            open(os.path.join(td, "p.ptx"), "w").write(ptx)
            subprocess.run(["ptxas", "-arch=sm_100a", "-O1", "-o", cub_path, os.path.join(td, "p.ptx")], check=True)
    cubin = open(cub_path, "rb").read()
    if os.environ.get("ICACHE_COMPILE_ONLY"):
        print("compiled", len(cubin), "bytes")
        return
    import torch
    from cuda.bindings import driver as cu

    torch.cuda.init()
    torch.zeros(1, device="cuda")

    def chk(r):
        if isinstance(r, tuple):
            err, rest = r[0], r[1:]
        else:
            err, rest = r, ()
        if int(err) != 0:
            raise RuntimeError(f"CUDA driver error {err}")
        return rest[0] if len(rest) == 1 else rest

    mod = chk(cu.cuModuleLoadData(cubin))
    fn = chk(cu.cuModuleGetFunction(mod, b"probe"))
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    NS = 4096
    tab = torch.zeros(NS, dtype=torch.int32, device="cuda")
    clk = torch.zeros(NS, dtype=torch.int64, device="cuda")
    sink = torch.zeros(NS * 256, dtype=torch.float64, device="cuda")
    smem = 120 * 1024  # one block per SM
    chk(cu.cuFuncSetAttribute(fn, cu.CUfunction_attribute.CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, smem))

    def launch(table: np.ndarray) -> np.ndarray:
        tab.copy_(torch.from_numpy(table.astype(np.int32)))
        clk.zero_()
        args = np.array([tab.data_ptr(), clk.data_ptr(), sink.data_ptr(), a.iters], dtype=np.uint64)
        ptrs = np.array([args[i:].ctypes.data for i in range(4)], dtype=np.uint64)
        chk(cu.cuLaunchKernel(fn, n_sm, 1, 1, 256, 1, 1, smem, 0, ptrs.ctypes.data, 0))
        torch.cuda.synchronize()
        return clk.cpu().numpy()

    # which SM ids exist
    t = launch(np.ones(NS))
    t = launch(np.ones(NS))
    smids = np.nonzero(t)[0]
    res = {"n_sm": int(n_sm), "kb": a.kb, "smids": smids.tolist(), "iters": a.iters}
    print(f"# {len(smids)} SMs report, ids {smids.min()}..{smids.max()}; {a.kb} KB per function, {a.funcs} functions", flush=True)
    same = launch(np.ones(NS))[smids]
    print(f"all SMs, same function: clocks median {np.median(same):.0f} min {same.min()} max {same.max()}", flush=True)
    res["all_same"] = float(np.median(same))
    # groups: SM s runs function (rank(s) // g) % funcs
    rank = np.zeros(NS, dtype=np.int64)
    rank[smids] = np.arange(len(smids))
    res["groups"] = {}
    for g in (1, 2, 4, 8, 16, 32, 74, 148):
        table = np.zeros(NS)
        table[smids] = (rank[smids] // g) % a.funcs + 1
        launch(table)
        tt = launch(table)[smids]
        res["groups"][g] = float(np.median(tt))
        print(f"all SMs, function = (rank // {g:3d}) % {a.funcs}: clocks median {np.median(tt):.0f} ({np.median(tt) / np.median(same):.2f}x) max {tt.max()}", flush=True)
    # how many distinct functions fit: SMs alternate between the first n functions
    res["n_distinct"] = {}
    for n in range(1, a.funcs + 1):
        table = np.zeros(NS)
        table[smids] = rank[smids] % n + 1
        launch(table)
        tt = launch(table)[smids]
        res["n_distinct"][n] = float(np.median(tt))
        print(f"all SMs, {n} distinct functions dealt round-robin: clocks median {np.median(tt):.0f} ({np.median(tt) / np.median(same):.2f}x)", flush=True)
    # pairs: SM i with function 1, SM j with function 2, everything else idle
    res["pairs"] = {}
    todo = list(smids)
    probes = []
    while todo and len(probes) < a.pairs_for:
        probes.append(todo[0])
        i = todo[0]
        table = np.zeros(NS)
        table[i] = 1
        launch(table)
        alone = launch(table)[i]
        grp = [int(i)]
        for j in todo[1:]:
            table = np.zeros(NS)
            table[i], table[j] = 1, 2
            launch(table)
            if launch(table)[i] > 1.15 * alone:
                grp.append(int(j))
        res.setdefault("groups_found", []).append(grp)
        print(f"group of SM {i} ({len(grp)} SMs): {grp}", flush=True)
        todo = [x for x in todo if int(x) not in grp]
    for i in []:
        table = np.zeros(NS)
        table[i] = 1
        launch(table)
        alone = launch(table)[i]
        slow = []
        ratios = []
        for j in smids:
            if j == i:
                ratios.append(1.0)
                continue
            table = np.zeros(NS)
            table[i], table[j] = 1, 2
            launch(table)
            tt = launch(table)[i]
            ratios.append(float(tt) / float(alone))
            if tt > 1.15 * alone:
                slow.append(int(j))
        res["pairs"][int(i)] = {"alone": int(alone), "slowed_by": slow, "ratios": [round(x, 3) for x in ratios]}
        print(f"SM {i}: alone {alone} clocks; slowed (>15 %) by SMs {slow}", flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"))


if __name__ == "__main__":
    main()
