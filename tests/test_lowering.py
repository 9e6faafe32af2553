"""CPU-only checks of the native lowering (fdg_compile): the packet program, run by the ISA emulator,
must reproduce the oracle's emitter-order values bit for bit, whatever the slot budget / prefetch distance."""
import ctypes as C

import numpy as np
import pytest

import fdgraph_b200 as fd
import graphgen
import vm_emulator as E
from fdgraph_b200 import _capi
from oracle import oracle as O


def _check(roots, dtype=np.float64, max_slots=0, prefetch=0, batch=7, signed=True, root=None, cse=False, schedule=0):
    raw, nodes = fd.flatten(roots, root)
    ev = fd.compile_raw(raw, dtype=dtype, max_slots=max_slots, prefetch=prefetch, cse=cse, schedule=schedule)
    orc = O.Oracle(raw)
    assert ev.n_leaves == orc.n_leaves and (ev.leaf_nodes == orc.leaf_nodes).all()
    assert ev.last_root == orc.last_root
    leaf = graphgen.leaf_values(11, max(ev.n_leaves, 1), batch, dtype=dtype, signed=signed)
    want = orc.eval(leaf, "emitter", root=np.full((orc.n_roots, batch), -3.0, dtype))
    got, mask, cnt = E.run(ev.program_words(), leaf, ev.n_roots)
    got[~mask] = -3.0
    assert got.tobytes() == want.tobytes(), E.disassemble(ev.program_words())
    assert cnt["max_slot"] + 1 <= ev.stats["n_slots"]
    assert cnt["ldl"] == ev.stats["leaf_loads"]
    return ev, cnt


@pytest.mark.parametrize("seed", range(12))
def test_random_dag_f64(seed):
    _check(graphgen.random_dag(seed, n_leaves=6 + seed, n_inner=30 + 5 * seed, n_roots=3))


@pytest.mark.parametrize("seed", range(6))
def test_random_dag_deep(seed):
    # deep nesting exercises the accumulator registers and the spilled-accumulator (XADDF/XMULF) path
    _check(graphgen.random_dag(100 + seed, n_leaves=5, n_inner=80, n_roots=2, max_fan=3, deep=True))


@pytest.mark.parametrize("seed", range(6))
def test_random_tree_nesting(seed):
    ev, _ = _check(graphgen.random_tree(300 + seed, depth=8), max_slots=6 + seed)
    assert ev.stats["max_depth"] >= 6
    dis = E.disassemble(ev.program_words())
    assert "RADDF" in dis and "XADDF" in dis and "[push]" in dis  # both the stack and the parked-in-a-slot combine paths


@pytest.mark.parametrize("seed", range(4))
def test_random_dag_c128(seed):
    _check(graphgen.random_dag(200 + seed, n_leaves=7, n_inner=40, n_roots=3, max_pow=5), dtype=np.complex128)


def test_power_ge_4_f64():
    _check(graphgen.random_dag(7, n_leaves=4, n_inner=30, n_roots=2, p_power=0.4, max_pow=7), signed=False)


@pytest.mark.parametrize("max_slots", [4, 5, 8, 16])
@pytest.mark.parametrize("prefetch", [-1, 1, 8, 64])
def test_small_slot_file_spills_and_prefetch(max_slots, prefetch):
    roots = graphgen.random_dag(42, n_leaves=20, n_inner=120, n_roots=4)
    ev, cnt = _check(roots, max_slots=max_slots, prefetch=prefetch)
    assert ev.stats["n_slots"] <= max(max_slots, 12)
    assert ev.stats["n_scratch"] > 0 or ev.stats["leaf_loads"] > ev.n_leaves


def test_sum_of_products_shape():
    roots = graphgen.sum_of_products(5, n_leaves=40, n_terms=300, term_len=6, n_roots=2)
    ev, cnt = _check(roots, max_slots=24, prefetch=16)
    st = ev.stats
    # single-use products are folded into the accumulators: no slot traffic for them
    assert st["n_scratch"] == 0
    assert st["flops_mul"] >= 300 * 5 and st["flops_add"] == 2 * 299


@pytest.mark.parametrize("term_len", [1, 2, 3, 4, 5, 7, 8, 11, 13])
def test_term_blocks_every_operand_count(term_len):
    ev, cnt = _check(graphgen.sum_of_products(6, n_leaves=30, n_terms=70, term_len=term_len, n_roots=2), max_slots=20, cse=False)
    if term_len <= 11:
        assert cnt["terms"] == 2 * 70


@pytest.mark.parametrize("name", ["gv_sigma_o2", "gv_sigma_o3", "gv_sigma_o4", "gv_ver4_o1", "gv_ver4_o2", "gv_ver4_o3", "gv_ver4I_o3",
                                  "parquet_ver3_o2", "parquet_ver3_o3", "parquet_polar_o3", "parquet_polar_o4"])
def test_real_workload_graphs(name):
    import json
    import os

    wl = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads")
    raw = fd.RawGraph.load(os.path.join(wl, name + ".npz"))
    ev = fd.compile_raw(raw)
    orc = O.Oracle(raw)
    leaf = graphgen.leaf_values(5, ev.n_leaves, 3, signed=True)
    got, _, _ = E.run(ev.program_words(), leaf, ev.n_roots)
    assert got.tobytes() == orc.eval(leaf).tobytes()
    # all-leaves-one checksum recorded when the workload was generated from the reference's .diag files
    man = json.load(open(os.path.join(wl, "MANIFEST.json")))[name]
    ones, _, _ = E.run(ev.program_words(), np.ones((ev.n_leaves, 1)), ev.n_roots)
    assert list(ones[:, 0]) == man["all_leaves_one"] and ev.n_leaves == man["n_leaves"]


@pytest.mark.parametrize("name", ["gv_sigma_o3", "gv_sigma_o4", "parquet_sigma_o3", "parquet_ver4_o2"])
@pytest.mark.parametrize("acc", [False, True])
def test_specialised_kernels_assemble_without_a_gpu(name, acc):
    """fdg_jit_prepare: the PTX written for the emitted function assembles for sm_100a with the toolkit's PTX compiler
    library (no driver needed); it contains only rn multiplies / adds (no fma outside Power{N>=4})."""
    import os

    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", name + ".npz"))
    evc = fd.compile_raw(raw, jit_segment=300, dtype=np.complex128)
    assert evc.jit_prepare(1, acc)["kernels"] >= 1  # ComplexF64: (re, im) pairs, one sample per thread
    ev = fd.compile_raw(raw, jit_segment=300)
    info = ev.jit_prepare(2, acc)
    assert info["kernels"] >= 1 and info["cubin_bytes"] > 0
    muls = 0
    for i in range(info["kernels"]):
        ptx, log = ev.jit_ptx(2, acc, i)
        assert ".target sm_100a" in ptx and "fma" not in ptx
        muls += ptx.count("mul.rn.f64")  # (a short last kernel may hold nothing but the tail of a sum)
        assert "registers" in log
    assert muls > 0


@pytest.mark.parametrize("name,dtype", [("parquet_ver4_o2", np.float64), ("gv_sigma_o4", np.float64), ("taylor_sigma_o2", np.complex128)])
@pytest.mark.parametrize("acc", [False, True])
def test_bulk_form_assembles_without_a_gpu(name, dtype, acc, monkeypatch):
    """The bulk form of the specialised kernels (persistent blocks: consumer warps + a producer warpgroup that fills a
    shared-memory ring with cp.async.bulk row copies, mbarrier-signalled) assembles for sm_100a on the host; nothing
    spills, the consumers' straight-line code holds no address arithmetic for its inputs."""
    import os

    monkeypatch.setenv("FDG_JIT_BULK", "1")
    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", name + ".npz"))
    ev = fd.compile_raw(raw, jit_segment=300, dtype=dtype, backend=2)
    info = ev.jit_prepare(1, acc)
    assert info["bulk"] and info["kernels"] >= 1 and 64 * 1024 < info["bulk_smem"] <= 227 * 1024
    for i in range(info["kernels"]):
        ptx, log = ev.jit_ptx(1, acc, i)
        assert ".maxntid 384" in ptx and "setmaxnreg.dec.sync.aligned.u32 24" in ptx and "setmaxnreg.inc.sync.aligned.u32 240" in ptx
        assert "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes" in ptx and "mbarrier.arrive.expect_tx" in ptx
        assert "cp.async.ca" not in ptx and "cp.async.cg" not in ptx and "fma" not in ptx
        assert "Used 168 registers" in log and " 0 bytes spill stores" in log
    monkeypatch.setenv("FDG_JIT_BULK", "0")
    assert not ev.jit_prepare(1, acc)["bulk"]  # the ring form of the same program is a separate variant


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("cse,schedule", [(True, 0), (False, 1), (True, 1)])
def test_lowering_options_do_not_change_values(seed, cse, schedule):
    _check(graphgen.random_dag(500 + seed, n_leaves=8, n_inner=90, n_roots=3), cse=cse, schedule=schedule)


def test_common_subexpressions_are_evaluated_once():
    a, b, c = fd.Graph([]), fd.Graph([]), fd.Graph([])
    def term():
        return fd.Graph([a, b], operator=fd.Prod(), subgraph_factors=[1.0, -2.0])  # three distinct node objects, same expression
    top = fd.Graph([term(), fd.Graph([term(), c], operator=fd.Prod()), term()], operator=fd.Sum(), subgraph_factors=[1.0, 3.0, 0.5])
    import os

    os.environ["FDG_CSE_MIN_COST"] = "0"
    try:
        ev, _ = _check([top], cse=True)
    finally:
        del os.environ["FDG_CSE_MIN_COST"]
    assert ev.stats["cse_removed"] == 2 and ev.stats["n_inner"] == 5
    ev2, _ = _check([top], cse=False)
    assert ev2.stats["cse_removed"] == 0
    # flop counts are the reference function's, whatever the back end saves
    assert ev.stats["flops_mul"] == ev2.stats["flops_mul"] and ev.stats["flops_add"] == ev2.stats["flops_add"]


def test_stats_counts_match_reference_operation_count():
    # count_operation (tree_properties.jl:165-185): sum (fan_in - 1) adds, prod (fan_in - 1) muls; + 1 mul per factor != 1
    a, b, c = fd.Graph([]), fd.Graph([]), fd.Graph([])
    s = fd.Graph([a, b, c], operator=fd.Sum(), subgraph_factors=[1.0, 2.0, 1.0])
    p = fd.Graph([s, a, b], operator=fd.Prod(), subgraph_factors=[1.0, 1.0, -1.0])
    ev, _ = _check([p])
    assert ev.stats["flops_add"] == 2 and ev.stats["flops_mul"] == 2 + 1 + 1
    assert ev.stats["bytes_in"] == 24 and ev.stats["bytes_out"] == 8
    evc = fd.compile_raw(ev.raw, dtype=np.complex128)
    assert evc.stats["bytes_in"] == 48 and evc.stats["flops_mul"] == 4 * 2 + 2 * 2


def test_roots_leaf_shared_and_unset():
    a, b = fd.Graph([]), fd.Graph([])
    s = fd.Graph([a, b], operator=fd.Sum())
    p = fd.Graph([s, s, a], operator=fd.Prod())
    _check([p, s, a], root=[a.id, 999, p.id, s.id, a.id])


def test_unary_chains_and_dead_nodes():
    a = fd.Graph([])
    chain = a
    for f in (2.0, 1.0, -3.0, 1.0):
        chain = fd.Graph([chain], operator=fd.Prod() if f != 1.0 else fd.Sum(), subgraph_factors=[f])
    dead = fd.Graph([a, a], operator=fd.Prod())  # computed by the emitter, never observable
    top = fd.Graph([chain, dead], operator=fd.Sum())
    _check([top, chain], root=[chain.id])


def test_long_left_deep_chain_is_not_recursive():
    g = fd.Graph([])
    acc = g
    for i in range(20000):
        acc = fd.Graph([acc, g], operator=fd.Sum(), subgraph_factors=[1.0, 0.5])
    raw, _ = fd.flatten([acc])
    ev = fd.compile_raw(raw)
    assert ev.stats["n_inner"] == 20000
    orc = O.Oracle(raw)
    leaf = np.array([[1.0, 2.0]])
    got, _, _ = E.run(ev.program_words(), leaf, 1)
    assert got.tobytes() == orc.eval(leaf).tobytes()


def test_empty_and_trivial():
    ev = fd.compile_raw(fd.flatten([])[0])
    assert ev.stats["n_leaves"] == 0 and ev.stats["n_packets"] == 1
    a = fd.Graph([])
    _check([a])  # a leaf that is itself the root


def _desc_err(raw):
    with pytest.raises(_capi.FdgError) as e:
        fd.compile_raw(raw)
    return e.value


def test_error_codes():
    a, b = fd.Graph([]), fd.Graph([])
    s = fd.Graph([a, b], operator=fd.Sum())
    raw, _ = fd.flatten([s])
    bad = fd.RawGraph(**{k: getattr(raw, k).copy() for k in raw.__dataclass_fields__})
    bad.child_node[0] = 77
    assert _desc_err(bad).code == 2  # FDG_ERR_BAD_GRAPH
    bad = fd.RawGraph(**{k: getattr(raw, k).copy() for k in raw.__dataclass_fields__})
    bad.node_op[-1] = 9  # unknown operator: static.jl:6-11 error(...)
    assert _desc_err(bad).code == 2
    # a cycle
    bad = fd.RawGraph(**{k: getattr(raw, k).copy() for k in raw.__dataclass_fields__})
    bad.child_node[0] = len(bad.node_id) - 1
    assert _desc_err(bad).code == 2
    # Power{N<2}
    p = fd.Graph([a], operator=fd.Power(2))
    rawp, _ = fd.flatten([p])
    rawp.node_pow[:] = -1
    assert _desc_err(rawp).code == 2
    with pytest.raises(TypeError):
        fd.compile_raw(raw, dtype=np.float32)  # static.jl:151 "Unsupported type"
    with pytest.raises(NotImplementedError):
        class Weird(fd.graph.Operator):
            pass
        fd.flatten([fd.Graph([a], operator=Weird())])


def test_library_exports_every_declared_symbol():
    import os
    import re

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(here, "include", "fdgraph.h")).read()
    declared = set(re.findall(r"\b(fdg_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"fdg_graph_desc", "fdg_options"}
    L = _capi.lib()
    assert declared == set(_capi.EXPORTS)
    for name in declared:
        assert getattr(L, name) is not None
    assert L.fdg_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    a, b = fd.Graph([]), fd.Graph([])
    ev, _ = fd.compile([a + b])
    with pytest.raises(_capi.FdgError) as e:
        ev(np.zeros(1), np.array([1.0, 2.0]))
    assert e.value.code in (6, 4)  # FDG_ERR_NO_DEVICE (or a CUDA error): never a silent CPU result


def test_fdgraph_file_round_trip(tmp_path):
    """a flattened graph as a file (fdg_graph_write / fdg_compile_file), written by the C library and read
    back by an independent numpy reader and by the library itself: same arrays, same lowered program, same leafmap."""
    roots = graphgen.random_dag(21, n_leaves=7, n_inner=45, n_roots=3, p_power=0.2)
    raw, _ = fd.flatten(roots, root=[roots[1].id, 424242, roots[0].id])
    path = str(tmp_path / "g.fdgraph")
    raw.save_fdg(path)
    back = fd.RawGraph.load_fdg(path)
    for k in raw.__dataclass_fields__:
        assert np.array_equal(getattr(raw, k), getattr(back, k)) and getattr(raw, k).dtype == getattr(back, k).dtype
    a, b = fd.compile_raw(raw), fd.compile_file(path)
    assert a.stats == b.stats and list(a.leaf_nodes) == list(b.leaf_nodes) and a.last_root == b.last_root
    assert np.array_equal(a.program_words(), b.program_words())
    with open(path, "ab") as fh:
        fh.write(b"x")
    with pytest.raises(_capi.FdgError) as e:
        fd.compile_file(path)  # trailing bytes
    assert e.value.code == 2
    with pytest.raises(ValueError):
        fd.RawGraph.load_fdg(path)
    with open(path, "wb") as fh:
        fh.write(b"not a graph")
    with pytest.raises(_capi.FdgError):
        fd.compile_file(path)


def test_planner_of_the_specialised_back_end_on_the_headline_graph(monkeypatch):
    """Host-side regression guard for the kernel plan of Parquet vertex4 order 4: root ordering
    keeps the cross traffic small, rows are reused, kernels stay inside the instruction-cache budget."""
    import os

    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", "parquet_ver4_o4.npz"))
    ev = fd.compile_raw(raw, backend=2, cse=False)
    info = ev.jit_prepare(1, True)
    assert info["operations"] == 94500 and not info["grid_stride"]
    assert 15 <= info["kernels"] <= 24
    assert info["cross_values"] <= 400 and info["cross_rows"] <= 160
    assert info["leaf_loads"] <= 3400                      # 984 leaves, each read by ~3.3 kernels
    assert info["cubin_bytes"] / info["kernels"] < 110e3   # straight-line code above ~100 KB stalls on instruction fetch
    ptx, log = ev.jit_ptx(1, True, 3)
    assert "cp.async.ca.shared.global" in ptx and "cp.async.wait_group" in ptx and "neg.f64" in ptx and "mad.wide.u32" in ptx
    assert "fma.rn.f64" not in ptx                         # nothing is contracted
    assert "bytes spill stores" not in log or " 0 bytes spill stores" in log
    # the automatic choice (cse=None): equal sub-expressions merged where the copies sit within one kernel of each other --
    # a third of the arithmetic goes, the traffic does not grow
    auto = fd.compile_raw(raw, backend=2).jit_prepare(1, True)
    assert auto["cse"] and auto["fp64_instr"] < 0.75 * info["fp64_instr"] and auto["kernels"] < info["kernels"]
    assert auto["leaf_loads"] + auto["cross_loads"] + auto["cross_stores"] <= info["leaf_loads"] + info["cross_loads"] + info["cross_stores"]
    assert auto["max_code_bytes"] <= 120 * 1024
    monkeypatch.setenv("FDG_JIT_ROOT_ORDER", "0")
    worse = fd.compile_raw(raw, backend=2, cse=False).jit_prepare(1, True)
    assert worse["cross_values"] > 4 * info["cross_values"]
    monkeypatch.delenv("FDG_JIT_ROOT_ORDER")
    # a budget that would overflow the 128 KB instruction cache is refitted from the machine code ptxas produced
    assert info["max_code_bytes"] <= 120 * 1024
    big = fd.compile_raw(raw, backend=2, jit_segment=9000, cse=False).jit_prepare(1, True)
    assert big["max_code_bytes"] <= 120 * 1024 and big["kernels"] >= 12
    monkeypatch.setenv("FDG_JIT_NO_REFIT", "1")
    assert fd.compile_raw(raw, backend=2, jit_segment=9000, cse=False).jit_prepare(1, True)["max_code_bytes"] > 128 * 1024


def test_root_ordering_brings_sharing_roots_together(monkeypatch):
    """Roots i and 100 + i share a sub-graph; the emitter's order keeps 100 shared values alive across ~100 roots, the
    planner's order (exact scoring, or the short list it uses for very large root sets) evaluates the pairs back to back."""
    fd.uidreset()
    rng = np.random.default_rng(3)
    leaves = [fd.Graph([]) for _ in range(40)]

    def tree(k):
        terms = [fd.Graph([leaves[int(j)] for j in rng.choice(40, 3, replace=False)], operator=fd.Prod()) for _ in range(k)]
        return fd.Graph(terms, operator=fd.Sum(), subgraph_factors=[1.0 + 0.5 * t for t in range(k)])

    shared = [tree(6) for _ in range(100)]
    roots = [fd.Graph([shared[i], tree(5)], operator=fd.Prod()) for i in range(100)]
    roots += [fd.Graph([shared[i], tree(4)], operator=fd.Prod()) for i in range(100)]
    raw, _ = fd.flatten(roots)

    def cross(env):
        for k in ("FDG_JIT_ROOT_ORDER", "FDG_JIT_ROOT_SHORTLIST"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        return fd.compile_raw(raw, backend=2, jit_segment=300).jit_prepare(1, False)["cross_values"]

    emitter, exact, short = cross({"FDG_JIT_ROOT_ORDER": "0"}), cross({}), cross({"FDG_JIT_ROOT_SHORTLIST": "1"})
    assert emitter >= 90
    assert exact <= 15 and short <= 25


def test_corrupt_fdgraph_header_is_an_error_not_an_abort(tmp_path):
    """A file whose header announces 2^31 - 1 nodes must come back as FDG_ERR_BAD_GRAPH before anything is sized from it
    (the C ABI never throws and never aborts the host process)."""
    path = str(tmp_path / "bad.fdgraph")
    with open(path, "wb") as fh:
        fh.write(b"FDGRAPH\x01" + np.array([2 ** 31 - 1, 0, 0, 0], "<i8").tobytes())
    with pytest.raises(_capi.FdgError) as e:
        fd.compile_file(path)
    assert e.value.code == 2 and "truncated" in str(e.value)
    with open(path, "wb") as fh:  # counts that agree with nothing
        fh.write(b"FDGRAPH\x01" + np.array([3, 2 ** 31 - 1, 5, 1], "<i8").tobytes() + b"\0" * 64)
    with pytest.raises(_capi.FdgError) as e:
        fd.compile_file(path)
    assert e.value.code == 2


def test_id_shared_by_a_leaf_and_an_inner_node_is_rejected():
    """static.jl:116,122 keep separate visited lists for leaves and inner nodes, so an id carried by both kinds of object
    is emitted twice by the reference (later readers see the last assignment).  The evaluator refuses the ambiguity."""
    a, b, c = fd.Graph([]), fd.Graph([]), fd.Graph([])
    inner = a * b
    c.id = inner.id  # a leaf object with the id of an inner node, e.g. graphs merged from two sessions
    raw, _ = fd.flatten([fd.Graph([inner, c], operator=fd.Sum())])  # (`inner + c` would merge the two by id)
    with pytest.raises(_capi.FdgError) as e:
        fd.compile_raw(raw)
    assert e.value.code == 2 and "leaf and by an inner node" in str(e.value)


def test_pipeline_plan_of_the_headline_graph():
    """Host-side guard for the pipeline form on Parquet vertex4 order 4 for the B200 layout of
    instruction-cache groups: one stage per group and pass, stage sizes follow the group sizes, code fits the cache."""
    import os

    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", "parquet_ver4_o4.npz"))
    ev = fd.compile_raw(raw, backend=2)
    info = ev.pipeline_prepare(True, 148)
    assert info["passes"] == 2 and info["stages"] == 16 and info["operations"] == 94500
    assert info["stage_blocks"] == [12, 18, 18, 20, 20, 20, 20, 20] * 2
    assert info["max_code_bytes"] <= 120 * 1024 and info["ring_bytes"] <= 200 * 1024
    cost, blocks = np.array(info["stage_cost"], float), np.array(info["stage_blocks"], float)
    per_pass = [(cost[p:p + 8] / blocks[p:p + 8]).max() for p in (0, 8)]
    assert sum(per_pass) / (cost.sum() / 148) < 1.08  # the slowest stage of each pass is within 8 % of a perfect split
    assert 0 < info["boundary_rows"] <= 80            # values that go from the first pass to the second


def test_julia_shim_struct_layouts_and_call_sequence(tmp_path):
    """The Julia shim cannot run here (no julia binary).  tests/abi/julia_layout.c asserts, at compile time, that the byte
    offsets of the shim's structs (= Julia's fieldoffsets: natural C layout) are those of include/fdgraph.h, then makes
    FDGraphB200.compile()'s calls in its order on the 4.5 graph of test/compiler.jl:4-15 and prints what the shim reads."""
    import os
    import subprocess

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = _capi.lib()._name
    exe = str(tmp_path / "julia_layout")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(here, "include"), os.path.join(here, "tests", "abi", "julia_layout.c"),
                    "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)], check=True, capture_output=True, text=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    assert out[0] == "L=2 R=1 leafmap=0,1 last_root=0 abi=1" and out[1] == "ok"


def test_derivative_graphs_lower_and_evaluate():
    """Graphs made by the restated graph-level AD (oracle/frontend/ad.py; operation.jl:478-543) go through the lowering
    like any other: placeholders filled in place, sums of product-rule terms, dual leaves as ordinary leaves."""
    from oracle.frontend import ad, parquet as pq

    fd.uidreset()
    pq._ver4I.clear()
    graphs = [r["diagram"] for r in pq.sigma(pq.DiagPara(type=pq.SigmaDiag, innerLoopNum=2))]
    dual = ad.build_derivative_graph(graphs, (2, 1))
    root_ids = {g.id for g in graphs}
    roots = list(graphs) + [d for (nid, _), d in sorted(dual.items(), key=lambda kv: (kv[0][0], kv[0][1])) if nid in root_ids]
    ev, _ = _check(roots)
    assert ev.n_roots == 8
