"""Parity tests proper: the CUDA path, called through the C ABI (ctypes / torch pointers), against the CPU
oracle on the same seeded inputs.  Bar: BIT-EXACT for Float64 and ComplexF64 (Sum, Prod, Power{2,3}; Power{N>=4}
uses the same compensated algorithm on both sides and is compared bit-exactly as well)."""
import numpy as np
import pytest

import fdgraph_b200 as fd
import graphgen
from fdgraph_b200 import _capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev_eval(ev, leaf_host: np.ndarray, batch: int, spt: int = 0, threads: int = 0, fill=-3.0):
    """leaf_host (L, ld) -> root (R, batch) through fdg_eval with device buffers."""
    ev.set_launch(threads, spt, 0)
    dt = torch.float64 if leaf_host.dtype == np.float64 else torch.complex128
    leaf = torch.from_numpy(leaf_host).cuda()
    ld = leaf_host.shape[1]
    root = torch.full((max(ev.n_roots, 1), ld), fill, dtype=dt, device="cuda")
    ev.eval_device(leaf.data_ptr(), ld, root.data_ptr(), ld, batch, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return root.cpu().numpy()[: ev.n_roots, :batch]


VM, JIT = 1, 2  # FDG_BACKEND_VM (packet interpreter kernel), FDG_BACKEND_JIT (specialised PTX kernels)
BACKENDS = [VM, JIT]


def _parity(roots, dtype=np.float64, batch=1000, ld=0, spt=0, threads=0, max_slots=0, prefetch=0, signed=True, root=None,
            backend=VM, jit_segment=0):
    raw, _ = fd.flatten(roots, root)
    ev = fd.compile_raw(raw, dtype=dtype, max_slots=max_slots, prefetch=prefetch, backend=backend, jit_segment=jit_segment)
    orc = O.Oracle(raw)
    leaf = graphgen.leaf_values(5, max(ev.n_leaves, 1), batch, dtype=dtype, signed=signed, ld=ld)
    want = orc.eval(np.ascontiguousarray(leaf[:, :batch]), "emitter", root=np.full((orc.n_roots, batch), -3.0, dtype))
    got = _dev_eval(ev, leaf, batch, spt, threads)
    assert got.tobytes() == want.tobytes()
    return ev


def test_known_answers_through_the_public_call():
    # reference test/compiler.jl:18-31 (4.5) and test/computational_graph.jl:874-887 (26, 27, 702)
    v1, v2 = fd.FeynmanGraph([]), fd.FeynmanGraph([])
    g = fd.FeynmanGraph([v1, v2], factor=1.5)
    eval_graph, leafmap = fd.Compilers.compile([g])
    root = np.array([0.0])
    leaf = np.array([1.0, 2.0])
    assert eval_graph(root, leaf) == 4.5 == (leaf[0] + leaf[1]) * 1.5
    assert root[0] == 4.5 and leafmap[0] is v1 and leafmap[1] is v2
    g1, g2 = fd.Graph([]), fd.Graph([], factor=2)
    g3 = 2 * (3 * g1 + 5 * g2)
    g4 = g1 + 2 * (3 * g1 + 5 * g2)
    g5 = g4 * g3
    f, _ = fd.compile([g3, g4, g5])
    root = np.zeros(3)
    assert f(root, np.ones(2)) == 702.0
    assert list(root) == [26.0, 27.0, 702.0]


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("spt", [1, 2, 4])
def test_random_dag_f64(seed, spt):
    _parity(graphgen.random_dag(seed, n_leaves=6 + seed, n_inner=40 + 10 * seed, n_roots=3), spt=spt, batch=1536 + 4 * seed)


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("spt", [1, 2])
@pytest.mark.parametrize("jit_segment", [0, 37])
def test_random_dag_f64_jit(seed, spt, jit_segment):
    # jit_segment=37 cuts the program into many small kernels: exercises the cross-segment buffer
    _parity(graphgen.random_dag(seed, n_leaves=6 + seed, n_inner=40 + 10 * seed, n_roots=3), spt=spt, batch=1536 + 4 * seed,
            backend=JIT, jit_segment=jit_segment)


@pytest.mark.parametrize("spt", [1, 2])
@pytest.mark.parametrize("jit_segment", [0, 300])
def test_many_inputs_wrap_the_ring(spt, jit_segment):
    # more input rows than the cp.async ring holds (32 rows of 8 B, 24 of 16 B): slots are refilled while the kernel runs
    _parity(graphgen.random_dag(77, n_leaves=190, n_inner=900, n_roots=7), spt=spt, batch=4100, ld=4100, backend=JIT,
            jit_segment=jit_segment)


def test_many_inputs_wrap_the_ring_c128():
    _parity(graphgen.random_dag(78, n_leaves=120, n_inner=500, n_roots=5), dtype=np.complex128, batch=1031, ld=1031, backend=JIT)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("backend,jit_segment", [(VM, 0), (JIT, 0), (JIT, 41)])
def test_random_dag_c128(seed, backend, jit_segment):
    _parity(graphgen.random_dag(200 + seed, n_leaves=7, n_inner=60, n_roots=3, max_pow=5), dtype=np.complex128, batch=777,
            backend=backend, jit_segment=jit_segment)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("backend", BACKENDS)
def test_random_tree_nesting(seed, backend):
    _parity(graphgen.random_tree(300 + seed, depth=8), max_slots=6 + seed, batch=515, ld=516, backend=backend)


@pytest.mark.parametrize("backend", BACKENDS)
def test_power_ge_4(backend):
    _parity(graphgen.random_dag(7, n_leaves=4, n_inner=30, n_roots=2, p_power=0.4, max_pow=7), signed=False, backend=backend)


@pytest.mark.parametrize("max_slots,prefetch", [(4, -1), (5, 8), (8, 64), (16, 1), (64, 24)])
def test_spills_and_prefetch(max_slots, prefetch):
    roots = graphgen.random_dag(42, n_leaves=20, n_inner=120, n_roots=4)
    ev = _parity(roots, max_slots=max_slots, prefetch=prefetch, batch=4100, ld=4102)
    if max_slots <= 16:
        assert ev.stats["n_scratch"] > 0 or ev.stats["leaf_loads"] > ev.n_leaves  # the small slot file was really exercised


@pytest.mark.parametrize("batch,ld", [(1, 1), (1, 2), (2, 2), (3, 3), (3, 4), (31, 32), (33, 40), (255, 256), (257, 257), (1025, 1026)])
@pytest.mark.parametrize("backend", BACKENDS)
def test_ragged_batches(batch, ld, backend):
    # odd / tiny batches, padded leading dimensions, the one- and two-samples-per-thread kernels
    _parity(graphgen.random_dag(9, n_leaves=9, n_inner=50, n_roots=3), batch=batch, ld=ld, backend=backend)


@pytest.mark.parametrize("threads", [32, 64, 128, 256])
@pytest.mark.parametrize("spt", [1, 2, 4])
def test_block_shapes(threads, spt):
    _parity(graphgen.sum_of_products(5, n_leaves=40, n_terms=300, term_len=6, n_roots=2), threads=threads, spt=spt,
            batch=5000, max_slots=24, prefetch=16)


@pytest.mark.parametrize("term_len", [1, 2, 3, 4, 5, 7, 8, 11, 13])
def test_term_blocks_every_operand_count(term_len):
    _parity(graphgen.sum_of_products(6, n_leaves=30, n_terms=70, term_len=term_len, n_roots=2), batch=2050, ld=2052)


@pytest.mark.parametrize("name", ["gv_sigma_o3", "gv_ver4_o2", "gv_ver4_o3", "gv_sigma_o5", "parquet_sigma_o2", "parquet_sigma_o3",
                                  "parquet_sigma_o4", "parquet_ver4_o3", "taylor_sigma_o3", "parquet_ver3_o3", "parquet_ver3_o4",
                                  "parquet_polar_o4", "parquet_polar_o5"])
@pytest.mark.parametrize("spt,backend", [(2, VM), (4, VM), (1, JIT), (2, JIT)])
def test_real_workload_graphs(name, spt, backend):
    import os

    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", name + ".npz"))
    ev = fd.compile_raw(raw, backend=backend)
    orc = O.Oracle(raw)
    batch = 4096
    leaf = graphgen.leaf_values(5, ev.n_leaves, batch, signed=True)
    got = _dev_eval(ev, leaf, batch, spt)
    assert got.tobytes() == orc.eval(leaf).tobytes()


@pytest.mark.parametrize("name", ["taylor_sigma_o2", "taylor_sigma_o3", "taylor_sigma_o4", "parquet_ver4_o2"])
@pytest.mark.parametrize("backend", BACKENDS)
def test_real_workload_graphs_complex(name, backend):
    # BASELINE config 5: the Taylor-mode AD self-energy graphs with ComplexF64 leaves
    import os

    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", name + ".npz"))
    ev = fd.compile_raw(raw, dtype=np.complex128, backend=backend)
    orc = O.Oracle(raw)
    batch = 2048
    leaf = graphgen.leaf_values(6, ev.n_leaves, batch, dtype=np.complex128, signed=True)
    got = _dev_eval(ev, leaf, batch)
    assert got.tobytes() == orc.eval(leaf).tobytes()


def test_roots_unset_columns_are_left_untouched():
    a, b = fd.Graph([]), fd.Graph([])
    s = fd.Graph([a, b], operator=fd.Sum())
    p = fd.Graph([s, s, a], operator=fd.Prod())
    _parity([p, s, a], root=[a.id, 999, p.id, s.id, a.id], batch=100)


def test_empty_batch_and_empty_program():
    a, b = fd.Graph([]), fd.Graph([])
    ev, _ = fd.compile([a + b])
    ev.eval_device(0, 0, 0, 0, 0)  # batch 0: no launch, no error
    ev0 = fd.compile_raw(fd.flatten([])[0])
    ev0.eval_device(0, 8, 0, 8, 8)
    torch.cuda.synchronize()


def test_bad_arguments_are_reported_not_executed():
    a, b = fd.Graph([]), fd.Graph([])
    ev, _ = fd.compile([a + b])
    x = torch.zeros(64, dtype=torch.float64, device="cuda")
    with pytest.raises(_capi.FdgError) as e:
        ev.eval_device(x.data_ptr(), 4, x.data_ptr(), 8, 8)  # ld_leaf < batch
    assert e.value.code == 1
    with pytest.raises(_capi.FdgError):
        ev.eval_device(0, 8, x.data_ptr(), 8, 8)  # null leaf


@pytest.mark.parametrize("backend", BACKENDS)
def test_accumulate_is_the_sum_of_eval_and_deterministic(backend):
    roots = graphgen.random_dag(21, n_leaves=10, n_inner=60, n_roots=4)
    raw, _ = fd.flatten(roots)
    for dtype, w in ((np.float64, 1), (np.complex128, 2)):
        ev = fd.compile_raw(raw, dtype=dtype, backend=backend, jit_segment=50 if backend == JIT else 0)
        batch = 100_003
        leaf_h = graphgen.leaf_values(3, ev.n_leaves, batch, dtype=dtype, signed=True, ld=batch + 1)
        per_sample = _dev_eval(ev, leaf_h, batch)
        leaf = torch.from_numpy(leaf_h).cuda()
        accs = []
        for _ in range(2):
            acc = torch.zeros(ev.n_roots * w, dtype=torch.float64, device="cuda")
            ev.accumulate(leaf.T[:batch], acc)
            ev.accumulate(leaf.T[:batch], acc)  # accumulates on top
            torch.cuda.synchronize()
            accs.append(acc.cpu().numpy())
        assert accs[0].tobytes() == accs[1].tobytes()  # fixed-order reduction: run-to-run identical
        want = 2 * per_sample.sum(axis=1)
        scale = 2 * np.abs(per_sample).sum(axis=1)
        got = accs[0].view(dtype) if w == 2 else accs[0]
        assert (np.abs(got - want) <= 1e-12 * scale).all()


def test_host_buffers_through_eval_graph_call():
    # the reference-facing call with ordinary host arrays: (B, L) in, (B, R) out, batch-major or not
    roots = graphgen.random_dag(31, n_leaves=12, n_inner=70, n_roots=3)
    f, leafmap = fd.compile(roots)
    orc = O.Oracle(f.raw)
    B = 70_001
    leafT = graphgen.leaf_values(8, f.n_leaves, B)
    want = orc.eval(leafT).T
    for order in ("F", "C"):
        leafVal = np.asarray(leafT.T, order=order)
        root = np.zeros((B, f.n_roots), order=order)
        ret = f(root, leafVal)
        assert root.tobytes(order="C") == want.tobytes(order="C")
        assert (ret == want[:, f.last_root]).all()


def test_torch_tensors_batch_major():
    roots = graphgen.random_dag(32, n_leaves=12, n_inner=70, n_roots=3)
    f, _ = fd.compile(roots)
    orc = O.Oracle(f.raw)
    B = 4097
    leafT = graphgen.leaf_values(8, f.n_leaves, B)
    leafVal = torch.from_numpy(leafT).cuda().T  # (B, L) with strides (1, B)
    root = torch.empty(f.n_roots, B, dtype=torch.float64, device="cuda").T
    f(root, leafVal)
    torch.cuda.synchronize()
    assert root.T.contiguous().cpu().numpy().tobytes() == orc.eval(leafT).tobytes()
    with pytest.raises(ValueError):
        f(torch.empty(B, f.n_roots, dtype=torch.float64, device="cuda"), leafVal.contiguous())


def test_auto_backend_and_backends_agree():
    roots = graphgen.random_dag(55, n_leaves=10, n_inner=60, n_roots=3)
    _parity(roots, backend=0, batch=3000)  # AUTO: specialised kernels
    _parity(roots, backend=0, batch=3000, dtype=np.complex128)
    # the two back ends agree bit for bit on the same program and inputs
    raw, _ = fd.flatten(roots)
    leaf = graphgen.leaf_values(1, 11, 512, dtype=np.complex128, signed=True)
    a = _dev_eval(fd.compile_raw(raw, dtype=np.complex128, backend=VM), leaf, 512)
    b = _dev_eval(fd.compile_raw(raw, dtype=np.complex128, backend=JIT), leaf, 512)
    assert a.tobytes() == b.tobytes()


def test_large_batch_properties():
    """Full-size style checks that do not need the oracle on every sample: linearity in the root factors and
    agreement of a strided subsample with the oracle."""
    roots = graphgen.sum_of_products(77, n_leaves=32, n_terms=60, term_len=4, n_roots=2)
    scaled = [fd.Graph([r], operator=fd.Prod(), subgraph_factors=[2.0]) for r in roots]  # exact: *2
    f, _ = fd.compile(roots + scaled)
    B = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(1234)
    leaf = (torch.rand(f.n_leaves, B, dtype=torch.float64, device="cuda", generator=g) + 0.5)
    root = torch.empty(f.n_roots, B, dtype=torch.float64, device="cuda")
    f(root.T, leaf.T)
    torch.cuda.synchronize()
    assert torch.equal(root[2], 2 * root[0]) and torch.equal(root[3], 2 * root[1])
    idx = torch.arange(0, B, 4099, device="cuda")
    sub = leaf[:, idx].cpu().numpy()
    want = O.Oracle(f.raw).eval(np.ascontiguousarray(sub))
    assert root[:, idx].cpu().numpy().tobytes() == want.tobytes()


@pytest.mark.parametrize("bulk", [False, True])
def test_leading_dimension_of_4_gib_and_more(bulk, monkeypatch):
    """Row offsets are one 32-bit multiply-add in the specialised kernels; a leaf matrix whose leading dimension
    reaches 4 GiB takes the 64-bit variant (the bulk form's producer always forms 64-bit addresses; its consumers' cross
    stores switch).  Only the first `batch` columns of each row are touched."""
    roots = graphgen.random_dag(91, n_leaves=3, n_inner=12, n_roots=2)
    raw, _ = fd.flatten(roots)
    if bulk:
        monkeypatch.setenv("FDG_JIT_BULK", "1")
    ev = fd.compile_raw(raw, backend=JIT, jit_segment=6 if bulk else 0)
    if bulk:
        ev.set_launch(0, 1, 0)
    batch, ld = 2050, (1 << 29) + 16
    host = graphgen.leaf_values(3, ev.n_leaves, batch, signed=True)
    leaf = torch.empty(ev.n_leaves, ld, dtype=torch.float64, device="cuda")
    leaf[:, :batch] = torch.from_numpy(host).cuda()
    root = torch.zeros(ev.n_roots, batch, dtype=torch.float64, device="cuda")
    ev.eval_device(leaf.data_ptr(), ld, root.data_ptr(), batch, batch, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert root.cpu().numpy().tobytes() == O.Oracle(raw).eval(host).tobytes()
    if bulk:
        last = ev.jit_last()
        assert last["bulk"] and last["kernels"] >= 2 and last["cross_rows"] > 0


@pytest.mark.parametrize("name,dtype,log2_batch", [("parquet_sigma_o3", np.float64, 24), ("parquet_ver4_o4", np.float64, 20),
                                                   ("gv_ver4_o4", np.float64, 19), ("taylor_sigma_o3", np.complex128, 22)])
def test_full_size_checksums_on_the_baseline_graphs(name, dtype, log2_batch):
    """BASELINE-size batches (cfg 2 at its full 2^24 samples; the order-4 graphs at one resident batch) through
    size-independent properties: with every leaf equal to one each root is the signed diagram count of
    workloads/MANIFEST.json at EVERY sample and the accumulators are count x batch exactly (integers below 2^53);
    with random leaves a strided subsample is bit-equal to the oracle and accumulate equals the sum of eval."""
    import json
    import os

    root_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = fd.RawGraph.load(os.path.join(root_dir, "workloads", name + ".npz"))
    want = np.array(json.load(open(os.path.join(root_dir, "workloads", "MANIFEST.json")))[name]["all_leaves_one"])
    ev = fd.compile_raw(raw, dtype=dtype)
    B = 1 << log2_batch
    tdt = torch.float64 if dtype == np.float64 else torch.complex128
    W = 1 if dtype == np.float64 else 2
    s = torch.cuda.current_stream().cuda_stream
    leaf = torch.ones(ev.n_leaves, B, dtype=tdt, device="cuda")
    acc = torch.zeros(ev.n_roots * W, dtype=torch.float64, device="cuda")
    ev.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), s)
    torch.cuda.synchronize()
    got = acc.cpu().numpy().reshape(ev.n_roots, W)
    assert np.array_equal(got[:, 0], want * B)
    if W == 2:
        assert not got[:, 1].any()
    # eval mode on a slice of the same batch: the count at every sample
    nb = min(B, 1 << 18)
    root = torch.zeros(ev.n_roots, nb, dtype=tdt, device="cuda")
    ev.eval_device(leaf.data_ptr(), B, root.data_ptr(), nb, nb, s)
    torch.cuda.synchronize()
    assert torch.equal(root, torch.from_numpy(want.astype(dtype)).cuda()[:, None].expand(ev.n_roots, nb))
    # random leaves: strided subsample against the oracle, accumulate against the sum of eval
    g = torch.Generator(device="cuda").manual_seed(99)
    torch.view_as_real(leaf).copy_(torch.rand(ev.n_leaves, B, 2, dtype=torch.float64, device="cuda", generator=g) + 0.5) if W == 2 else \
        leaf.copy_(torch.rand(ev.n_leaves, B, dtype=torch.float64, device="cuda", generator=g) + 0.5)
    ev.eval_device(leaf.data_ptr(), B, root.data_ptr(), nb, nb, s)
    acc.zero_()
    ev.accumulate_device(leaf.data_ptr(), B, nb, acc.data_ptr(), s)
    torch.cuda.synchronize()
    idx = torch.arange(0, nb, 1021, device="cuda")
    sub = np.ascontiguousarray(leaf[:, idx].cpu().numpy())
    assert root[:, idx].cpu().numpy().tobytes() == O.Oracle(raw).eval(sub).tobytes()
    ref = torch.view_as_real(root).sum(dim=1).reshape(-1) if W == 2 else root.sum(dim=1)
    scale = (torch.view_as_real(root).abs().sum(dim=1).reshape(-1) if W == 2 else root.abs().sum(dim=1)) + 1e-300
    assert bool(((acc - ref).abs() <= 1e-12 * scale).all())


@pytest.mark.parametrize("backend", BACKENDS)
def test_overflow_infinities_signed_zeros_and_nans(backend):
    """Values that overflow, cancel or vanish: +-inf and +-0 must match the oracle bit for bit (x * (-1.0) is written
    as a negation in the specialised kernels: same bits for every finite and infinite double); where the result is NaN
    both sides must say NaN -- its sign and payload are hardware conventions (x86 and NVIDIA differ) and are not compared."""
    roots = graphgen.random_dag(123, n_leaves=10, n_inner=120, n_roots=6, p_power=0.15, max_pow=5)
    raw, _ = fd.flatten(roots)
    ev = fd.compile_raw(raw, backend=backend, jit_segment=60 if backend == JIT else 0)
    batch = 2048
    rng = np.random.default_rng(17)
    leaf = graphgen.leaf_values(9, ev.n_leaves, batch, signed=True)
    kind = rng.integers(0, 6, size=leaf.shape)
    leaf = np.where(kind == 0, leaf * 1e200, leaf)       # overflow in products
    leaf = np.where(kind == 1, leaf * 1e-200, leaf)      # underflow to subnormals / zero
    leaf = np.where(kind == 2, 0.0 * np.sign(leaf), leaf)  # +-0
    leaf = np.ascontiguousarray(leaf)
    with np.errstate(all="ignore"):
        want = O.Oracle(raw).eval(leaf)
    got = _dev_eval(ev, leaf, batch)
    nan_w, nan_g = np.isnan(want), np.isnan(got)
    assert np.array_equal(nan_w, nan_g)
    assert nan_w.any() and np.isinf(want).any() and (want == 0).any()  # the case really exercises all three
    assert np.array_equal(want.view(np.uint64)[~nan_w], got.view(np.uint64)[~nan_w])


def test_host_arrays_in_many_chunks_on_a_multi_kernel_program():
    """fdg_eval_host pipelines 64 MiB chunks over two streams; with a multi-kernel program both streams run launch
    sequences that need a cross buffer -- each stream has its own.  Bit-exact against the oracle on every sample."""
    import os

    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", "parquet_ver4_o3.npz"))
    ev = fd.compile_raw(raw, backend=JIT, jit_segment=700)
    assert ev.jit_prepare(1, False)["kernels"] >= 4 and ev.jit_prepare(1, False)["cross_rows"] > 0
    B = 150000  # 4 chunks of 46080 samples
    leafT = graphgen.leaf_values(11, ev.n_leaves, B, signed=True)      # (L, B)
    leafVal = np.asfortranarray(leafT.T)                                # (B, L), batch unit-stride
    root = np.asfortranarray(np.zeros((B, ev.n_roots)))
    for _ in range(3):  # repeated: a race would not show every time
        root[...] = 0.0
        ev(root, leafVal)
        want = O.Oracle(raw).eval(leafT, nthreads=8)
        assert np.ascontiguousarray(root.T).tobytes() == want.tobytes()


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_launch_sequences_over_sub_batches(dtype, monkeypatch):
    """A batch larger than the cross buffer allows is processed as several launch sequences (FDG_JIT_CROSS_GB caps the
    buffer; the floor is 4 blocks per SM = 75 776 samples): eval and accumulate must not notice."""
    import os

    monkeypatch.setenv("FDG_JIT_CROSS_GB", "0.0001")
    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", "parquet_ver4_o3.npz"))
    ev = fd.compile_raw(raw, dtype=dtype, backend=JIT, jit_segment=900)
    B = 3 * 75776 + 1234
    tdt = torch.float64 if dtype == np.float64 else torch.complex128
    W = 1 if dtype == np.float64 else 2
    g = torch.Generator(device="cuda").manual_seed(5)
    leaf = torch.empty(ev.n_leaves, B, dtype=tdt, device="cuda")
    torch.view_as_real(leaf).copy_(torch.rand(ev.n_leaves, B, 2, dtype=torch.float64, device="cuda", generator=g) - 0.5) if W == 2 else \
        leaf.copy_(torch.rand(ev.n_leaves, B, dtype=torch.float64, device="cuda", generator=g) - 0.5)
    root = torch.zeros(ev.n_roots, B, dtype=tdt, device="cuda")
    acc = torch.zeros(ev.n_roots * W, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ev.eval_device(leaf.data_ptr(), B, root.data_ptr(), B, B, s)
    ev.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), s)
    torch.cuda.synchronize()
    idx = torch.cat([torch.arange(0, B, 997, device="cuda"), torch.tensor([75775, 75776, 151551, 151552, B - 1], device="cuda")])
    sub = np.ascontiguousarray(leaf[:, idx].cpu().numpy())
    assert root[:, idx].cpu().numpy().tobytes() == O.Oracle(raw).eval(sub).tobytes()
    ref = torch.view_as_real(root).sum(dim=1).reshape(-1) if W == 2 else root.sum(dim=1)
    scale = (torch.view_as_real(root).abs().sum(dim=1).reshape(-1) if W == 2 else root.abs().sum(dim=1)) + 1e-300
    assert bool(((acc - ref).abs() <= 1e-12 * scale).all())


@pytest.mark.parametrize("name,dtype", [("parquet_ver4_o3", np.float64), ("taylor_sigma_o3", np.complex128), ("gv_sigma_o5", np.float64)])
def test_opt_in_fma_is_within_the_tolerance_but_not_the_default(name, dtype):
    """fdg_options.fma = 1 lets the specialised kernels fuse multiplies into adds.  It is an opt-in: not bit-identical
    (one rounding fewer per fused pair), error bounded on the scale of the sum of absolute terms well inside the 1e-12
    of the north star; the default stays bit-exact."""
    import os

    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", name + ".npz"))
    orc = O.Oracle(raw)
    batch = 512
    leaf = graphgen.leaf_values(21, orc.n_leaves, batch, dtype=dtype, signed=True)
    want = orc.eval(leaf)
    exact = _dev_eval(fd.compile_raw(raw, dtype=dtype, backend=JIT), leaf, batch)
    assert exact.tobytes() == want.tobytes()
    fused = _dev_eval(fd.compile_raw(raw, dtype=dtype, backend=JIT, fma=True), leaf, batch)
    assert fused.tobytes() != want.tobytes()                                   # contraction really happened
    for b in range(0, batch, 97):
        bound = np.array(O.eval_abs_bound(orc, np.abs(leaf[:, b]) if dtype == np.float64 else np.abs(leaf[:, b].real) + np.abs(leaf[:, b].imag)))
        assert (np.abs(fused[:, b] - want[:, b]) <= 1e-12 * bound).all()
    with pytest.raises(_capi.FdgError):
        fd.compile_raw(raw, dtype=dtype, backend=VM, fma=True)                  # the packet VM has the exact arithmetic only


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_grid_stride_kernel_over_several_launch_sequences(dtype, monkeypatch):
    """A single small accumulate kernel runs as a grid-stride loop whose warps STORE their sums.  When the batch is
    processed as several launch sequences (FDG_JIT_MAX_SUB forces that here; huge batches do it on their own) every
    sequence must be folded into the result before the next one overwrites the rows -- also when the last sequence runs
    a smaller grid than the others."""
    monkeypatch.setenv("FDG_JIT_MAX_SUB", "40000")
    roots = graphgen.random_dag(11, n_leaves=9, n_inner=40, n_roots=3)
    raw, _ = fd.flatten(roots)
    ev = fd.compile_raw(raw, dtype=dtype, backend=JIT)
    W = 1 if dtype == np.float64 else 2
    info = ev.jit_prepare(1 if W == 2 else 2, True)
    assert info["grid_stride"] and info["kernels"] == 1
    B = 3 * 40000 + 777  # three full launch sequences and a short one
    leaf_h = graphgen.leaf_values(3, ev.n_leaves, B, dtype=dtype, signed=True)
    leaf = torch.from_numpy(leaf_h).cuda()
    acc = torch.zeros(ev.n_roots * W, dtype=torch.float64, device="cuda")
    ev.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = O.Oracle(raw).eval(leaf_h)
    ref = want.sum(axis=1)
    ref = np.stack([ref.real, ref.imag], axis=1).reshape(-1) if W == 2 else ref
    scale = np.abs(want).sum(axis=1)
    scale = np.repeat(scale, 2) if W == 2 else scale
    assert np.all(np.abs(acc.cpu().numpy() - ref) <= 1e-12 * (scale + 1e-300))


@pytest.mark.parametrize("name,dtype", [("parquet_ver4_o3", np.float64), ("taylor_sigma_o4", np.complex128)])
def test_pipeline_form_is_bit_exact(name, dtype, monkeypatch):
    """The pipeline form of the specialised kernels (one cooperative kernel per pass, one code stage per instruction-cache
    group, tiles of 32 samples flowing between the stages; opt-in with FDG_JIT_PIPE=1): same bits as the oracle in eval
    mode, same sums in accumulate mode, no stage ever gave up waiting."""
    import os

    monkeypatch.setenv("FDG_JIT_PIPE", "1")
    monkeypatch.setenv("FDG_PIPE_MIN_BATCH", "1")
    raw = fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", name + ".npz"))
    ev = fd.compile_raw(raw, dtype=dtype, backend=JIT, jit_segment=1200)
    W = 1 if dtype == np.float64 else 2
    tdt = torch.float64 if W == 1 else torch.complex128
    B = 70001  # more tiles than the window of tiles in flight, and a ragged last tile
    leaf_h = graphgen.leaf_values(9, ev.n_leaves, B, dtype=dtype, signed=True)
    leaf = torch.from_numpy(leaf_h).cuda()
    root = torch.full((ev.n_roots, B), -3.0, dtype=tdt, device="cuda")
    acc = torch.zeros(ev.n_roots * W, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    launches = ev.launches
    ev.eval_device(leaf.data_ptr(), B, root.data_ptr(), B, B, s)
    n_eval = ev.pipeline_prepare(False, torch.cuda.get_device_properties(0).multi_processor_count)["stages"]
    assert not ev.pipeline_stats(s, n_eval)["stalled"]
    ev.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), s)
    torch.cuda.synchronize()
    assert ev.launches > launches
    idx = np.unique(np.concatenate([np.arange(0, B, 53), [31, 32, B - 2, B - 1]]))
    want = O.Oracle(raw).eval(np.ascontiguousarray(leaf_h[:, idx]))
    assert root.cpu().numpy()[:, idx].tobytes() == want.tobytes()
    got = root.cpu().numpy()
    ref = got.sum(axis=1)
    ref = np.stack([ref.real, ref.imag], axis=1).reshape(-1) if W == 2 else ref
    scale = np.abs(got).sum(axis=1)
    scale = np.repeat(scale, 2) if W == 2 else scale
    assert np.all(np.abs(acc.cpu().numpy() - ref) <= 1e-12 * (scale + 1e-300))


def test_two_host_callers_on_one_handle():
    """include/fdgraph.h: a handle may be used from several threads.  Two threads push different host batches through the
    SAME handle's host entry point (fdg_eval_host: chunked H2D -> kernels -> D2H on the handle's two streams) at the same
    time, several times over; each must get exactly the oracle's bytes for its own batch."""
    import threading

    roots = graphgen.random_dag(17, n_leaves=14, n_inner=120, n_roots=4)
    ev, _ = fd.compile(roots)
    orc = O.Oracle(ev.raw)
    B = 300_000  # several 64 MiB chunks per call would need more leaves; this is two chunks of ~150 k samples... per caller
    leaves = [np.asfortranarray(graphgen.leaf_values(100 + t, ev.n_leaves, B, signed=True).T) for t in range(2)]
    wants = [orc.eval(np.ascontiguousarray(lv.T)) for lv in leaves]
    errors = []

    def work(t):
        try:
            for _ in range(4):
                root = np.asfortranarray(np.zeros((B, ev.n_roots)))
                ev(root, leaves[t])
                if np.ascontiguousarray(root.T).tobytes() != wants[t].tobytes():
                    errors.append(f"thread {t}: result differs from the oracle")
        except Exception as ex:  # noqa: BLE001
            errors.append(f"thread {t}: {ex!r}")

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    assert ev.launches >= 8


def _workload(name):
    import os

    return fd.RawGraph.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "workloads", name + ".npz"))


@pytest.mark.parametrize("name,dtype", [("parquet_ver4_o3", np.float64), ("parquet_ver4_o4", np.float64), ("gv_ver4_o4", np.float64),
                                        ("taylor_sigma_o3", np.complex128), ("taylor_sigma_o4", np.complex128), ("parquet_sigma_o4", np.float64)])
@pytest.mark.parametrize("cse", [True, None])
def test_common_subexpressions_merged_same_bits(name, dtype, cse):
    """SURVEY section 8f N4 (optimize.jl:345-390): merging equal sub-expressions keeps every bit (operand order is part of
    the key).  cse=True forces it, cse=None lets the planner's model decide; both against the oracle of the UNMERGED graph,
    eval mode bit for bit and accumulate against the sum."""
    raw = _workload(name)
    ev = fd.compile_raw(raw, dtype=dtype, backend=JIT, cse=cse)
    W = 1 if dtype == np.float64 else 2
    info = ev.jit_prepare(1 if (W == 2 or ev.stats["n_operands"] > 400) else 2, False)
    if cse:
        assert info["cse"]
    B = 4096
    leaf_h = graphgen.leaf_values(21, ev.n_leaves, B, dtype=dtype, signed=True)
    got = _dev_eval(ev, leaf_h, B)
    want = O.Oracle(raw).eval(leaf_h)
    assert got.tobytes() == want.tobytes()
    leaf = torch.from_numpy(leaf_h).cuda()
    acc = torch.zeros(ev.n_roots * W, dtype=torch.float64, device="cuda")
    ev.accumulate_device(leaf.data_ptr(), B, B, acc.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = want.sum(axis=1)
    ref = np.stack([ref.real, ref.imag], axis=1).reshape(-1) if W == 2 else ref
    scale = np.abs(want).sum(axis=1)
    scale = np.repeat(scale, 2) if W == 2 else scale
    assert np.all(np.abs(acc.cpu().numpy() - ref) <= 1e-12 * (scale + 1e-300))


def test_automatic_cse_follows_the_model():
    """The planner keeps the merged program where its modelled time (traffic, FP64 instructions, registers spilled) is at
    least 5 % lower -- measured on this hardware: Parquet sigma order 4 +9 %, Taylor-AD sigma order 4 +10 % -- and the
    unmerged one elsewhere: the memory-bound order-4 vertices (merging makes more values cross kernels) and Taylor-AD
    sigma order 3, where the shared values push ptxas into spilling (-9 % measured)."""
    picks = {}
    for name, dtype in (("taylor_sigma_o4", np.complex128), ("parquet_sigma_o4", np.float64), ("taylor_sigma_o3", np.complex128),
                        ("parquet_ver4_o4", np.float64), ("gv_ver4_o4", np.float64)):
        ev = fd.compile_raw(_workload(name), dtype=dtype, backend=JIT)
        picks[name] = ev.jit_prepare(1, True)["cse"]
    # (Parquet vertex4 order 4: merged only within the reach of one kernel -- "scoped" -- which saves a third of the
    # arithmetic without more values crossing kernels)
    assert picks == {"taylor_sigma_o4": True, "parquet_sigma_o4": True, "taylor_sigma_o3": False, "parquet_ver4_o4": True, "gv_ver4_o4": False}


@pytest.mark.parametrize("name", ["parquet_ver4_o4", "gv_ver4_o4"])
@pytest.mark.parametrize("backend", BACKENDS)
def test_order_four_vertices_against_the_full_oracle(name, backend):
    """The headline graph (Parquet vertex4 order 4) and GV vertex4 order 4 at 4096 samples, EVERY sample against the
    oracle, on both back ends."""
    raw = _workload(name)
    ev = fd.compile_raw(raw, backend=backend)
    B = 4096
    leaf = graphgen.leaf_values(31, ev.n_leaves, B, signed=True)
    got = _dev_eval(ev, leaf, B)
    assert got.tobytes() == O.Oracle(raw).eval(leaf).tobytes()


def test_opt_in_fma_on_the_headline_graph():
    """fdg_options.fma = 1 on Parquet vertex4 order 4: not the reference's bits, but within 1e-12 of them on the scale of
    the sum of the absolute terms (leaves in [0.5, 1.5): no cancellation to amplify the single roundings saved)."""
    raw = _workload("parquet_ver4_o4")
    ev = fd.compile_raw(raw, backend=JIT, fma=True)
    B = 2048
    leaf = graphgen.leaf_values(41, ev.n_leaves, B, signed=False)
    got = _dev_eval(ev, leaf, B)
    want = O.Oracle(raw).eval(leaf)
    assert got.tobytes() != want.tobytes()
    assert np.all(np.abs(got - want) <= 1e-12 * np.abs(want).max(axis=1, keepdims=True))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_graph_file_compiled_by_the_library_evaluates_to_the_oracle(tmp_path, dtype):
    """a graph flattened elsewhere travels as an FDGRAPH file (RawGraph.save_fdg here,
    FDGraphB200.save_graph in Julia), the library reads the file itself (fdg_compile_file) and the program it builds
    evaluates on the device to the oracle's bytes -- eval mode and the host entry point."""
    raw = _workload("parquet_ver4_o3")
    path = str(tmp_path / "g.fdgraph")
    raw.save_fdg(path)
    ev = fd.compile_file(path, dtype=dtype)
    assert ev.n_leaves == 175 and ev.n_roots == 84
    B = 3000
    leaf = graphgen.leaf_values(12, ev.n_leaves, B, dtype=dtype, signed=True)
    want = O.Oracle(raw).eval(leaf)
    assert _dev_eval(ev, leaf, B).tobytes() == want.tobytes()
    root = np.asfortranarray(np.zeros((B, ev.n_roots), dtype))
    ev(root, np.asfortranarray(leaf.T))
    assert np.ascontiguousarray(root.T).tobytes() == want.tobytes()


# ---- bulk form: persistent warp-specialised kernels, rows fetched by cp.async.bulk into a ring (DESIGN.md 4b') ----------
@pytest.mark.parametrize("name,dtype,seg", [("parquet_ver4_o3", np.float64, 700), ("parquet_ver4_o3", np.float64, 0), ("taylor_sigma_o3", np.complex128, 900),
                                            ("gv_sigma_o5", np.float64, 500)])
@pytest.mark.parametrize("batch,ld", [(256, 256), (1000, 1000), (1001, 1002), (4099, 4100), (77777, 77778)])
def test_bulk_form_is_bit_exact(name, dtype, seg, batch, ld, monkeypatch):
    """FDG_JIT_BULK=1 forces the bulk form whatever the batch: whole and ragged tiles (the copy of a ragged tile is rounded
    up to 16 bytes, inside the row), fewer tiles than SMs and more, real and complex; every sample against the oracle."""
    monkeypatch.setenv("FDG_JIT_BULK", "1")
    raw = _workload(name)
    ev = fd.compile_raw(raw, dtype=dtype, backend=JIT, jit_segment=seg)
    leaf = graphgen.leaf_values(41, ev.n_leaves, batch, dtype=dtype, signed=True, ld=ld)
    got = _dev_eval(ev, leaf, batch, spt=1)
    last = ev.jit_last()
    assert last["bulk"] and last["bulk_smem"] > 64 * 1024 and not last["grid_stride"]
    want = O.Oracle(raw).eval(np.ascontiguousarray(leaf[:, :batch]), "emitter", root=np.full((ev.n_roots, batch), -3.0, dtype))
    assert got.tobytes() == want.tobytes()
    ptx, log = ev.jit_ptx(0, False, 0)
    assert "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes" in ptx and "setmaxnreg.inc.sync.aligned.u32 240" in ptx
    assert "mbarrier.try_wait.parity" in ptx and "cp.async.ca" not in ptx


@pytest.mark.parametrize("name,dtype", [("parquet_ver4_o3", np.float64), ("taylor_sigma_o3", np.complex128)])
def test_bulk_form_accumulates_over_launch_sequences(name, dtype, monkeypatch):
    """Accumulate mode of the bulk form: per-thread running sums in shared memory, added into the warp's partial row at
    the end of the kernel -- so several launch sequences (FDG_JIT_MAX_SUB) add up; deterministic from run to run."""
    monkeypatch.setenv("FDG_JIT_BULK", "1")
    raw = _workload(name)
    W = 2 if dtype == np.complex128 else 1
    ev = fd.compile_raw(raw, dtype=dtype, backend=JIT, jit_segment=800)
    ev.set_launch(0, 1, 0)
    B = 50001
    leaf = graphgen.leaf_values(43, ev.n_leaves, B, dtype=dtype, signed=True, ld=B + 1)
    want = O.Oracle(raw).eval(np.ascontiguousarray(leaf[:, :B]))
    ref = want.sum(axis=1)
    scale = np.abs(want).sum(axis=1)
    dleaf = torch.from_numpy(leaf).cuda()
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    for max_sub in (0, 0, 4096):
        if max_sub:
            monkeypatch.setenv("FDG_JIT_MAX_SUB", str(max_sub))
        acc = torch.zeros(ev.n_roots * W, dtype=torch.float64, device="cuda")
        ev.accumulate_device(dleaf.data_ptr(), B + 1, B, acc.data_ptr(), s)
        torch.cuda.synchronize()
        assert ev.jit_last()["bulk"]
        a = acc.cpu().numpy()
        outs.append(a)
        got = a.view(np.complex128) if W == 2 else a
        assert np.all(np.abs(got - ref) <= 1e-12 * (scale + 1e-300))
    assert outs[0].tobytes() == outs[1].tobytes()  # same launch shape, same bits


def test_bulk_form_is_chosen_for_big_batches_and_needs_aligned_rows(monkeypatch):
    """Without the override the bulk form runs where it pays -- programs of several kernels on batches that give every SM a
    few tiles -- and never on rows a bulk copy cannot address (16-byte alignment): those take the ring form.  Same bits."""
    raw = _workload("parquet_ver4_o3")
    ev = fd.compile_raw(raw, backend=JIT, jit_segment=700)
    ev.set_launch(0, 1, 0)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    B = 256 * sm * 8
    leaf = graphgen.leaf_values(47, ev.n_leaves, B + 2, signed=True)          # (L, B + 2), rows 16-byte aligned
    dleaf = torch.from_numpy(leaf).cuda()
    s = torch.cuda.current_stream().cuda_stream
    root = torch.empty(ev.n_roots, B, dtype=torch.float64, device="cuda")
    ev.eval_device(dleaf.data_ptr(), B + 2, root.data_ptr(), B, B, s)
    torch.cuda.synchronize()
    assert ev.jit_last()["bulk"]
    big = root.cpu().numpy()
    ev.eval_device(dleaf.data_ptr(), B + 2, root.data_ptr(), B, 4096, s)            # a small batch: ring form
    torch.cuda.synchronize()
    assert not ev.jit_last()["bulk"]
    assert root.cpu().numpy()[:, :4096].tobytes() == big[:, :4096].tobytes()
    root2 = torch.empty(ev.n_roots, B, dtype=torch.float64, device="cuda")
    ev.eval_device(dleaf.data_ptr() + 8, B + 2, root2.data_ptr(), B, B, s)          # rows start 8 bytes off a 16-byte boundary
    torch.cuda.synchronize()
    assert not ev.jit_last()["bulk"]
    want = O.Oracle(raw).eval(np.ascontiguousarray(leaf[:, 1:4097]))
    assert root2.cpu().numpy()[:, :4096].tobytes() == want.tobytes()
    assert big[:, 1:4096].tobytes() == root2.cpu().numpy()[:, :4095].tobytes()
    monkeypatch.setenv("FDG_JIT_BULK", "0")
    ev.eval_device(dleaf.data_ptr(), B + 2, root2.data_ptr(), B, B, s)
    torch.cuda.synchronize()
    assert not ev.jit_last()["bulk"] and root2.cpu().numpy().tobytes() == big.tobytes()


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("loops,orders", [(2, (2, 1)), (3, (1, 1))])
def test_derivative_graphs_of_a_self_energy(loops, orders, backend):
    """Another producer of evaluator inputs: the reference's graph-level AD (build_derivative_graph,
    operation.jl:478-543, restated in oracle/frontend/ad.py and pinned on the reference's known answers in the CPU suite)
    applied to the Parquet self-energy; every derivative graph of every root is a root of the compiled program, the dual
    leaves are ordinary leaves.  Device bytes == oracle bytes."""
    from oracle.frontend import ad, parquet as pq

    fd.uidreset()
    pq._ver4I.clear()
    graphs = [r["diagram"] for r in pq.sigma(pq.DiagPara(type=pq.SigmaDiag, innerLoopNum=loops))]
    dual = ad.build_derivative_graph(graphs, orders)
    root_ids = {g.id for g in graphs}
    roots = list(graphs) + [d for (nid, _), d in sorted(dual.items(), key=lambda kv: (kv[0][0], kv[0][1])) if nid in root_ids]
    assert len(roots) == len(graphs) * (1 + sum(orders))
    _parity(roots, batch=3000, backend=backend)
