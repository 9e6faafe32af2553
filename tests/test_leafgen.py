"""N1 -- leaf values computed on the device from the Monte-Carlo variables (fdg_leafgen_*).

CPU part: the restated integrand (oracle/leafgen.py) against a scalar transcription of example/benchmark.jl:44-127 and
the restated `leafstates` against the committed sidecars.  GPU part: device leaf values against the CPU restatement
within 2e-14 relative (exp() is the device's, <= 1 ulp; the reference's own `mul!` is a BLAS call whose summation order
is not specified either), and the graph evaluation on top of the device's leaves bit-exact against the oracle."""
import math
import os

import numpy as np
import pytest

import fdgraph_b200 as fd
from fdgraph_b200 import _capi
from oracle import leafgen
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KF, BETA, LAM = 1.919, 3.0, 1.2  # example/benchmark.jl:11-18


def _load(name):
    raw = fd.RawGraph.load(os.path.join(ROOT, "workloads", name + ".npz"))
    meta = dict(np.load(os.path.join(ROOT, "workloads", name + ".leaves.npz")))
    return raw, meta


def _variables(meta, B, seed=1234, dim=3):
    """varK / varT of example/benchmark.jl:45-53: K1 = K2 = (kF, 0, 0), K3 = 0, inner loops random; T[0] = 0."""
    rng = np.random.default_rng(seed)
    n_loops = meta["loop_basis"].shape[1]
    n_tau = int(max(meta["tau_in"].max(), meta["tau_out"].max())) + 1
    K = rng.random((dim, n_loops, B)) * 2 - 0.7
    K[:, 0, :] = K[:, 1, :] = np.array([KF, 0.0, 0.0])[:, None]
    if n_loops > 2:
        K[:, 2, :] = 0.0
    T = rng.random((n_tau, B)) * BETA
    T[0] = 0.0
    T[-1, ::7] = T[0, ::7]  # some exactly equal times: the tau == 0 branch (TAU_CUTOFF)
    return K, T


def _green_scalar(tau, w, beta):
    if tau == 0.0:
        tau = -1e-10
    if tau > 0.0:
        return math.exp(-w * tau) / (1 + math.exp(-w * beta)) if w > 0.0 else math.exp(w * (beta - tau)) / (1 + math.exp(w * beta))
    return -math.exp(-w * (tau + beta)) / (1 + math.exp(-w * beta)) if w > 0.0 else -math.exp(-w * tau) / (1 + math.exp(w * beta))


def test_cpu_restatement_against_a_scalar_transcription():
    raw, meta = _load("parquet_sigma_o3")
    K, T = _variables(meta, 5)
    got = leafgen.leaf_values(meta, K, T, KF, BETA, LAM)
    for b in range(5):
        for l in range(len(meta["leaf_type"])):
            basis = meta["loop_basis"][meta["loop_index"][l]]
            kq = [sum(K[c, j, b] * basis[j] for j in range(K.shape[1])) for c in range(3)]
            q2 = sum(x * x for x in kq)
            if meta["leaf_type"][l] == 1:
                assert meta["leaf_order"][l][0] == 0
                want = _green_scalar(T[meta["tau_out"][l], b] - T[meta["tau_in"][l], b], q2 - KF * KF, BETA)
            else:
                inv = 1.0 / (q2 + LAM)
                want = 8 * math.pi / inv * (LAM * inv) ** int(meta["leaf_order"][l][1])
            assert got[l, b] == pytest.approx(want, rel=1e-15)
    # fermionic sign structure: G(tau -> 0-) = -(1 - n_F), G(tau > 0) > 0
    assert (leafgen.green(np.array([0.0]), np.array([0.5]), BETA) < 0).all()
    assert (leafgen.green(np.array([0.3]), np.array([-0.5]), BETA) > 0).all()


def test_green_derivatives_against_high_precision_differentiation():
    """green_derive (benchmark.jl:93-111), orders 1..5: the closed form of oracle/leafgen.py (and of the device kernel)
    against 50-digit numerical differentiation of `green` itself.  The reference's own numbers come from
    Lehmann.Spectral (not vendored): this pins the FUNCTION, not that dependency's bits."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50

    def green_mp(tau, w, beta):
        tau = mp.mpf(tau)
        if tau == 0:
            tau = mp.mpf("-1e-10")
        if tau > 0:
            return mp.e ** (-w * tau) / (1 + mp.e ** (-w * beta))
        return -mp.e ** (-w * (tau + beta)) / (1 + mp.e ** (-w * beta))

    rng = np.random.default_rng(0)
    for trial in range(40):
        tau = 0.0 if trial % 10 == 0 else float(rng.uniform(-BETA, BETA))
        w = float(rng.uniform(-6, 6))
        for order in range(1, 6):
            d = mp.diff(lambda x: green_mp(tau, x, BETA), mp.mpf(w), order)
            want = (-1) ** order * d / math.factorial(order)
            got = leafgen.green_derive(np.array([tau]), np.array([w]), BETA, order)[0]
            scale = abs(green_mp(tau, mp.mpf(w), BETA)) * BETA ** order
            assert abs(mp.mpf(float(got)) - want) <= 1e-14 * scale
    with pytest.raises(NotImplementedError):
        leafgen.green_derive(np.array([0.1]), np.array([0.2]), BETA, 6)


def test_sidecars_match_the_workloads():
    for name in ("parquet_sigma_o3", "parquet_ver4_o3", "parquet_ver4_o4", "gv_sigma_o4", "gv_ver4_o3", "taylor_sigma_o2", "taylor_sigma_o3"):
        raw, meta = _load(name)
        L = O.Oracle(raw).n_leaves
        assert len(meta["leaf_type"]) == L == len(meta["tau_in"]) == len(meta["loop_index"])
        assert set(np.unique(meta["leaf_type"])) <= {1, 2} and ((meta["leaf_order"] == 0).all() or name.startswith("taylor"))
        assert meta["loop_index"].max() == meta["loop_basis"].shape[0] - 1           # every basis vector is used
        assert len({tuple(b) for b in meta["loop_basis"]}) == meta["loop_basis"].shape[0]  # and distinct
    assert _load("parquet_ver4_o4")[1]["loop_basis"].shape[1] == 7                  # MaxLoopNum of example/benchmark.jl:21
    # the Taylor-AD graphs: every propagator / interaction leaf also appears with its counter-term orders (g, v)
    _, meta = _load("taylor_sigma_o3")
    orders = {tuple(o) for o in meta["leaf_order"].tolist()}
    assert orders == {(0, 0), (1, 0), (2, 0), (0, 1)}
    assert all(o[1] == 0 for o, t in zip(meta["leaf_order"], meta["leaf_type"]) if t == 1)


def test_create_rejects_what_the_reference_cannot_compute():
    _, meta = _load("parquet_sigma_o3")
    g = fd.LeafGenerator(meta, kF=KF, beta=BETA, lam=LAM)
    assert (g.n_leaves, g.n_loops, g.n_tau, g.var_rows) == (27, 4, 3, 15)
    bad = {k: v.copy() for k, v in meta.items()}
    bad["leaf_order"][np.argmax(meta["leaf_type"] == 1), 0] = 6  # "not implemented!" (benchmark.jl:107)
    with pytest.raises(_capi.FdgError) as e:
        fd.LeafGenerator(bad)
    assert e.value.code == 3
    bad = {k: v.copy() for k, v in meta.items()}
    bad["leaf_type"][0] = 3  # BareGreenNId: "this leaftype not implemented" (benchmark.jl:76)
    with pytest.raises(_capi.FdgError):
        fd.LeafGenerator(bad)
    bad = {k: v.copy() for k, v in meta.items()}
    bad["loop_index"][3] = 99
    with pytest.raises(_capi.FdgError) as e:
        fd.LeafGenerator(bad)
    assert e.value.code == 1


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["parquet_sigma_o3", "parquet_ver4_o3", "gv_ver4_o3", "taylor_sigma_o2", "taylor_sigma_o3"])
def test_device_leaves_and_the_graph_on_top_of_them(name):
    torch = pytest.importorskip("torch")
    raw, meta = _load(name)
    B = 4097
    K, T = _variables(meta, B)
    gen = fd.LeafGenerator(meta, kF=KF, beta=BETA, lam=LAM)
    ev = fd.compile_raw(raw)
    dK = torch.from_numpy(K.transpose(1, 0, 2).reshape(-1, B).copy()).cuda()  # row (j * dim + c)
    dT = torch.from_numpy(T).cuda()
    leaf = torch.zeros(gen.n_leaves, B, dtype=torch.float64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    gen.fill_device(dK.data_ptr(), dT.data_ptr(), B, B, leaf.data_ptr(), B, s)
    torch.cuda.synchronize()
    got = leaf.cpu().numpy()
    want = leafgen.leaf_values(meta, K, T, KF, BETA, LAM)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 2e-14 * np.abs(want).max()
    # element by element: relative for the plain leaves; the counter-term (derivative) leaves are differences of terms of
    # size |green| * beta^order, which is the scale their rounding error lives on
    meta0 = dict(meta, leaf_order=np.zeros_like(meta["leaf_order"]))
    scale = np.abs(leafgen.leaf_values(meta0, K, T, KF, BETA, LAM)) * (BETA ** meta["leaf_order"].sum(axis=1))[:, None]
    derived = (meta["leaf_order"].sum(axis=1) > 0)[:, None]
    tol = np.where(derived, 1e-13 * scale, 2e-14 * np.abs(want))
    assert (np.abs(got - want) <= tol + 1e-300).all()
    # evaluation on the device's own leaves is bit-exact against the oracle on the same leaves
    root = torch.zeros(ev.n_roots, B, dtype=torch.float64, device="cuda")
    ev.eval_device(leaf.data_ptr(), B, root.data_ptr(), B, B, s)
    acc = torch.zeros(ev.n_roots, dtype=torch.float64, device="cuda")
    gen.accumulate_device(ev, dK.data_ptr(), dT.data_ptr(), B, B, acc.data_ptr(), s)
    torch.cuda.synchronize()
    ref = O.Oracle(raw).eval(got)
    assert root.cpu().numpy().tobytes() == ref.tobytes()
    scale = np.abs(ref).sum(axis=1) + 1e-300
    assert (np.abs(acc.cpu().numpy() - ref.sum(axis=1)) <= 1e-12 * scale).all()
    # the host entry point: (K, T) in, R sums out
    acc_h = gen.accumulate_host(ev, dK.cpu().numpy(), T)
    assert (np.abs(acc_h - ref.sum(axis=1)) <= 1e-12 * scale).all()


@pytest.mark.gpu
def test_generated_sub_batches_and_ragged_sizes(monkeypatch):
    torch = pytest.importorskip("torch")
    raw, meta = _load("parquet_sigma_o3")
    gen = fd.LeafGenerator(meta, kF=KF, beta=BETA, lam=LAM)
    ev = fd.compile_raw(raw)
    monkeypatch.setenv("FDG_LEAFGEN_GB", "0.0001")  # sub-batches of 4096 samples
    B = 3 * 4096 + 37
    K, T = _variables(meta, B, seed=5)
    Kr = K.transpose(1, 0, 2).reshape(-1, B).copy()
    acc = gen.accumulate_host(ev, Kr, T)
    leaf = leafgen.leaf_values(meta, K, T, KF, BETA, LAM)
    ref = O.Oracle(raw).eval(leaf)
    assert (np.abs(acc - ref.sum(axis=1)) <= 1e-11 * (np.abs(ref).sum(axis=1) + 1e-300)).all()


def test_specialised_generator_assembles_without_a_gpu():
    """The kernels written for a graph's leaves (csrc/fdg_lgjit.cpp) assemble for sm_100a on the host, cover every order-0
    leaf, stay inside the instruction cache, and hold the graph's data as immediates (no table is read)."""
    for name, n_cov in (("parquet_ver4_o4", 984), ("parquet_sigma_o3", 27), ("taylor_sigma_o3", 27)):
        _, meta = _load(name)
        gen = fd.LeafGenerator(meta, kF=KF, beta=BETA, lam=LAM)
        info, ptx = gen.jit_prepare(False, 0)
        assert info["leaves_covered"] == n_cov and info["kernels"] >= 1 and info["max_code_bytes"] < 120 * 1024
        assert ".target sm_100a" in ptx and "rcp.approx.ftz.f64" in ptx and "copysign.f64" in ptx and "mad.wide.u32" in ptx
        assert "ld.global.nc.f64 %fk" in ptx and "ld.global.nc.f64 %ft" in ptx  # the sample's variables, once, into registers
        wide, ptxw = gen.jit_prepare(True, 0)
        assert wide["kernels"] == info["kernels"] and "mad.lo.u64 %rd8, %rd7" in ptxw
    gen = fd.LeafGenerator(_load("parquet_ver4_o4")[1], kF=KF, beta=BETA, lam=LAM)
    assert gen.jit_prepare()["kernels"] >= 8  # ~40 momenta per kernel


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["parquet_ver4_o4", "gv_ver4_o3", "taylor_sigma_o3"])
def test_specialised_and_table_driven_generators_agree(name, monkeypatch):
    """FDG_LG_JIT=0 runs the table-driven kernel on every leaf; the default runs the kernels specialised for the graph
    (and the table-driven one on the counter-term leaves only).  Same formulas, same order of operations: the leaves agree
    to rounding (the specialised exp flushes results below 2^-1022 to zero), on whole and ragged blocks."""
    torch = pytest.importorskip("torch")
    raw, meta = _load(name)
    B = 5000 + 77
    K, T = _variables(meta, B, seed=11)
    gen = fd.LeafGenerator(meta, kF=KF, beta=BETA, lam=LAM)
    dK = torch.from_numpy(K.transpose(1, 0, 2).reshape(-1, B).copy()).cuda()
    dT = torch.from_numpy(T).cuda()
    s = torch.cuda.current_stream().cuda_stream
    out = []
    for jit in ("1", "0"):
        monkeypatch.setenv("FDG_LG_JIT", jit)
        leaf = torch.full((gen.n_leaves, B + 3), -7.0, dtype=torch.float64, device="cuda")
        gen.fill_device(dK.data_ptr(), dT.data_ptr(), B, B, leaf.data_ptr(), B + 3, s)
        torch.cuda.synchronize()
        got = leaf.cpu().numpy()
        assert (got[:, B:] == -7.0).all()  # nothing written past the batch
        out.append(got[:, :B])
    want = leafgen.leaf_values(meta, K, T, KF, BETA, LAM)
    derived = (meta["leaf_order"].sum(axis=1) > 0)[:, None]
    assert (out[0][derived[:, 0]] == out[1][derived[:, 0]]).all()          # counter-term leaves: the same kernel either way
    plain = ~derived[:, 0]
    assert (np.abs(out[0][plain] - out[1][plain]) <= 4e-16 * np.abs(out[1][plain]) + 1e-300).all()
    # against the numpy restatement: exp(a) carries the rounding of its argument, |a| eps relative, and |a| = |w| x is
    # |log| of the leaf (order-4 momenta reach |a| > 100, where one rounding of a dot product is already 2e-14 of the value)
    w = np.abs(want[plain])
    assert (np.abs(out[0][plain] - want[plain]) <= (2e-14 + 4e-16 * np.abs(np.log(w + 1e-300))) * w + 1e-300).all()
