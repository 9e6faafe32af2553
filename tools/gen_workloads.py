"""Generates the workload graphs under workloads/ from the reference's own data files.

Run in the build container (needs /root/reference for the GV `.diag` files); the GPU box only reads the
committed .npz files.  Each file is a flattened graph (the arrays of fdg_graph_desc, include/fdgraph.h) produced by
the restated front end + optimize! (oracle/frontend/), i.e. what `Compilers.compile(diags)` receives in
example/benchmark_GV.jl:23-39 / example/benchmark.jl:23-39.  workloads/MANIFEST.json records provenance, sizes and
the all-leaves-one value of every root (an exact integer-valued checksum of the diagram content).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import fdgraph_b200 as fd  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.frontend import gv, optimize as opt  # noqa: E402

OUT = os.path.join(ROOT, "workloads")


def emit(name, builder, source, manifest, optimize=True):
    t = time.time()
    fd.uidreset()
    graphs = builder()
    if optimize:
        opt.optimize(graphs)
    raw, _ = fd.flatten(graphs)
    raw.save(os.path.join(OUT, name + ".npz"))
    orc = O.Oracle(raw)
    ones = orc.eval(np.ones((max(orc.n_leaves, 1), 1)))[:, 0]
    ev = fd.compile_raw(raw)
    manifest[name] = {
        "source": source, "n_leaves": orc.n_leaves, "n_statements": orc.n_stmts, "n_roots": orc.n_roots,
        "n_nodes": raw.n_nodes, "n_edges": raw.n_edges, "flops_add": ev.stats["flops_add"], "flops_mul": ev.stats["flops_mul"],
        "all_leaves_one": [float(x) for x in ones],
    }
    print(f"{name}: L={orc.n_leaves} stmts={orc.n_stmts} R={orc.n_roots} edges={raw.n_edges}  ({time.time() - t:.1f}s)")


def taylor_workloads(manifest):
    """BASELINE config 5: README.md:59-86 -- Parquet self-energy, optimize!, then taylorAD with Green's-function
    counter-terms up to order 2 and interaction counter-terms up to order 1; every coefficient graph is a root."""
    from oracle.frontend import parquet as pq, taylor
    from oracle.frontend.ids import BareGreenId, BareInteractionId

    def build(order):
        pq._ver4I.clear()
        graphs = [r["diagram"] for r in pq.sigma(pq.DiagPara(type=pq.SigmaDiag, innerLoopNum=order))]
        opt.optimize(graphs)
        d = taylor.taylorAD(graphs, [2, 1], [lambda p: isinstance(p, BareGreenId), lambda p: isinstance(p, BareInteractionId)])
        return [g for k in sorted(d) for g in d[k]]

    for order in (2, 3, 4):
        emit(f"taylor_sigma_o{order}", lambda o=order: build(o),
             f"taylorAD(optimize!(Parquet.sigma(DiagPara(type=SigmaDiag, innerLoopNum={order}))), [2, 1], [BareGreenId, BareInteractionId]); "
             "roots = the coefficient graphs of orders (0,0) (0,1) (1,0) (1,1) (2,0) (2,1)  (README.md:59-86, src/utility.jl:48-93)",
             manifest, optimize=False)


def ver3_polar_workloads(manifest):
    """The other two Parquet front ends (vertex3.jl, polarization.jl) with their default parameters + optimize!, as
    parity cases of the evaluator: not BASELINE configurations, their all-leaves-one values are the diagram counts
    of test/front_end.jl:701-825 up to the spin / sign conventions written there."""
    from oracle.frontend import parquet as pq

    def ver3(order):
        pq._ver4I.clear()
        return [r["diagram"] for r in pq.vertex3(pq.DiagPara(type=pq.Ver3Diag, innerLoopNum=order))]

    def polar(order):
        pq._ver4I.clear()
        return [r["diagram"] for r in pq.polarization(pq.DiagPara(type=pq.PolarDiag, innerLoopNum=order))]

    for order in (2, 3, 4):
        emit(f"parquet_ver3_o{order}", lambda o=order: ver3(o),
             f"Parquet.vertex3(DiagPara(type=Ver3Diag, innerLoopNum={order})) + optimize!  (src/frontend/parquet/vertex3.jl:20-112)", manifest)
    for order in (3, 4, 5):
        emit(f"parquet_polar_o{order}", lambda o=order: polar(o),
             f"Parquet.polarization(DiagPara(type=PolarDiag, innerLoopNum={order})) + optimize!  (src/frontend/parquet/polarization.jl:16-127)", manifest)


def leaf_sidecars():
    """workloads/<name>.leaves.npz: the `leafstates` metadata (frontends.jl:175-232) of a workload's leaves, in leafVal
    order, for the on-device leaf generation (N1).  The graphs are rebuilt and must flatten to the committed arrays."""
    from oracle.frontend import leafstates as ls, parquet as pq

    def pq_sigma(order):
        pq._ver4I.clear()
        return [r["diagram"] for r in pq.sigma(pq.DiagPara(type=pq.SigmaDiag, innerLoopNum=order))]

    def pq_ver4(order):
        pq._ver4I.clear()
        return [r["diagram"] for r in pq.vertex4(pq.DiagPara(type=pq.Ver4Diag, innerLoopNum=order))]

    from oracle.frontend import taylor
    from oracle.frontend.ids import BareGreenId, BareInteractionId

    def taylor_sigma(order):  # as in taylor_workloads(): counter-term leaves carry derivative orders (g, v)
        graphs = pq_sigma(order)
        opt.optimize(graphs)
        d = taylor.taylorAD(graphs, [2, 1], [lambda p: isinstance(p, BareGreenId), lambda p: isinstance(p, BareInteractionId)])
        return [g for k in sorted(d) for g in d[k]]

    jobs = [("parquet_sigma_o3", lambda: pq_sigma(3), 4, True), ("parquet_ver4_o3", lambda: pq_ver4(3), 6, True),
            ("parquet_ver4_o4", lambda: pq_ver4(4), 7, True),  # example/benchmark.jl:13,21: MaxLoopNum = 7
            ("gv_sigma_o4", lambda: gv.diagsGV("sigma", 4), 5, True), ("gv_ver4_o3", lambda: gv.diagsGV_ver4(3), 6, True),
            ("taylor_sigma_o2", lambda: taylor_sigma(2), 3, False), ("taylor_sigma_o3", lambda: taylor_sigma(3), 4, False)]
    for name, builder, max_loops, optimize in jobs:
        fd.uidreset()
        graphs = builder()
        if optimize:
            opt.optimize(graphs)
        raw, nodes = fd.flatten(graphs)
        ref = fd.RawGraph.load(os.path.join(OUT, name + ".npz"))
        same = all(np.array_equal(getattr(raw, k), getattr(ref, k)) for k in raw.__dataclass_fields__)
        assert same, f"{name}: the rebuilt graph differs from the committed workload"
        orc = O.Oracle(raw)
        meta = ls.leafstates([nodes[i] for i in orc.leaf_nodes], max_loops)
        np.savez_compressed(os.path.join(OUT, name + ".leaves.npz"), **meta)
        print(f"{name}.leaves: L={len(meta['leaf_type'])} G={int((meta['leaf_type'] == 1).sum())} W={int((meta['leaf_type'] == 2).sum())} "
              f"basis={meta['loop_basis'].shape} n_tau={int(max(meta['tau_in'].max(), meta['tau_out'].max())) + 1}")


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "leaves":
        leaf_sidecars()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "taylor":  # add / refresh the Taylor workloads only
        with open(os.path.join(OUT, "MANIFEST.json")) as fh:
            manifest = json.load(fh)
        taylor_workloads(manifest)
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "ver3polar":  # add / refresh the vertex3 / polarization workloads only
        with open(os.path.join(OUT, "MANIFEST.json")) as fh:
            manifest = json.load(fh)
        ver3_polar_workloads(manifest)
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
            json.dump(manifest, fh, indent=1, sort_keys=True)
        return
    manifest = {}
    for order in (2, 3, 4, 5, 6):
        emit(f"gv_sigma_o{order}", lambda o=order: gv.diagsGV("sigma", o),
             f"GV.diagsGV(:sigma, {order}) + optimize!  (src/frontend/GV_diagrams/groups_sigma/Sigma{order}_0_0.diag)", manifest)
    for order in (1, 2, 3, 4):
        emit(f"gv_ver4_o{order}", lambda o=order: gv.diagsGV_ver4(o),
             f"GV.diagsGV_ver4({order}) + optimize!  (groups_vertex4/Vertex4{order}_0_0.diag; example/benchmark_GV.jl:23-25)",
             manifest)
    for order in (3, 4):
        emit(f"gv_ver4I_o{order}", lambda o=order: gv.diagsGV_ver4(o, channels=[gv.Alli]),
             f"GV.diagsGV_ver4({order}, channels=[Alli]) + optimize!  (groups_vertex4/Vertex4I{order}_0_0.diag)", manifest)
    from oracle.frontend import parquet as pq

    def pq_sigma(order):
        pq._ver4I.clear()
        return [r["diagram"] for r in pq.sigma(pq.DiagPara(type=pq.SigmaDiag, innerLoopNum=order))]

    def pq_ver4(order):
        pq._ver4I.clear()
        return [r["diagram"] for r in pq.vertex4(pq.DiagPara(type=pq.Ver4Diag, innerLoopNum=order))]

    for order in (2, 3, 4):
        emit(f"parquet_sigma_o{order}", lambda o=order: pq_sigma(o),
             f"Parquet.sigma(DiagPara(type=SigmaDiag, innerLoopNum={order})) + optimize!  (README.md:59-68)", manifest)
    for order in (2, 3, 4):
        emit(f"parquet_ver4_o{order}", lambda o=order: pq_ver4(o),
             f"Parquet.vertex4(DiagPara(type=Ver4Diag, innerLoopNum={order})) + optimize!  (example/benchmark.jl:13,23-25)",
             manifest)
    taylor_workloads(manifest)
    ver3_polar_workloads(manifest)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    leaf_sidecars()


if __name__ == "__main__":
    main()
