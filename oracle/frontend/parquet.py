"""Restated Parquet front end (4-point vertex, self-energy, Green's function, 3-point vertex, polarisation).
TEST / WORKLOAD INFRASTRUCTURE.

Producer of the evaluator's input for BASELINE.json's Parquet configurations -- not part of the hot path.

Reference: src/frontend/parquet/parquet.jl:102-122 (DiagPara), :60-92 (Interaction, ParquetBlocks),
           common.jl:28-130 (partitions, tau / loop index bookkeeping), operation.jl:1-106 (mergeby),
           operation.jl:108-182 (update_extKT!), filter.jl:30-64, vertex4.jl:27-485, green.jl:21-115, sigma.jl:20-137,
           vertex3.jl:20-123, polarization.jl:16-135.

Known, documented deviations from a Julia run (they permute operands inside merged Sum nodes or root columns,
never the set of terms): `orderedPartition` iterates a Julia `Set` of permutations and `bubble!` iterates a Julia
`Dict` (hash order); here both use a deterministic order.  The all-leaves-one value of every graph is therefore
identical (checked against the closed-form diagram counts of benchmark/diagram_count.jl), while floating-point
rounding of a merged sum can differ from a particular Julia session in the last bits.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field, replace
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

from fdgraph_b200 import graph as G
from fdgraph_b200.graph import Graph, Prod, Sum

from . import gv
from .ids import (Alli, AnalyticProperty, AnyChan, BareGreenId, BareInteractionId, ChargeCharge, DiagramId, DirectOnly,
                  Dynamic, GenericId, Girreducible, GreenId, Instant, NoBubble, NoFock, NoHartree, PHEr, PHr, PolarId, PPr,
                  Proper, Response, SigmaId, SpinSpin, TwoBodyChannel, UpDown, UpUp, Ver3Id, Ver4Id, Wirreducible)

DI, EX = 0, 1  # 0-based positions of Julia's DI = 1, EX = 2
INL, OUTL, INR, OUTR = 0, 1, 2, 3
SYM_FACTOR = [1.0, -1.0, 1.0, -0.5, +1.0, -1.0]  # indexed by Int(chan) - 1
Di, Ex = "Di", "Ex"

VacuumDiag, SigmaDiag, GreenDiag, PolarDiag, Ver3Diag, Ver4Diag = range(6)


class Interaction:
    def __init__(self, response, types):
        self.response = response
        self.type = frozenset([types] if isinstance(types, AnalyticProperty) else types)

    def key(self):
        return (int(self.response), tuple(sorted(int(t) for t in self.type)))


class ParquetBlocks:
    def __init__(self, phi=(Alli, PHEr, PPr), ppi=(Alli, PHr, PHEr), gamma4=None):
        self.phi = list(phi)
        self.ppi = list(ppi)
        if gamma4 is None:
            gamma4 = list(self.phi) + [c for c in self.ppi if c not in self.phi]  # union(phi, ppi)
        self.gamma4 = list(gamma4)

    def key(self):
        return (frozenset(self.phi), frozenset(self.ppi), frozenset(self.gamma4))


def interaction_tau_num(has_tau: bool, interactions) -> int:
    if not has_tau:
        return 0
    for inter in interactions:
        if Dynamic in inter.type:
            return 2
    return 1


def inner_tau_num(kind: int, inner_loop_num: int, itn: int) -> int:
    if kind == Ver4Diag:
        return (inner_loop_num + 1) * itn
    if kind in (SigmaDiag, GreenDiag):
        return inner_loop_num * itn
    if kind == VacuumDiag:
        return (inner_loop_num - 1) * itn
    if kind == PolarDiag:
        return 1 + inner_tau_num(Ver3Diag, inner_loop_num - 1, itn)
    if kind == Ver3Diag:
        return 1 + inner_tau_num(Ver4Diag, inner_loop_num - 1, itn)
    raise NotImplementedError


def first_tau_idx(kind: int, offset: int = 0) -> int:
    return (3 if kind == GreenDiag else 1) + offset


def first_loop_idx(kind: int, offset: int = 0) -> int:
    return {Ver4Diag: 4, SigmaDiag: 2, GreenDiag: 2, PolarDiag: 2, Ver3Diag: 3, VacuumDiag: 1}[kind] + offset


@dataclass
class DiagPara:
    """parquet.jl:102-122 (@with_kw defaults)."""
    type: int
    innerLoopNum: int
    isFermi: bool = True
    spin: int = 2
    interaction: Tuple[Interaction, ...] = None
    firstLoopIdx: int = None
    totalLoopNum: int = None
    hasTau: bool = True
    firstTauIdx: int = None
    totalTauNum: int = None
    filter: Tuple = (NoHartree,)
    transferLoop: Tuple[float, ...] = ()
    extra: Any = None

    def __post_init__(self):
        if self.interaction is None:
            self.interaction = (Interaction(ChargeCharge, [Instant]),)
        self.interaction = tuple(self.interaction)
        self.filter = tuple(self.filter)
        self.transferLoop = tuple(float(x) for x in self.transferLoop)
        if self.firstLoopIdx is None:
            self.firstLoopIdx = first_loop_idx(self.type)
        if self.totalLoopNum is None:
            self.totalLoopNum = self.firstLoopIdx + self.innerLoopNum - 1
        if self.firstTauIdx is None:
            self.firstTauIdx = first_tau_idx(self.type)
        if self.totalTauNum is None:
            self.totalTauNum = self.firstTauIdx + inner_tau_num(self.type, self.innerLoopNum, self.interactionTauNum) - 1
        self._key = None

    @property
    def interactionTauNum(self) -> int:
        return interaction_tau_num(self.hasTau, self.interaction)

    def reconstruct(self, **kw) -> "DiagPara":
        """Parameters.reconstruct (parquet.jl:127-147): unspecified fields keep the parent's values."""
        d = {f: getattr(self, f) for f in ("type", "innerLoopNum", "isFermi", "spin", "interaction", "firstLoopIdx",
                                            "totalLoopNum", "hasTau", "firstTauIdx", "totalTauNum", "filter",
                                            "transferLoop", "extra")}
        d.update(kw)
        return DiagPara(**d)

    def key(self):
        """Base.isequal(::DiagPara, ::DiagPara) (parquet.jl:170-197)."""
        if self._key is None:
            extra = self.extra.key() if hasattr(self.extra, "key") else self.extra
            self._key = ("DiagPara", self.type, self.innerLoopNum, self.isFermi, self.spin,
                         frozenset(i.key() for i in self.interaction), self.firstLoopIdx, self.totalLoopNum, self.hasTau,
                         self.firstTauIdx, self.totalTauNum, frozenset(int(f) for f in self.filter),
                         self.transferLoop if self.transferLoop else None, extra)
        return self._key


# ---------------------------------------------------------------------------------------------------------
# common.jl
# ---------------------------------------------------------------------------------------------------------


def _partitions_fixed(total: int, n: int):
    """Integer partitions of `total` into exactly n positive parts, parts non-increasing."""
    def rec(rem, parts, maxpart):
        if parts == 1:
            if 1 <= rem <= maxpart:
                yield (rem,)
            return
        for first in range(min(rem - parts + 1, maxpart), 0, -1):
            for rest in rec(rem - first, parts - 1, first):
                yield (first,) + rest
    if n >= 1 and total >= n:
        yield from rec(total, n, total)


def ordered_partition(total_: int, n: int, lowerbound: int = 1) -> List[Tuple[int, ...]]:
    """common.jl:28-44."""
    assert lowerbound >= 0
    total = total_ - n * (lowerbound - 1)
    assert total >= n
    out: List[Tuple[int, ...]] = []
    for p in _partitions_fixed(total, n):
        p = tuple(x + lowerbound - 1 for x in p)
        assert sum(p) == total_
        out.extend(sorted(set(itertools.permutations(p))))
    return out


def get_k(loop_num: int, loop_idx: int) -> List[float]:
    k = [0.0] * loop_num
    k[loop_idx - 1] = 1.0
    return k


def find_first_loop_idx(partition, firstidx: int):
    acc = list(itertools.accumulate(partition, initial=firstidx))[1:]
    return [firstidx] + acc[:-1], acc[-1] - 1


def find_first_tau_idx(partition, kinds, firstidx: int, tau_num: int):
    taup = [inner_tau_num(kinds[i], p, tau_num) for i, p in enumerate(partition)]
    acc = list(itertools.accumulate(taup, initial=firstidx))[1:]
    return [firstidx] + acc[:-1], acc[-1] - 1


# ---------------------------------------------------------------------------------------------------------
# operation.jl: mergeby on a list-of-dict "DataFrame"
# ---------------------------------------------------------------------------------------------------------


def _sort_key(v):
    if isinstance(v, tuple):
        return tuple(_sort_key(x) for x in v)
    return int(v) if hasattr(v, "__int__") and not isinstance(v, float) else v


def _mergediag(group: List[dict], id_, operator, name) -> Graph:
    if len(group) == 1:
        if isinstance(id_, GenericId) or type(id_) is type(group[0]["diagram"].properties):
            return group[0]["diagram"]
    return Graph([r["diagram"] for r in group], properties=id_, operator=operator, name=name)


def mergeby_df(rows: List[dict], fields: Sequence[str], operator=None, name="none", getid: Optional[Callable] = None) -> List[dict]:
    """operation.jl:59-74 with DataFrames.groupby(df, fields, sort=true)."""
    if not rows:
        return rows
    operator = operator or Sum()
    fields = list(fields)
    if getid is None:
        def getid(g):
            return GenericId(g[0]["diagram"].properties.para, tuple(g[0][f] for f in fields))
    groups: Dict[Tuple, List[dict]] = {}
    for r in rows:
        groups.setdefault(tuple(r[f] for f in fields), []).append(r)
    out = []
    for key in sorted(groups, key=_sort_key):
        grp = groups[key]
        d = {f: v for f, v in zip(fields, key)}
        d["diagram"] = _mergediag(grp, getid(grp), operator, name)
        d["hash"] = d["diagram"].id
        out.append(d)
    return out


def mergeby_vec(diags: List[Graph], operator=None, name="none", getid: Optional[Callable] = None) -> List[Graph]:
    """operation.jl:78-92."""
    if not diags:
        return diags
    operator = operator or Sum()
    id_ = getid(diags) if getid else GenericId(diags[0].properties.para)
    if len(diags) == 1 and (isinstance(id_, GenericId) or type(id_) is type(diags[0].properties)):
        return diags
    return [Graph(diags, properties=id_, operator=operator, name=name)]


def _pre_order(graph: Graph):
    """AbstractTrees.PreOrderDFS over the tree expansion, pruned per object after the first visit."""
    seen = set()
    stack = [graph]
    while stack:
        node = stack.pop()
        yield node
        if id(node) in seen:
            continue
        seen.add(id(node))
        stack.extend(reversed(node.subgraphs))


def _deepcopy_graphs(graphs: List[Graph]) -> List[Graph]:
    memo: Dict[int, Graph] = {}
    for n in G.post_order_unique(graphs):
        c = object.__new__(type(n))
        c.id, c.name, c.orders = n.id, n.name, list(n.orders)
        c.subgraphs = [memo[id(s)] for s in n.subgraphs]
        c.subgraph_factors = list(n.subgraph_factors)
        c.operator, c.weight, c.properties = n.operator, n.weight, n.properties
        memo[id(n)] = c
    return [memo[id(g)] for g in graphs]


def _raw_id(prop, **updates):
    """FrontEnds.reconstruct (diagram_id.jl:366-383): positional constructor -> mirror-symmetrises propagator momenta."""
    c = object.__new__(type(prop))
    c.__dict__.update(prop.__dict__)
    c.__dict__.update(updates)
    if isinstance(c, (BareGreenId, BareInteractionId, GreenId, SigmaId, PolarId)):
        from .ids import mirror_symmetrize
        c.extK = mirror_symmetrize(c.extK)
    return c


def update_extKT(diags: List[Graph], para: DiagPara, leg_k: List[List[float]], extra_loop_idx: Optional[int] = None) -> List[Graph]:
    """operation.jl:108-188 on a deep copy: re-labels ids, re-expresses leaf momenta in the caller's loop basis."""
    graphs = _deepcopy_graphs(diags)
    visited = set()
    tau_idx = para.firstTauIdx
    n = len(leg_k[0])
    ext_k = leg_k[:-1]
    for graph in graphs:
        tau_shift = tau_idx - graph.properties.extT[0]
        for node in _pre_order(graph):
            if node.id in visited:
                continue
            node.id = G.uid()
            visited.add(node.id)
            prop = node.properties
            if prop is None or not hasattr(prop, "extK") or not hasattr(prop, "extT"):
                continue
            T = prop.extT
            if isinstance(prop, (Ver4Id, Ver3Id)):
                newk = tuple(tuple(float(x) for x in leg_k[i][:n]) for i in range(len(prop.extK)))
                upd = {"extK": newk, "para": para}
                if tau_shift != 0:
                    upd["extT"] = tuple(t + tau_shift for t in T)
                c = object.__new__(type(prop))
                c.__dict__.update(prop.__dict__)
                c.__dict__.update(upd)
                node.properties = c
            elif isinstance(prop, (BareGreenId, BareInteractionId, GreenId, SigmaId, PolarId)):
                K = list(prop.extK)
                orig = len(K)
                if orig < n:
                    K = K + [0.0] * (n - orig)
                    if extra_loop_idx is not None:
                        K[-1] = K[extra_loop_idx - 1]
                        K[extra_loop_idx - 1] = 0.0
                else:
                    K = K[:n]
                sum_k = [0.0] * n
                for i, k in enumerate(ext_k):
                    for q in range(n):
                        sum_k[q] += K[i] * k[q]
                order = sorted(range(len(ext_k)), key=lambda i: sum(1 for x in ext_k[i] if x != 0))  # sortperm (stable)
                indep: List[int] = []
                for i in order:
                    j = next(idx for idx in range(n) if idx not in indep and ext_k[i][idx] != 0)
                    indep.append(j)
                    K[i], K[j] = K[j], K[i]
                k_inner = [0.0] * n
                for idx in range(n):
                    if idx not in indep:
                        k_inner[idx] = K[idx]
                newk = tuple(sum_k[q] + k_inner[q] for q in range(n))
                if tau_shift != 0:
                    node.properties = _raw_id(prop, extK=newk, extT=tuple(t + tau_shift for t in T))
                else:  # K is mutated in place in the reference: no re-symmetrisation
                    c = object.__new__(type(prop))
                    c.__dict__.update(prop.__dict__)
                    c.extK = newk
                    node.properties = c
    return graphs


# ---------------------------------------------------------------------------------------------------------
# filter.jl
# ---------------------------------------------------------------------------------------------------------


def not_proper(para: DiagPara, K) -> bool:
    if Proper in para.filter:
        tl = para.transferLoop
        assert tl, "Please initialize para.transferLoop to check proper diagrams."
        return all(abs(a - b) <= 1.5e-8 * max(abs(a), abs(b), 0.0) or a == b for a, b in zip(tl[: len(K)], K))
    return False


def is_valid_g(flt, inner_loop_num: int) -> bool:
    if (NoFock in flt) and (NoHartree in flt) and inner_loop_num == 1:
        return False
    if (Girreducible in flt) and inner_loop_num > 0:
        return False
    return True


def is_valid_sigma(flt, inner_loop_num: int, subdiagram: bool) -> bool:
    if inner_loop_num == 0:
        return False
    if subdiagram and (Girreducible in flt):
        return False
    if subdiagram and (NoFock in flt) and (NoHartree in flt) and inner_loop_num == 1:
        return False
    return True


# ---------------------------------------------------------------------------------------------------------
# vertex4.jl
# ---------------------------------------------------------------------------------------------------------
_ver4I: Dict[int, List[Graph]] = {}


def get_ver4I() -> Dict[int, List[Graph]]:
    """parquet.jl:222-235: the fully irreducible GV vertices of order 3 and 4, loaded once."""
    if not _ver4I:
        _ver4I[3] = gv.diagsGV_ver4(3, channels=[Alli])
        _ver4I[4] = gv.diagsGV_ver4(4, channels=[Alli])
    return _ver4I


def max_ver4_tau_idx(para: DiagPara) -> int:
    return (para.innerLoopNum + 1) * para.interactionTauNum + para.firstTauIdx - 1


def max_ver4_loop_idx(para: DiagPara) -> int:
    return para.firstLoopIdx + para.innerLoopNum - 1


def _vadd(a, b):
    return [x + y for x, y in zip(a, b)]


def _vsub(a, b):
    return [x - y for x, y in zip(a, b)]


def leg_basis(chan, leg_k, loop_idx: int):
    k_in_l, k_out_l, k_in_r, k_out_r = leg_k
    K = [0.0] * len(k_in_l)
    K[loop_idx - 1] = 1.0
    if chan == PHr:
        Kx = _vsub(_vadd(k_out_l, K), k_in_l)
        L, R = [k_in_l, k_out_l, Kx, K], [K, Kx, k_in_r, k_out_r]
    elif chan == PHEr:
        Kx = _vsub(_vadd(k_out_r, K), k_in_l)
        L, R = [k_in_l, k_out_r, Kx, K], [K, Kx, k_in_r, k_out_l]
    elif chan == PPr:
        Kx = _vsub(_vadd(k_in_l, k_in_r), K)
        L, R = [k_in_l, Kx, k_in_r, K], [K, k_out_l, Kx, k_out_r]
    else:
        raise NotImplementedError
    return L, K, R, Kx


def tau_basis(chan, LvT, RvT):
    G0T = (LvT[OUTR], RvT[INL])
    if chan == PHr:
        extT = (LvT[INL], LvT[OUTL], RvT[INR], RvT[OUTR])
        GxT = (RvT[OUTL], LvT[INR])
    elif chan == PHEr:
        extT = (LvT[INL], RvT[OUTR], RvT[INR], LvT[OUTL])
        GxT = (RvT[OUTL], LvT[INR])
    elif chan == PPr:
        extT = (LvT[INL], RvT[OUTL], LvT[INR], RvT[OUTR])
        GxT = (LvT[OUTL], RvT[INR])
    else:
        raise NotImplementedError
    assert sorted(G0T + GxT + extT) == sorted(tuple(LvT) + tuple(RvT))
    return extT, G0T, GxT


def _factor(para: DiagPara, chan) -> float:
    f = SYM_FACTOR[int(chan) - 1]
    return f if para.isFermi else abs(f)


def _bare(para, diex, response, kind, which, inner_t, q, factor=1.0):
    sign = -1.0 if which == Di else (1.0 if para.isFermi else -1.0)
    if not not_proper(para, q) and which in diex:
        vid = BareInteractionId(response, kind, k=q, t=inner_t)
        return Graph([], factor=sign * factor, properties=vid)
    return None


def _push_bare(para, nodes, response, kind, ext_t, leg_k, vd, ve):
    if vd is not None:
        nodes.append(dict(response=response, type=kind, extT=ext_t[DI],
                          diagram=Graph([vd], operator=Sum(), properties=Ver4Id(para, response, kind, k=leg_k, t=ext_t[DI]))))
    if ve is not None:
        nodes.append(dict(response=response, type=kind, extT=ext_t[EX],
                          diagram=Graph([ve], operator=Sum(), properties=Ver4Id(para, response, kind, k=leg_k, t=ext_t[EX]))))


def _push_bare_with_response(para, nodes, response, kind, leg_k, q, diex, ext_t, inner_t):
    if response == UpUp:
        vd = _bare(para, diex, response, kind, Di, inner_t[DI], q[DI])
        ve = _bare(para, diex, response, kind, Ex, inner_t[EX], q[EX])
        _push_bare(para, nodes, UpUp, kind, ext_t, leg_k, vd, ve)
    elif response == UpDown:
        vd = _bare(para, diex, UpDown, kind, Di, inner_t[DI], q[DI])
        _push_bare(para, nodes, UpDown, kind, ext_t, leg_k, vd, None)
    elif response == ChargeCharge:
        vuud = _bare(para, diex, ChargeCharge, kind, Di, inner_t[DI], q[DI])
        vuue = _bare(para, diex, ChargeCharge, kind, Ex, inner_t[EX], q[EX])
        _push_bare(para, nodes, UpUp, kind, ext_t, leg_k, vuud, vuue)
        vupd = _bare(para, diex, ChargeCharge, kind, Di, inner_t[DI], q[DI])
        _push_bare(para, nodes, UpDown, kind, ext_t, leg_k, vupd, None)
    elif response == SpinSpin:
        vuud = _bare(para, diex, SpinSpin, kind, Di, inner_t[DI], q[DI])
        vuue = _bare(para, diex, SpinSpin, kind, Ex, inner_t[EX], q[EX])
        _push_bare(para, nodes, UpUp, kind, ext_t, leg_k, vuud, vuue)
        vupd = _bare(para, diex, SpinSpin, kind, Di, inner_t[DI], q[DI], -1.0)
        vupe = _bare(para, diex, SpinSpin, kind, Ex, inner_t[EX], q[EX], 2.0)
        _push_bare(para, nodes, UpDown, kind, ext_t, leg_k, vupd, vupe)
    else:
        raise NotImplementedError


def bare_ver4(nodes, para: DiagPara, leg_k, diex=(Di, Ex), leftalign=True):
    """vertex4.jl:350-408."""
    k_in_l, k_out_l, k_in_r = leg_k[0], leg_k[1], leg_k[2]
    t0 = para.firstTauIdx
    q = [_vsub(k_in_l, k_out_l), _vsub(k_in_r, k_out_l)]
    if para.hasTau:
        ext_ins = [(t0,) * 4, (t0,) * 4]
        ext_ins_right = [(t0 + 1,) * 4, (t0 + 1,) * 4]
        ext_dyn = [(t0, t0, t0 + 1, t0 + 1), (t0, t0 + 1, t0 + 1, t0)]
        inner_ins = [(1, 1), (1, 1)]
        inner_dyn = [(t0, t0 + 1), (t0, t0 + 1)]
    else:
        ext_ins = [(t0,) * 4, (t0,) * 4]
        ext_dyn = ext_ins
        inner_ins = [(1, 1), (1, 1)]
        inner_dyn = inner_ins
        ext_ins_right = ext_ins
    for inter in para.interaction:
        response, tv = inter.response, inter.type
        if Instant in tv and Dynamic not in tv:
            _push_bare_with_response(para, nodes, response, Instant, leg_k, q, diex, ext_ins, inner_ins)
        elif Instant not in tv and Dynamic in tv:
            _push_bare_with_response(para, nodes, response, Dynamic, leg_k, q, diex, ext_dyn, inner_dyn)
        elif Instant in tv and Dynamic in tv:
            _push_bare_with_response(para, nodes, response, Instant, leg_k, q, diex, ext_ins if leftalign else ext_ins_right, inner_dyn)
            _push_bare_with_response(para, nodes, response, Dynamic, leg_k, q, diex, ext_dyn, inner_dyn)
    return nodes


def _bubble2diag(ver8, para, chan, ldiag, rdiag, g0, gx, extrafactor):
    """vertex4.jl:230-285."""
    lid, rid = ldiag.properties, rdiag.properties
    ln, rn = lid.response, rid.response
    vtype = Dynamic  # typeMap
    ext_t, g0t, gxt = tau_basis(chan, lid.extT, rid.extT)
    fac = _factor(para, chan) * extrafactor

    def spin(r):
        return "↑↑" if r == UpUp else "↑↓"

    def add(lr, rr, vr, factor=1.0):
        key = (g0t, gxt, ext_t, vr, vtype)
        ver8.setdefault(key, [])
        if ln == lr and rn == rr:
            name = f"{spin(lr)}x{spin(rr)} → {chan.name},"
            ver8[key].append(Graph([ldiag, rdiag], properties=GenericId(para), operator=Prod(), factor=factor * fac, name=name))

    if chan == PHr:
        add(UpUp, UpUp, UpUp)
        add(UpDown, UpDown, UpUp)
        add(UpUp, UpDown, UpDown)
        add(UpDown, UpUp, UpDown)
    elif chan == PHEr:
        add(UpUp, UpUp, UpUp)
        add(UpDown, UpDown, UpUp)
        add(UpUp, UpUp, UpDown)
        add(UpDown, UpDown, UpDown)
        add(UpUp, UpDown, UpDown, -1.0)
        add(UpDown, UpUp, UpDown, -1.0)
    elif chan == PPr:
        add(UpUp, UpUp, UpUp)
        add(UpDown, UpDown, UpDown, -2.0)
        add(UpUp, UpDown, UpDown)
        add(UpDown, UpUp, UpDown)
    else:
        raise NotImplementedError


def bubble(ver4df, para: DiagPara, leg_k, chan, partition, level, name, blocks, blockstoplevel, extrafactor=1.0):
    """vertex4.jl:125-202."""
    tau_num = para.interactionTauNum
    oL, oG0, oR, oGx = partition
    if not is_valid_g(para.filter, oG0) or not is_valid_g(para.filter, oGx):
        return
    loop_idx = para.firstLoopIdx
    idx, max_loop = find_first_loop_idx(partition, loop_idx + 1)
    l_loop, g0_loop, r_loop, gx_loop = idx
    assert max_loop == max_ver4_loop_idx(para)
    idx, max_tau = find_first_tau_idx(partition, [Ver4Diag, GreenDiag, Ver4Diag, GreenDiag], para.firstTauIdx, tau_num)
    l_tau, g0_tau, r_tau, gx_tau = idx
    assert max_tau == max_ver4_tau_idx(para)

    l_para = para.reconstruct(type=Ver4Diag, innerLoopNum=oL, firstLoopIdx=l_loop, firstTauIdx=l_tau)
    r_para = para.reconstruct(type=Ver4Diag, innerLoopNum=oR, firstLoopIdx=r_loop, firstTauIdx=r_tau)
    gx_para = para.reconstruct(type=GreenDiag, innerLoopNum=oGx, firstLoopIdx=gx_loop, firstTauIdx=gx_tau)
    g0_para = para.reconstruct(type=GreenDiag, innerLoopNum=oG0, firstLoopIdx=g0_loop, firstTauIdx=g0_tau)

    if chan in (PHr, PHEr):
        gi = blockstoplevel.phi if level == 1 else blocks.phi
        gf = blockstoplevel.gamma4 if level == 1 else blocks.gamma4
    elif chan == PPr:
        gi = blockstoplevel.ppi if level == 1 else blocks.ppi
        gf = blockstoplevel.gamma4 if level == 1 else blocks.gamma4
    else:
        raise NotImplementedError

    l_leg, K, r_leg, Kx = leg_basis(chan, leg_k, loop_idx)
    lver = vertex4(l_para, l_leg, True, channels=gi, level=level + 1, name="Γi", blocks=blocks)
    if not lver:
        return
    rver = vertex4(r_para, r_leg, True, channels=gf, level=level + 1, name="Γf", blocks=blocks)
    if not rver:
        return

    ver8: Dict[Tuple, List[Graph]] = {}
    for lrow in lver:
        for rrow in rver:
            ldiag, rdiag = lrow["diagram"], rrow["diagram"]
            ext_t, g0t, gxt = tau_basis(chan, ldiag.properties.extT, rdiag.properties.extT)
            g0 = green(g0_para, K, g0t, True, name="G0", blocks=blocks)  # built (ids consumed) and dropped, like the reference
            gx = green(gx_para, Kx, gxt, True, name="Gx", blocks=blocks)
            _bubble2diag(ver8, para, chan, ldiag, rdiag, g0, gx, extrafactor)

    for key, lst in ver8.items():
        g0t, gxt, ext_t, vresponse, vtype = key
        g0 = green(g0_para, K, g0t, True, name="G0", blocks=blocks)
        gx = green(gx_para, Kx, gxt, True, name="Gx", blocks=blocks)
        id_ = Ver4Id(para, vresponse, vtype, k=leg_k, t=ext_t, chan=chan)
        if len(lst) == 1:
            diag = Graph([lst[0], g0, gx], properties=id_, operator=Prod())
        elif not lst:
            continue
        else:
            inner = Graph(lst, properties=GenericId(para), operator=Sum())
            diag = Graph([inner, g0, gx], properties=id_, operator=Prod())
        ver4df.append(dict(response=vresponse, type=vtype, extT=ext_t, diagram=diag))


def rpa_chain(ver4df, para, leg_k, chan, level, name, extrafactor=1.0):
    """vertex4.jl:204-213."""
    if chan not in (PHr, PHEr):
        return
    new_filter = tuple(dict.fromkeys(list(para.filter) + [Girreducible, DirectOnly]))
    para_rpa = para.reconstruct(filter=new_filter)
    blocks = ParquetBlocks(phi=[], ppi=[], gamma4=[PHr])
    bubble(ver4df, para_rpa, leg_k, chan, [0, 0, para.innerLoopNum - 1, 0], level, f"{name}_RPA_CT", blocks, blocks, extrafactor)


def add_alli(ver4df, para: DiagPara, leg_k):
    """vertex4.jl:113-123."""
    graphvec = update_extKT(get_ver4I()[para.innerLoopNum], para, leg_k, para.firstLoopIdx - 1)
    for d in graphvec:
        id_ = d.properties
        ver4df.append(dict(response=id_.response, type=id_.type, extT=id_.extT, diagram=d))


def merge_vertex4(para, ver4df, name, leg_k):
    if ver4df:
        ver4df = mergeby_df(ver4df, ["response", "type", "extT"], name=name,
                            getid=lambda g: Ver4Id(para, g[0]["response"], g[0]["type"], k=leg_k, t=g[0]["extT"]))
    return ver4df


def vertex4(para: DiagPara, ext_k=None, subdiagram=False, channels=(PHr, PHEr, PPr, Alli), level=1, name="none",
            blocks: Optional[ParquetBlocks] = None, blockstoplevel: Optional[ParquetBlocks] = None) -> List[dict]:
    """vertex4.jl:27-99 -> rows {response, type, extT, diagram, hash}."""
    blocks = blocks or ParquetBlocks()
    blockstoplevel = blockstoplevel or blocks
    if ext_k is None:
        ext_k = [get_k(para.totalLoopNum, 1), get_k(para.totalLoopNum, 2), get_k(para.totalLoopNum, 3)]
    for k in ext_k:
        assert len(k) >= para.totalLoopNum
    leg_k = [[float(x) for x in k[: para.totalLoopNum]] for k in ext_k[:3]]
    leg_k.append(_vsub(_vadd(leg_k[0], leg_k[2]), leg_k[1]))
    assert para.totalTauNum >= max_ver4_tau_idx(para), "Increase totalTauNum!"
    assert para.totalLoopNum >= max_ver4_loop_idx(para), "Increase totalLoopNum"
    assert PHr not in blocks.phi and PPr not in blocks.ppi
    loop_num = para.innerLoopNum
    ver4df: List[dict] = []
    if loop_num == 0:
        bare_ver4(ver4df, para, leg_k, [Di] if DirectOnly in para.filter else [Di, Ex])
    else:
        for c in channels:
            if c == Alli:
                if 3 <= loop_num <= 4:
                    add_alli(ver4df, para, leg_k)
                else:
                    continue
            for p in ordered_partition(loop_num - 1, 4, 0):
                if c in (PHr, PHEr, PPr):
                    bubble(ver4df, para, leg_k, c, list(p), level, name, blocks, blockstoplevel, 1.0)
            if NoBubble in para.filter and c in (PHr, PHEr):
                rpa_chain(ver4df, para, leg_k, c, level, name, -1.0)
    ver4df = merge_vertex4(para, ver4df, name, leg_k)
    assert all(r["extT"][0] == para.firstTauIdx for r in ver4df)
    return ver4df


# ---------------------------------------------------------------------------------------------------------
# green.jl / sigma.jl
# ---------------------------------------------------------------------------------------------------------


def green(para: DiagPara, ext_k=None, ext_t=None, subdiagram=False, name="G", blocks: Optional[ParquetBlocks] = None) -> Graph:
    """green.jl:21-115."""
    blocks = blocks or ParquetBlocks()
    if ext_k is None:
        ext_k = get_k(para.totalLoopNum, 1)
    if ext_t is None:
        ext_t = (1, 2) if para.hasTau else (0, 0)
    assert para.type == GreenDiag and is_valid_g(para.filter, para.innerLoopNum)
    assert len(ext_k) >= para.totalLoopNum
    ext_k = [float(x) for x in ext_k[: para.totalLoopNum]]
    tin, tout = ext_t
    t0 = para.firstTauIdx
    if para.innerLoopNum == 0:
        return Graph([], properties=BareGreenId(k=ext_k, t=ext_t), name=name)

    def sigma_g(group, oG, t_idx, k_idx, sigma_t_idx):
        para_g = para.reconstruct(type=GreenDiag, firstTauIdx=t_idx, firstLoopIdx=k_idx, innerLoopNum=oG)
        g = green(para_g, ext_k, group["GT"], True, blocks=blocks)
        pair_t = (("t", (sigma_t_idx, group["GT"][1])),)
        return Graph([group["diagram"], g], properties=GenericId(para, pair_t), operator=Prod(), name="ΣG")

    g0 = Graph([], properties=BareGreenId(k=ext_k, t=(tin, t0)), name="g0")
    pairs: List[Graph] = []
    for p in ordered_partition(para.innerLoopNum, 2, 0):
        o_sigma, o_g = p
        if not is_valid_sigma(para.filter, o_sigma, True) or not is_valid_g(para.filter, o_g):
            continue
        idx, max_tau = find_first_tau_idx(p, [SigmaDiag, GreenDiag], t0, para.interactionTauNum)
        assert max_tau <= para.totalTauNum
        s_tau, g_tau = idx
        idx, max_loop = find_first_loop_idx(p, para.firstLoopIdx)
        assert max_loop <= para.totalLoopNum
        s_loop, g_loop = idx
        sigma_para = para.reconstruct(type=SigmaDiag, firstTauIdx=s_tau, firstLoopIdx=s_loop, innerLoopNum=o_sigma)
        sig = sigma(sigma_para, ext_k, True, name="Σ", blocks=blocks)
        assert all(r["extT"][0] == s_tau for r in sig)
        df = [dict(r, Tin=r["extT"][0], GT=(r["extT"][1], ext_t[1])) for r in sig]
        groups = mergeby_df(df, ["GT"], operator=Sum())
        pairs.extend(sigma_g(g, o_g, g_tau, g_loop, s_tau) for g in groups)
    merged = mergeby_vec(pairs, operator=Sum(), name="gΣG")[0]
    return Graph([g0, merged], properties=GreenId(para, k=ext_k, t=ext_t), operator=Prod(), name=name)


def sigma(para: DiagPara, ext_k=None, subdiagram=False, name="Σ", blocks: Optional[ParquetBlocks] = None) -> List[dict]:
    """sigma.jl:20-137 -> rows {type, extT, diagram, hash}."""
    blocks = blocks or ParquetBlocks()
    assert para.type == SigmaDiag and para.innerLoopNum >= 1
    if ext_k is None:
        ext_k = get_k(para.totalLoopNum, 1)
    assert len(ext_k) >= para.totalLoopNum
    ext_k = [float(x) for x in ext_k[: para.totalLoopNum]]
    composite: List[dict] = []
    if not is_valid_sigma(para.filter, para.innerLoopNum, subdiagram):
        return composite
    K = [0.0] * len(ext_k)
    loop_idx = para.firstLoopIdx
    K[loop_idx - 1] = 1.0
    assert K != ext_k, "K and extK can not be the same"
    leg_k = [ext_k, K, K, ext_k]

    def gw_to_sigma(group, oW, para_g):
        response, kind = group["response"], group["type"]
        assert response in (UpUp, UpDown)
        sid = SigmaId(para, kind, k=ext_k, t=group["extT"])
        g = green(para_g, K, group["GT"], True, name="Gfock" if oW == 0 else "G_Σ", blocks=blocks)
        spinfactor = 2 if response == UpUp else -1
        if oW > 0:
            spinfactor *= 0.5
        return dict(type=kind, extT=group["extT"],
                    diagram=Graph([g, group["diagram"]], properties=sid, operator=Prod(), factor=spinfactor, name=name))

    for oG, oW in ordered_partition(para.innerLoopNum - 1, 2, 0):
        idx, max_loop = find_first_loop_idx([oW, oG], loop_idx + 1)
        assert max_loop <= para.totalLoopNum
        w_loop, g_loop = idx
        idx, max_tau = find_first_tau_idx([oW, oG], [Ver4Diag, GreenDiag], para.firstTauIdx, para.interactionTauNum)
        assert max_tau <= para.totalTauNum
        w_tau, g_tau = idx
        para_g = para.reconstruct(type=GreenDiag, innerLoopNum=oG, firstLoopIdx=g_loop, firstTauIdx=g_tau)
        para_w = para.reconstruct(type=Ver4Diag, innerLoopNum=oW, firstLoopIdx=w_loop, firstTauIdx=w_tau)
        if not is_valid_g(para_g.filter, para_g.innerLoopNum):
            continue
        if oW == 0:  # Fock-type Σ
            if NoHartree in para_w.filter:
                flt = tuple(dict.fromkeys(list(para_w.filter) + [Proper]))
                para_w0 = para_w.reconstruct(filter=flt, transferLoop=tuple(0.0 for _ in K))
                ver4 = vertex4(para_w0, leg_k, True, channels=[])
            else:
                ver4 = vertex4(para_w, leg_k, True, channels=[])
        else:  # composite Σ
            ver4 = vertex4(para_w, leg_k, True, channels=[PHr], blocks=blocks,
                           blockstoplevel=ParquetBlocks(phi=[], gamma4=[PHr, PHEr, PPr, Alli]))
        df = [dict(r, extT=(r["extT"][INL], r["extT"][OUTR]), GT=(r["extT"][OUTL], r["extT"][INR])) for r in ver4]
        groups = mergeby_df(df, ["response", "type", "GT", "extT"], operator=Sum())
        for row in groups:
            composite.append(gw_to_sigma(row, oW, para_g))
    if not composite:
        return composite
    sigmadf = mergeby_df(composite, ["type", "extT"], name=name,
                         getid=lambda g: SigmaId(para, g[0]["type"], k=ext_k, t=g[0]["extT"]))
    assert all(r["extT"][0] == para.firstTauIdx for r in sigmadf)
    return sigmadf


# ---------------------------------------------------------------------------------------------------------
# vertex3.jl / polarization.jl
# ---------------------------------------------------------------------------------------------------------


def _approx_vec(a, b) -> bool:
    """isapprox of two real vectors (norm(a - b) <= sqrt(eps) * max(norm(a), norm(b)))."""
    if len(a) != len(b):
        return False
    d = sum((x - y) ** 2 for x, y in zip(a, b)) ** 0.5
    return d <= 1.4901161193847656e-08 * max(sum(x * x for x in a) ** 0.5, sum(y * y for y in b) ** 0.5)


def _union_filter(first, flt):
    return tuple(dict.fromkeys([first] + list(flt)))


def vertex3(para: DiagPara, ext_k=None, subdiagram=False, name="Γ3", channels=(PHr, PHEr, PPr, Alli),
            blocks: Optional[ParquetBlocks] = None) -> List[dict]:
    """vertex3.jl:20-123 -> rows {response, extT, diagram, hash}.  extT = (bosonic, fermionic in, fermionic out)."""
    blocks = blocks or ParquetBlocks()
    if ext_k is None:
        ext_k = [get_k(para.totalLoopNum, 1), get_k(para.totalLoopNum, 2)]
    assert para.type == Ver3Diag
    assert para.innerLoopNum >= 1, "Only generates vertex corrections with more than one internal loops."
    for k in ext_k:
        assert len(k) >= para.totalLoopNum
    q = [float(x) for x in ext_k[0][: para.totalLoopNum]]
    k_in = [float(x) for x in ext_k[1][: para.totalLoopNum]]
    k_out = _vsub(k_in, q)
    assert not _approx_vec(q, k_in) and not _approx_vec(q, k_out), "The bosonic q cann't be same as the fermionic k."
    legs = [q, k_in, k_out]
    if Proper in para.filter and (len(para.transferLoop) != len(q) or not _approx_vec(para.transferLoop, q)):
        para = para.reconstruct(transferLoop=tuple(q))  # vertex3.jl:114-123
    t0 = para.firstTauIdx
    rows: List[dict] = []
    K = [0.0] * len(q)
    loop_idx = para.firstLoopIdx
    K[loop_idx - 1] = 1.0
    leg_k = [k_in, k_out, K, _vadd(K, q)]
    for part in ordered_partition(para.innerLoopNum - 1, 3, 0):
        o_ver4, o_gin, o_gout = part
        idx, max_loop = find_first_loop_idx(part, loop_idx + 1)
        assert max_loop <= para.totalLoopNum
        ver4_loop, gin_loop, gout_loop = idx
        ver4_t0 = para.firstTauIdx + 1 if para.hasTau else para.firstTauIdx
        idx, max_tau = find_first_tau_idx(part, [Ver4Diag, GreenDiag, GreenDiag], ver4_t0, para.interactionTauNum)
        assert max_tau <= para.totalTauNum
        ver4_tau, gin_tau, gout_tau = idx
        if not (is_valid_g(para.filter, o_gin) and is_valid_g(para.filter, o_gout)):
            continue
        para_gin = para.reconstruct(type=GreenDiag, innerLoopNum=o_gin, firstLoopIdx=gin_loop, firstTauIdx=gin_tau)
        para_gout = para.reconstruct(type=GreenDiag, innerLoopNum=o_gout, firstLoopIdx=gout_loop, firstTauIdx=gout_tau)
        para_ver4 = para.reconstruct(type=Ver4Diag, innerLoopNum=o_ver4, firstLoopIdx=ver4_loop, firstTauIdx=ver4_tau)
        ver4 = vertex4(para_ver4, leg_k, True, channels=channels, blocks=blocks)
        if not ver4:
            continue
        if para.hasTau:
            assert all(r["extT"][INL] == ver4_t0 for r in ver4), "The TinL of the inner Γ4 must be firstTauIdx+1"
        df = [dict(r, extT=(t0, r["extT"][INL], r["extT"][OUTL]), GinT=(t0, r["extT"][INR]), GoutT=(r["extT"][OUTR], t0))
              for r in ver4]
        for v4 in mergeby_df(df, ["response", "GinT", "GoutT", "extT"], operator=Sum()):
            response = v4["response"]
            assert response in (UpUp, UpDown)
            gin = green(para_gin, K, v4["GinT"], True, name="Gin", blocks=blocks)
            gout = green(para_gout, _vadd(K, q), v4["GoutT"], True, name="Gout", blocks=blocks)
            diag = Graph([gin, gout, v4["diagram"]], properties=Ver3Id(para, response, k=legs, t=v4["extT"]),
                         operator=Prod(), name=name)
            rows.append(dict(response=response, extT=v4["extT"], diagram=diag))
    if rows:
        rows = mergeby_df(rows, ["response", "extT"], name=name,
                          getid=lambda g: Ver3Id(para, g[0]["response"], k=legs, t=g[0]["extT"]))
    return rows


def polarization(para: DiagPara, ext_k=None, subdiagram=False, name="Π", blocks: Optional[ParquetBlocks] = None) -> List[dict]:
    """polarization.jl:16-135 -> rows {response, extT, diagram, hash}; every row has extT = (firstTauIdx, firstTauIdx + 1)."""
    blocks = blocks or ParquetBlocks()
    if ext_k is None:
        ext_k = get_k(para.totalLoopNum, 1)
    assert para.type == PolarDiag and para.innerLoopNum >= 1
    assert len(ext_k) >= para.totalLoopNum
    q_full = [float(x) for x in ext_k]
    if Proper not in para.filter or len(para.transferLoop) != len(q_full) or _approx_vec(para.transferLoop, q_full):
        # polarization.jl:129-135, condition as written there (a transfer loop EQUAL to q is replaced by q as well)
        para = para.reconstruct(transferLoop=tuple(q_full), filter=_union_filter(Proper, para.filter))
    q = q_full[: para.totalLoopNum]
    K = [0.0] * len(q)
    loop_idx = para.firstLoopIdx
    K[loop_idx - 1] = 1.0
    assert not _approx_vec(K, q)
    t0 = para.firstTauIdx
    ext_t = (t0, t0 + 1) if para.hasTau else (t0, t0)
    leg_k = [q, K, _vsub(K, q)]
    rows: List[dict] = []
    for part in ordered_partition(para.innerLoopNum - 1, 3, 0):
        o_ver3, o_gin, o_gout = part
        idx, max_loop = find_first_loop_idx(part, loop_idx + 1)
        assert max_loop <= para.totalLoopNum
        ver3_loop, gin_loop, gout_loop = idx
        if not (is_valid_g(para.filter, o_gin) and is_valid_g(para.filter, o_gout)):
            continue
        if o_ver3 == 0:  # Π0 = G G
            gt0 = ext_t[1] + 1 if para.hasTau else ext_t[0]
            idx, max_tau = find_first_tau_idx([o_gin, o_gout], [GreenDiag, GreenDiag], gt0, para.interactionTauNum)
            assert max_tau <= para.totalTauNum
            gin_tau, gout_tau = idx
            para_gin = para.reconstruct(type=GreenDiag, innerLoopNum=o_gin, firstLoopIdx=gin_loop, firstTauIdx=gin_tau)
            para_gout = para.reconstruct(type=GreenDiag, innerLoopNum=o_gout, firstLoopIdx=gout_loop, firstTauIdx=gout_tau)
            gin = green(para_gin, K, (ext_t[0], ext_t[1]), True, name="Gin")
            gout = green(para_gout, _vsub(K, q), (ext_t[1], ext_t[0]), True, name="Gout")
            sign = -1.0 if para.isFermi else 1.0
            diag = Graph([gin, gout], properties=PolarId(para, UpUp, k=q, t=ext_t), operator=Prod(), name=name, factor=sign)
            rows.append(dict(response=UpUp, extT=ext_t, diagram=diag))
            continue
        idx, max_tau = find_first_tau_idx(part, [Ver3Diag, GreenDiag, GreenDiag], ext_t[1], para.interactionTauNum)
        assert max_tau <= para.totalTauNum
        ver3_tau, gin_tau, gout_tau = idx
        para_gin = para.reconstruct(type=GreenDiag, innerLoopNum=o_gin, firstLoopIdx=gin_loop, firstTauIdx=gin_tau)
        para_gout = para.reconstruct(type=GreenDiag, innerLoopNum=o_gout, firstLoopIdx=gout_loop, firstTauIdx=gout_tau)
        para_ver3 = para.reconstruct(type=Ver3Diag, innerLoopNum=o_ver3, firstLoopIdx=ver3_loop, firstTauIdx=ver3_tau)
        ver3 = vertex3(para_ver3, leg_k, True, blocks=blocks)
        if not ver3:
            continue
        if para.hasTau:
            assert all(r["extT"][0] == ext_t[1] for r in ver3), "The bosonic T must be firstTauIdx+1 if hasTau"
            assert all(r["extT"][1] == ver3[0]["extT"][1] for r in ver3), "The TinL must be firstTauIdx+2 if hasTau"
        df = [dict(r, extT=ext_t, GinT=(ext_t[0], r["extT"][1]), GoutT=(r["extT"][2], ext_t[0])) for r in ver3]
        for v3 in mergeby_df(df, ["response", "GinT", "GoutT", "extT"], operator=Sum()):
            response = v3["response"]
            assert response in (UpUp, UpDown)
            gin = green(para_gin, K, v3["GinT"], True, name="Gin", blocks=blocks)
            gout = green(para_gout, _vsub(K, q), v3["GoutT"], True, name="Gout", blocks=blocks)
            diag = Graph([gin, gout, v3["diagram"]], properties=PolarId(para, response, k=q, t=v3["extT"]),
                         operator=Prod(), name=name)
            rows.append(dict(response=response, extT=v3["extT"], diagram=diag))
    if rows:
        rows = mergeby_df(rows, ["response", "extT"], name=name,
                          getid=lambda g: PolarId(para, g[0]["response"], k=q, t=ext_t))
    return rows
