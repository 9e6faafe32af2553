"""Restated GV front end (the `Graph`-producing readers).  TEST / WORKLOAD INFRASTRUCTURE.

Builds the reference's real workload graphs from its own `.diag` data files (read from the reference
checkout at generation time; the generated graphs are committed under workloads/ by tools/gen_workloads.py).

Reference: src/frontend/GV.jl:76-114 (diagsGV, diagsGV_ver4),
           src/frontend/GV_diagrams/readfile.jl:5-28 (_exchange), :191-265 (read_vertex4diagrams),
           :267-410 (read_one_vertex4diagram!), :412-473 (read_diagrams -> Graph), :475-588 (read_one_diagram!).
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, List, Sequence, Tuple

from fdgraph_b200.graph import Graph, Prod, Sum, linear_combination, multi_product

from .ids import (Alli, BareGreenId, BareInteractionId, ChargeCharge, Dynamic, GenericId, Instant, NoHartree, PHEr, PHr,
                  PolarId, PPr, Proper, SigmaId, SpinSpin, UpDown, UpUp, Ver4Id)

REF_DIAG_DIR = os.environ.get("FDG_REFERENCE_DIAGS", "/root/reference/src/frontend/GV_diagrams")
_INT = re.compile(r"[-+]?\d+")


def _ints(s: str) -> List[int]:
    return [int(m) for m in _INT.findall(s)]


def _exchange(perm: List[int], ver4legs: List[List[int]], index: int, ext_num: int = 2, offset_ver4: int = 0):
    """readfile.jl:15-28 (perm is 1-based, index is 1-based)."""
    pad = len(ver4legs) - offset_ver4
    v = index - 1
    inds = [(v >> k) & 1 for k in range(max(pad, v.bit_length()))]  # digits(index-1, base=2, pad=...)
    perm_ex = list(perm)
    legs_ex = [list(l) for l in ver4legs]
    for i, value in enumerate(reversed(inds), start=1):
        if value == 0:
            continue
        loc1 = perm.index(2 * i - 1 + ext_num)
        loc2 = perm.index(2 * i + ext_num)
        perm_ex[loc1], perm_ex[loc2] = perm_ex[loc2], perm_ex[loc1]
        j = i - 1 + offset_ver4
        legs_ex[j][1], legs_ex[j][3] = ver4legs[j][3], ver4legs[j][1]
    return perm_ex, legs_ex


def _read_header(block: str, keywords: Sequence[str]) -> Dict[str, List[int]]:
    out = {}
    for kw, line in zip(keywords, block.split("\n")):
        out[kw] = _ints(line)
    return out


def _blocks(path: str) -> List[str]:
    with open(path) as fh:
        text = fh.read()
    return [b for b in text.split("\n\n")]


class _Lines:
    def __init__(self, block: str):
        self.lines = block.split("\n")
        self.i = 0

    def next(self) -> str:
        s = self.lines[self.i] if self.i < len(self.lines) else ""
        self.i += 1
        return s

    def expect(self, word: str) -> None:
        line = self.next()
        assert word in line, f"expected {word!r} in {line!r}"


def _spin_weight(spin_factor: int, spin_polar: float) -> float:
    # sign(spinFactor) * (2 / (1 + spinPolarPara))^(log2(abs(spinFactor)))
    return math.copysign(1.0, spin_factor) * (2.0 / (1.0 + spin_polar)) ** math.log2(abs(spin_factor))


# ---------------------------------------------------------------------------------------------------------
# 4-point vertex
# ---------------------------------------------------------------------------------------------------------
_CHANNELS = {"PHr": PHr, "PHEr": PHEr, "PPr": PPr, "Alli": Alli}


def read_one_vertex4diagram(block: str, g_num: int, ver_num: int, loop_num: int, spin_polar: float = 0.0,
                            channels=(PHr, PHEr, PPr, Alli), filter=(NoHartree,), offset: int = -1):
    """readfile.jl:267-410 -> (g_Di, g_Ex) or () when the channel is filtered out."""
    flag_proper = Proper in filter
    is_dynamic = ver_num != 1
    io = _Lines(block)
    io.expect("Permutation")
    permutation = [x - offset for x in _ints(io.next())]
    assert len(permutation) == len(set(permutation)) == g_num
    io.expect("SymFactor")
    symfactor = float(io.next())
    io.expect("Channel")
    channel = _CHANNELS[io.next().strip()]
    if channel not in channels:
        return ()
    io.expect("GType")
    op_gtype = _ints(io.next())
    assert len(op_gtype) == g_num
    io.expect("VertexBasis")
    tau_labels = _ints(io.next())
    io.next()
    io.expect("LoopBasis")
    basis = [[0] * loop_num for _ in range(g_num)]  # currentBasis[g, loop]
    for i in range(loop_num):
        x = [int(t) for t in io.next().split()]
        assert len(x) == g_num
        for gi in range(g_num):
            basis[gi][i] = x[gi]
    io.expect("Ver4Legs")
    if ver_num == 0:
        ver4legs: List[List[int]] = []
        io_line = None
    else:
        strs = io.next().split("|")
        ver4legs = [_ints(s) for s in strs[:ver_num]]
    io.expect("WType")
    if ver_num > 0:
        io.next()
    io.expect("SpinFactor")
    spin_factors = _ints(io.next())
    io.expect("Di/Ex")
    di_ex = _ints(io.next())
    io.expect("Proper/ImProper")
    proper = _ints(io.next())

    inner_loop_num = loop_num - 3
    ext_k = [[0.0] * loop_num for _ in range(4)]
    for i in range(3):
        ext_k[i][i] = 1.0
        ext_k[3][i] = float((-1) ** i)
    ext_index = [1, 0, 2, 0]  # 1-based
    for ind1, ind2 in enumerate(permutation, start=1):
        if ind1 in (1, 2):
            continue
        if op_gtype[ind1 - 1] == -2:
            if ind2 == 1:
                ext_index[1] = ind1
            elif ind2 == 2:
                ext_index[3] = ind1
            else:
                raise ValueError(f"error GType for ({ind1}, {ind2}).")

    greens = []
    for ind1, ind2 in enumerate(permutation, start=1):
        if op_gtype[ind1 - 1] == -2:
            continue
        diagid = BareGreenId(k=basis[ind1 - 1], t=[tau_labels[ind1 - 1], tau_labels[ind2 - 1]])
        greens.append(Graph([], properties=diagid))
    fermi_green_prod = Graph(greens, operator=Prod())

    inter_di, inter_ex = [], []
    for iex, spin_factor in enumerate(spin_factors, start=1):
        if spin_factor == 0:
            continue
        if flag_proper and proper[iex - 1] == 1:
            continue
        permu, legs_ex = _exchange(permutation, ver4legs, iex)
        leafs = []
        ext_index[0] = permu[0]
        ext_index[2] = permu[1]
        for leg in legs_ex:
            ind1, ind2 = leg[1] - offset, leg[3] - offset
            current = [a - b for a, b in zip(basis[leg[0] - offset - 1], basis[ind1 - 1])]
            assert current == [a - b for a, b in zip(basis[ind2 - 1], basis[leg[2] - offset - 1])]  # momentum conservation
            diagid = BareInteractionId(ChargeCharge, k=current, t=[tau_labels[ind1 - 1], tau_labels[ind2 - 1]])
            leafs.append(Graph([], properties=diagid))
        node = Graph(leafs, operator=Prod(), factor=spin_factor * symfactor)
        (inter_di if di_ex[iex - 1] == 0 else inter_ex).append(node)

    ext_t = [tau_labels[i - 1] for i in ext_index]
    kind = Dynamic if is_dynamic else Instant
    id_di = Ver4Id((0, inner_loop_num), UpDown, kind, k=ext_k, t=ext_t, chan=channel)
    id_ex = Ver4Id((1, inner_loop_num), ChargeCharge, kind, k=ext_k, t=ext_t, chan=channel)
    if not fermi_green_prod.subgraphs:
        g_di = Graph(inter_di, operator=Sum(), properties=id_di)
        g_ex = Graph(inter_ex, operator=Sum(), properties=id_ex)
    else:
        g_di = multi_product(fermi_green_prod, Graph(inter_di, operator=Sum()), properties=id_di)
        g_ex = multi_product(fermi_green_prod, Graph(inter_ex, operator=Sum()), properties=id_ex)
    return g_di, g_ex


def read_vertex4diagrams(path: str, spin_polar: float = 0.0, filter=(NoHartree,), channels=(PHr, PHEr, PPr, Alli)):
    """readfile.jl:191-265 -> Vector{Graph} [guu, gud] per (extT, channel) group."""
    blocks = _blocks(path)
    hdr = _read_header(blocks[0], ["Vertex4", "DiagNum", "Order", "GNum", "Ver4Num", "LoopNum", "ExtLoopIndex",
                                   "DummyLoopIndex", "TauNum", "DummyTauIndex"])
    diag_num, g_num = hdr["DiagNum"][0], hdr["GNum"][0]
    ver_num, loop_num = hdr["Ver4Num"][1], hdr["LoopNum"][0]
    diagrams: List[Graph] = []
    for b in blocks[1:1 + diag_num]:
        diagrams.extend(read_one_vertex4diagram(b, g_num, ver_num, loop_num, spin_polar, channels=channels, filter=filter))
    inner_loop_num = loop_num - 3
    para = (2, inner_loop_num)
    # group by (extT, channel, Di/Ex) -- Julia Dict iteration order is not reproducible here; groups are emitted
    # in first-occurrence order of (extT, channel), which only permutes the root columns
    gr: Dict[Tuple, List[Graph]] = {}
    keys: List[Tuple] = []
    for d in diagrams:
        p = d.properties
        key = (p.extT, p.channel, p.para[0])
        gr.setdefault(key, []).append(d)
        if (p.extT, p.channel) not in keys:
            keys.append((p.extT, p.channel))
    out: List[Graph] = []
    for ext_t, channel in keys:
        g_di_list, g_ex_list = gr[(ext_t, channel, 0)], gr[(ext_t, channel, 1)]
        id_di = g_di_list[0].properties
        gud = linear_combination(g_di_list, properties=id_di)
        gex = linear_combination(g_ex_list, properties=g_ex_list[0].properties)
        guu_id = Ver4Id(para, UpUp, id_di.type, k=id_di.extK, t=id_di.extT, chan=id_di.channel)
        guu = Graph([gud, gex], properties=guu_id)
        out.extend([guu, gud])
    return out


# ---------------------------------------------------------------------------------------------------------
# self-energy / polarisation / free energy  (Graph output)
# ---------------------------------------------------------------------------------------------------------


def read_one_diagram(kind: str, block: str, g_num: int, ver_num: int, loop_num: int, ext_index: List[int],
                     spin_polar: float = 0.0, offset: int = -1, offset_ver4: int = 0) -> Graph:
    """readfile.jl:475-588."""
    is_dynamic = ver_num != 1
    io = _Lines(block)
    io.expect("Permutation")
    permutation = [x - offset for x in _ints(io.next())]
    assert len(permutation) == len(set(permutation)) == g_num
    io.expect("SymFactor")
    symfactor = float(io.next())
    io.expect("GType")
    op_gtype = _ints(io.next())
    assert len(op_gtype) == g_num
    io.expect("VertexBasis")
    tau_labels = [x - offset for x in _ints(io.next())]
    io.next()
    io.expect("LoopBasis")
    basis = [[0] * loop_num for _ in range(g_num)]
    for i in range(loop_num):
        x = [int(t) for t in io.next().split()]
        assert len(x) == g_num
        for gi in range(g_num):
            basis[gi][i] = x[gi]
    io.expect("Ver4Legs")
    if ver_num == 0:
        ver4legs: List[List[int]] = []
    else:
        ver4legs = [_ints(s) for s in io.next().split("|")[:ver_num]]
    io.expect("WType")
    if ver_num > 0:
        io.next()
    io.expect("SpinFactor")
    spin_factors = _ints(io.next())

    ext_index = [x - offset for x in ext_index]
    if kind == "sigma":
        ext_index[1] = permutation.index(ext_index[0]) + 1
    ext_num = len(ext_index)
    ext_k = [0.0] * loop_num

    greens = []
    for ind1, ind2 in enumerate(permutation, start=1):
        if op_gtype[ind1 - 1] == -2:
            continue
        diagid = BareGreenId(k=basis[ind1 - 1], t=[tau_labels[ind1 - 1], tau_labels[ind2 - 1]])
        greens.append(Graph([], properties=diagid))
    fermi_green_prod = Graph(greens, operator=Prod())

    interactions = []
    spinfactors_existed: List[float] = []
    for iex, spin_factor in enumerate(spin_factors, start=1):
        if spin_factor == 0:
            continue
        spinfactors_existed.append(_spin_weight(spin_factor, spin_polar))
        permu, legs_ex = _exchange(permutation, ver4legs, iex, ext_num, offset_ver4=offset_ver4)
        leafs = []
        for leg in legs_ex:
            ind1, ind2 = leg[1] - offset, leg[3] - offset
            current = [a - b for a, b in zip(basis[leg[0] - offset - 1], basis[ind1 - 1])]
            assert current == [a - b for a, b in zip(basis[ind2 - 1], basis[leg[2] - offset - 1])]
            diagid = BareInteractionId(ChargeCharge, k=current, t=[tau_labels[ind1 - 1], tau_labels[ind2 - 1]])
            leafs.append(Graph([], properties=diagid))
        if not leafs:
            continue
        interactions.append(Graph(leafs, operator=Prod()))

    inner_loop_num = loop_num - ext_num + 1
    ext_t = [tau_labels[i - 1] for i in ext_index]
    if kind == "freeEnergy":
        inner_loop_num -= 1
        diagid = GenericId(inner_loop_num)
    elif kind == "chargePolar":
        diagid = PolarId(inner_loop_num, ChargeCharge, k=ext_k, t=ext_t)
    elif kind == "spinPolar":
        diagid = PolarId(inner_loop_num, SpinSpin, k=ext_k, t=ext_t)
    elif kind == "sigma":
        diagid = SigmaId(inner_loop_num, Dynamic if is_dynamic else Instant, k=ext_k, t=ext_t)
    else:
        diagid = None
    facs = [w * symfactor for w in spinfactors_existed]
    if not interactions:
        return Graph([fermi_green_prod], subgraph_factors=facs, operator=Sum(), properties=diagid)
    inters = Graph(interactions, subgraph_factors=facs, operator=Sum())
    return multi_product(fermi_green_prod, inters, properties=diagid)


def read_diagrams(path: str, kind: str, spin_polar: float = 0.0) -> List[Graph]:
    """readfile.jl:412-473."""
    blocks = _blocks(path)
    hdr = _read_header(blocks[0], ["SelfEnergy", "DiagNum", "Order", "GNum", "Ver4Num", "LoopNum", "ExtLoopIndex",
                                   "DummyLoopIndex", "TauNum", "ExtTauIndex", "DummyTauIndex"])
    diag_num, g_num = hdr["DiagNum"][0], hdr["GNum"][0]
    ver_num, loop_num = hdr["Ver4Num"][1], hdr["LoopNum"][0]
    ext_index = hdr["ExtTauIndex"]
    offset_ver4 = 1 if kind == "sigma" else 0
    diagrams = [read_one_diagram(kind, b, g_num, ver_num, loop_num, list(ext_index), spin_polar, offset_ver4=offset_ver4)
                for b in blocks[1:1 + diag_num]]
    if kind == "freeEnergy":
        return [linear_combination(diagrams, properties=diagrams[0].properties)]
    gr: Dict[Tuple, List[Graph]] = {}
    order: List[Tuple] = []
    for d in diagrams:
        key = d.properties.extT
        if key not in gr:
            order.append(key)
        gr.setdefault(key, []).append(d)
    return [linear_combination(gr[k], properties=gr[k][0].properties) for k in order]


_DIRS = {"spinPolar": ("groups_spin", "Polar"), "chargePolar": ("groups_charge", "Polar"), "sigma": ("groups_sigma", "Sigma"),
         "green": ("groups_green", "Green"), "freeEnergy": ("groups_free_energy", "FreeEnergy")}


def diagsGV(kind: str, order: int, spin_polar: float = 0.0) -> List[Graph]:
    """GV.jl:76-93."""
    d, stem = _DIRS[kind]
    return read_diagrams(os.path.join(REF_DIAG_DIR, d, f"{stem}{order}_0_0.diag"), kind, spin_polar=spin_polar)


def diagsGV_ver4(order: int, spin_polar: float = 0.0, channels=(PHr, PHEr, PPr, Alli), filter=(NoHartree,)) -> List[Graph]:
    """GV.jl:106-114."""
    stem = "Vertex4I" if list(channels) == [Alli] else "Vertex4"
    path = os.path.join(REF_DIAG_DIR, "groups_vertex4", f"{stem}{order}_0_0.diag")
    return read_vertex4diagrams(path, spin_polar=spin_polar, channels=channels, filter=filter)
