"""Restated diagram identifiers and enums of the reference front ends.  TEST / WORKLOAD INFRASTRUCTURE.

The evaluator never looks at these; they exist because leaf de-duplication (`optimize!`) compares leaves by
their identifier, which decides how many leaf columns a real workload graph has.

Reference: src/frontend/frontends.jl:9-46 (enums), src/frontend/diagram_id.jl:19-96 (bare propagator ids,
mirror_symmetrize, the equal-time rule of BareInteractionId), :98-206 (composite ids).
"""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Any, Tuple


class TwoBodyChannel(enum.IntEnum):
    Alli = 1
    PHr = 2
    PHEr = 3
    PPr = 4
    AnyChan = 5


class Filter(enum.IntEnum):
    Wirreducible = 0
    Girreducible = 1
    NoHartree = 2
    NoFock = 3
    NoBubble = 4
    Proper = 5
    DirectOnly = 6


class Response(enum.IntEnum):
    Composite = 0
    ChargeCharge = 1
    SpinSpin = 2
    ProperChargeCharge = 3
    ProperSpinSpin = 4
    UpUp = 5
    UpDown = 6


class AnalyticProperty(enum.IntEnum):
    Instant = 0
    Dynamic = 1


Alli, PHr, PHEr, PPr, AnyChan = TwoBodyChannel
(Wirreducible, Girreducible, NoHartree, NoFock, NoBubble, Proper, DirectOnly) = Filter
(Composite, ChargeCharge, SpinSpin, ProperChargeCharge, ProperSpinSpin, UpUp, UpDown) = Response
Instant, Dynamic = AnalyticProperty


def mirror_symmetrize(k) -> Tuple[float, ...]:
    """diagram_id.jl:81-96: flip the sign so that the first non-zero component is positive (and -0.0 -> 0.0)."""
    k = [float(x) for x in k]
    idx = next((i for i, x in enumerate(k) if x != 0), None)
    if idx is None or k[idx] > 0:
        return tuple(k)
    return tuple((-x) + 0.0 for x in k)


class DiagramId:
    """Generic field-wise equality (diagram_id.jl:339-349)."""

    def key(self):
        return (type(self).__name__,) + tuple(_freeze(getattr(self, f)) for f in self.__dataclass_fields__)

    def __eq__(self, other):
        return isinstance(other, DiagramId) and self.key() == other.key()

    def __hash__(self):
        return hash(self.key())


def _freeze(x):
    if isinstance(x, (list, tuple)):
        return tuple(_freeze(y) for y in x)
    if isinstance(x, DiagramId):
        return x.key()
    if hasattr(x, "key") and callable(x.key):
        return x.key()
    return x


@dataclass(eq=False)
class BareGreenId(DiagramId):
    type: AnalyticProperty
    extK: Tuple[float, ...]
    extT: Tuple[int, int]

    def __init__(self, type=Dynamic, *, k, t):
        self.type, self.extK, self.extT = type, mirror_symmetrize(k), tuple(int(x) for x in t)


@dataclass(eq=False)
class BareInteractionId(DiagramId):
    response: Response
    type: AnalyticProperty
    extK: Tuple[float, ...]
    extT: Tuple[int, int]

    def __init__(self, response, type=Instant, *, k, t=(0, 0)):
        self.response, self.type = response, type
        self.extK, self.extT = mirror_symmetrize(k), tuple(int(x) for x in t)

    def key(self):
        # diagram_id.jl:49-69: two equal-time interactions are the same whatever the time label
        t = "eq" if self.extT[0] == self.extT[1] else self.extT
        return ("BareInteractionId", int(self.response), int(self.type), self.extK, t)


@dataclass(eq=False)
class GenericId(DiagramId):
    para: Any
    extra: Any = None


@dataclass(eq=False)
class GreenId(DiagramId):
    para: Any
    type: AnalyticProperty
    extK: Tuple[float, ...]
    extT: Tuple[int, int]

    def __init__(self, para, type=Dynamic, *, k, t):
        self.para, self.type, self.extK, self.extT = para, type, mirror_symmetrize(k), tuple(int(x) for x in t)


@dataclass(eq=False)
class SigmaId(DiagramId):
    para: Any
    type: AnalyticProperty
    extK: Tuple[float, ...]
    extT: Tuple[int, int]

    def __init__(self, para, type, *, k, t=(0, 0)):
        self.para, self.type, self.extK, self.extT = para, type, mirror_symmetrize(k), tuple(int(x) for x in t)


@dataclass(eq=False)
class PolarId(DiagramId):
    para: Any
    response: Response
    extK: Tuple[float, ...]
    extT: Tuple[int, int]

    def __init__(self, para, response, *, k, t=(0, 0)):
        self.para, self.response, self.extK, self.extT = para, response, mirror_symmetrize(k), tuple(int(x) for x in t)


@dataclass(eq=False)
class Ver3Id(DiagramId):
    para: Any
    response: Response
    extK: Tuple[Tuple[float, ...], ...]
    extT: Tuple[int, int, int]

    def __init__(self, para, response, *, k, t=(0, 0, 0)):
        self.para, self.response = para, response
        self.extK = tuple(tuple(float(y) for y in x) for x in k)
        self.extT = tuple(int(x) for x in t)


@dataclass(eq=False)
class Ver4Id(DiagramId):
    para: Any
    response: Response
    type: AnalyticProperty
    channel: TwoBodyChannel
    extK: Tuple[Tuple[float, ...], ...]
    extT: Tuple[int, int, int, int]

    def __init__(self, para, response, type=Dynamic, *, k, t=(0, 0, 0, 0), chan=AnyChan):
        self.para, self.response, self.type, self.channel = para, response, type, chan
        self.extK = tuple(tuple(float(y) for y in x) for x in k)
        self.extT = tuple(int(x) for x in t)
