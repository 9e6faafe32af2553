"""Restated `leafstates` (per-leaf metadata the integrand needs).  TEST / WORKLOAD INFRASTRUCTURE.

Reference: src/frontend/frontends.jl:175-232 (`leafstates(leaf_maps, maxloopNum)` for `Graph` leaves) and its copy in
example/benchmark.jl:129-187; type codes from `FrontEnds.index`, src/frontend/diagram_id.jl:342-354.
Indices are returned 0-based (the reference's are 1-based Julia indices).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .ids import BareGreenId, BareInteractionId


def leafstates(leaves: Sequence, max_loop_num: int) -> Dict[str, np.ndarray]:
    """`leaves[k]` = the leaf graph behind leafVal column k (the reference's `leafmap[k + 1]`)."""
    n = len(leaves)
    leaf_type = np.zeros(n, np.int32)
    leaf_order = np.zeros((n, 2), np.int32)
    tau_in = np.zeros(n, np.int32)
    tau_out = np.zeros(n, np.int32)
    loop_index = np.zeros(n, np.int32)
    basis: List[np.ndarray] = []
    for k, leaf in enumerate(leaves):
        assert leaf.isleaf()
        pid = leaf.properties
        if isinstance(pid, BareGreenId):
            leaf_type[k] = 1
        elif isinstance(pid, BareInteractionId):
            leaf_type[k] = 2
        else:
            raise NotImplementedError("Not Implemented!")  # diagram_id.jl:352
        loopmom = np.zeros(max_loop_num)
        assert max_loop_num >= len(pid.extK)
        loopmom[: len(pid.extK)] = pid.extK
        for bi, b in enumerate(basis):  # frontends.jl:207-213: first basis vector that is ≈ this one
            if np.allclose(b, loopmom, rtol=1.5e-8, atol=0.0):
                loop_index[k] = bi
                break
        else:
            basis.append(loopmom)
            loop_index[k] = len(basis) - 1
        tau_in[k], tau_out[k] = pid.extT[0] - 1, pid.extT[1] - 1
        orders = list(leaf.orders) + [0, 0]
        leaf_order[k] = orders[:2]
    return {"leaf_type": leaf_type, "leaf_order": leaf_order, "tau_in": tau_in, "tau_out": tau_out, "loop_index": loop_index,
            "loop_basis": np.asarray(basis, np.float64).reshape(len(basis), max_loop_num)}
