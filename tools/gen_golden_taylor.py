"""Writes tests/golden/taylor_sigma2.json: the fixture behind the Taylor-mode AD known-answer test.

Run in the build container (reads the reference's GV `.diag` files).  The reference's own test
(test/taylor.jl:96-112) asserts, for the two order-2 self-energy graphs and the counter-term orders (g, v) below,
    eval!(GV.diagsGV(:sigma, 2, g, v)[1][i]) == eval!(taylorexpansion!(diagsGV(:sigma, 2, 0, 0))[i].coeffs[[g, v]])
with every leaf equal to one.  The left-hand sides come from the files Sigma2_<v>_<g>.diag and are stored here as
`expected`; the order-(0, 0) graphs (node arrays + which leaf is a propagator / an interaction) are stored so that
the test can run the restated `taylorexpansion` without the reference checkout.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import fdgraph_b200 as fd  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.frontend import gv  # noqa: E402
from oracle.frontend.ids import BareGreenId, BareInteractionId  # noqa: E402

ORDERS = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (2, 0), (1, 2), (2, 2)]  # test/taylor.jl:97


def read(order, v, g):
    path = os.path.join(gv.REF_DIAG_DIR, "groups_sigma", f"Sigma{order}_{v}_{g}.diag")
    graphs = gv.read_diagrams(path, "sigma")
    graphs.sort(key=lambda x: 0 if x.properties.extT[0] == x.properties.extT[1] else 1)  # static first, readfile.jl:176-180
    return graphs


def ones_eval(graphs):
    raw, _ = fd.flatten(graphs)
    orc = O.Oracle(raw)
    return [float(x) for x in orc.eval(np.ones((max(orc.n_leaves, 1), 1)))[:, 0]]


def main():
    fd.uidreset()
    base = read(2, 0, 0)
    raw, nodes = fd.flatten(base)
    kind = []
    for n in nodes:
        p = n.properties
        kind.append("G" if isinstance(p, BareGreenId) else "W" if isinstance(p, BareInteractionId) else "")
    out = {
        "source": "src/frontend/GV_diagrams/groups_sigma/Sigma2_<v>_<g>.diag, all leaves one; test/taylor.jl:96-112",
        "graph": {k: getattr(raw, k).tolist() for k in raw.__dataclass_fields__},
        "node_kind": kind,
        "expected": {f"{g},{v}": ones_eval(read(2, v, g)) for g, v in ORDERS},
    }
    path = os.path.join(ROOT, "tests", "golden", "taylor_sigma2.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path, out["expected"])


if __name__ == "__main__":
    main()
