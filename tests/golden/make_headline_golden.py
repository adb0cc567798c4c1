#!/usr/bin/env python
"""Generate tests/golden/headline_<case>.npz: the ORACLE's result of one node update at the shapes the bench numbers
are quoted on (tests/headline_cases.py).  Run here (CPU, minutes per case); the GPU test only loads the files.

    python tests/golden/make_headline_golden.py [case ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import factors as OF, mpbp as O, tt as OT  # noqa: E402
from tests.headline_cases import CASES, case_inputs  # noqa: E402

FACTORS = {"glauber": OF.HomogeneousGlauberFactor, "sis": OF.SISFactor, "sirs": OF.SIRSFactor}


def oracle_case(name):
    """returns (oracle bp after ONE update of node 0, inputs dict).  Shared with the gpu test's device-side set-up."""
    c = CASES[name]
    inp = case_inputs(name)
    T, d = c["T"], c["d"]
    L = T + 1
    fac = FACTORS[c["kind"]](*c["params"])
    if c["infinite"]:
        bp = O.mpbp_infinite_graph(c["z"], [fac] * L, c["q"], phi=[p.copy() for p in inp["phi"][0]])
        A = OT.TT([a.copy() for a in inp["msgs"][0]])
        OT.normalize(A)
        bp.mu[0] = A
    else:
        g = O.BiDiGraph(inp["N"], inp["und"])
        bp = O.MPBP(g, [[fac] * L for _ in range(inp["N"])], inp["q"], T, phi=[[p.copy() for p in ph] for ph in inp["phi"]])
        for k, e in enumerate(g.in_edges[0]):
            A = OT.TT([a.copy() for a in inp["msgs"][k]])
            OT.normalize(A)
            bp.mu[e] = A
    return bp, inp


def main():
    names = sys.argv[1:] or list(CASES)
    for name in names:
        c = CASES[name]
        t0 = time.perf_counter()
        bp, inp = oracle_case(name)
        in_sum = float(sum(np.sum(t) for A in bp.mu for t in A))
        O.onebpiter(bp, 0, OT.TruncBond(c["d"]))
        t1 = time.perf_counter()
        b0 = np.array(O.beliefs(bp)[0])
        pb, lz = O.pair_beliefs(bp)
        out_bonds = np.array([bp.mu[e].bond_dims() for e in range(bp.g.ne)], dtype=object)
        np.savez(os.path.join(ROOT, "tests", "golden", f"headline_{name}.npz"), belief0=b0, f0=float(bp.f[0]),
                 pair=np.array([np.array(p) for p in pb]), pair_logz=np.array(lz), input_checksum=in_sum,
                 oracle_seconds=t1 - t0, nedges=bp.g.ne)
        print(f"{name}: node update {t1 - t0:.1f} s, total {time.perf_counter() - t0:.1f} s, f0={bp.f[0]:.12g}, "
              f"out bonds {[max(b) for b in out_bonds]}", flush=True)


if __name__ == "__main__":
    main()
