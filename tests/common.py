"""Shared builders: the same model handed to the oracle (CPU restatement) and to the CUDA engine."""
import numpy as np

import mpbp_b200 as M
from oracle import factors as OF, mpbp as O, tt as OT


def make_factor(kind, params, both=True):
    """returns (oracle_factor, product_factor)"""
    table = {
        "glauber": (OF.HomogeneousGlauberFactor, M.HomogeneousGlauberFactor),
        "pmj": (OF.PMJGlauberFactor, M.PMJGlauberFactor),
        "intglauber": (OF.IntegerGlauberFactor, M.IntegerGlauberFactor),
        "sis": (OF.SISFactor, M.SISFactor),
        "sishet": (OF.SIS_heterogeneousFactor, M.SIS_heterogeneousFactor),
        "sirs": (OF.SIRSFactor, M.SIRSFactor),
    }
    a, b = table[kind]
    return a(*params), b(*params)


def otrunc(tr):
    return {0: lambda: OT.TruncBond(tr.d), 1: lambda: OT.TruncThresh(tr.eps), 2: lambda: OT.TruncBondThresh(tr.d, tr.eps)}[tr.kind]()


def build_pair(N, und, T, kinds, q, phi, psi=None, dmax=8):
    """kinds: per node (kind, params).  Returns (oracle bp, device bp)."""
    go = O.BiDiGraph(N, und)
    gd = M.IndexedBiDiGraph(N, und)
    assert list(gd.src) == go.src and list(gd.dst) == go.dst and list(gd.rev) == go.rev
    wo, wd = [], []
    for i in range(N):
        a, b = make_factor(*kinds[i])
        wo.append([a] * (T + 1))
        wd.append([b] * (T + 1))
    bo = O.MPBP(go, wo, q, T, phi=[[p.copy() for p in ph] for ph in phi], psi=None if psi is None else [[p.copy() for p in ps] for ps in psi])
    bd = M.mpbp(gd, wd, q, T, phi=phi, psi=psi, dmax=dmax)
    return bo, bd


def compare(bo, bd, atol=1e-8, pair=True):
    b_o = np.concatenate([np.array(b).ravel() for b in O.beliefs(bo)])
    b_d = np.concatenate([np.array(b).ravel() for b in M.beliefs(bd)])
    err_b = float(np.max(np.abs(b_o - b_d)))
    f_d = M.api.free_energy_contributions(bd)
    err_f = float(np.max(np.abs(bo.f - f_d)))
    err_p = 0.0
    if pair:
        po, lzo = O.pair_beliefs(bo)
        pd, lzd = M.pair_beliefs(bd)
        err_p = max(float(np.max(np.abs(np.array(a) - np.array(b)))) for a, b in zip(po, pd))
        err_p = max(err_p, float(np.max(np.abs(lzo - lzd))))
    return err_b, err_f, err_p
