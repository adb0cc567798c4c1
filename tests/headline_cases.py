"""Node-update parity cases at the shapes the bench numbers are quoted on (BASELINE.json configs[1..4]).

Each case is ONE node update (src/recursive_bp_factor.jl:146-165) of the centre of a star (or of the single node of an
InfiniteRegularGraph) whose incoming messages are seeded random full-bond trains.  The expected beliefs / free energy /
pair beliefs come from the oracle and are committed under tests/golden/headline_<name>.npz by
tests/golden/make_headline_golden.py (the oracle needs minutes per case at these sizes; the GPU test only loads them).

numpy only: imported by the golden generator (oracle side) and by the gpu test (device side).
"""
import numpy as np

CASES = {
    # name: kind, factor params, q, z (star degree) or k (infinite), T, d, message profile
    "glauber_z3_d20_T50": dict(kind="glauber", params=(0.5, 0.1, 1.0), q=2, z=3, T=50, d=20, decay=0.8, infinite=False, hard_obs=False),
    "glauber_z5_d20_T50": dict(kind="glauber", params=(0.5, 0.1, 1.0), q=2, z=5, T=50, d=20, decay=0.85, infinite=False, hard_obs=False),
    "sis_z3_d10_T50": dict(kind="sis", params=(0.1, 0.05), q=2, z=3, T=50, d=10, decay=0.7, infinite=False, hard_obs=False),
    "sirs_z3_d15_T40_hardobs": dict(kind="sirs", params=(0.4, 0.15, 0.15), q=3, z=3, T=40, d=15, decay=0.8, infinite=False, hard_obs=True),
    "glauber_inf_k4_d30_T8": dict(kind="glauber", params=(0.2, 0.0, 1.0), q=2, z=4, T=8, d=30, decay=0.9, infinite=True, hard_obs=False),
}


def bond_profile(T, d, q):
    """steady-state bond profile of a TruncBond(d) message: min((q*q)^t, (q*q)^(L-t), d)"""
    L = T + 1
    P = q * q
    return [int(min(float(P) ** t, float(P) ** (L - t), d)) for t in range(L + 1)]


def random_message(bonds, qs, qd, decay, rng):
    """positive random train whose entries are graded by decay^(m+n): the unfoldings have a decaying spectrum, like
    a converged BP message (and the represented function stays positive), so that TruncBond(d) binds at every site.  NOT normalised (callers normalise)."""
    out = []
    for t in range(len(bonds) - 1):
        m, n = bonds[t], bonds[t + 1]
        g = decay ** (np.arange(m)[:, None] + np.arange(n)[None, :])
        out.append((0.1 + rng.random((m, n, qs, qd))) * g[:, :, None, None])
    return out


def case_inputs(name):
    """deterministic inputs of a case: (und edges, N, q list, phi, message tensors per directed in-edge of node 0)"""
    c = CASES[name]
    rng = np.random.default_rng(sum(ord(ch) for ch in name))
    T, d, q = c["T"], c["d"], c["q"]
    L = T + 1
    z = c["z"]
    N = 1 if c["infinite"] else z + 1
    phi = [[np.ones(q) for _ in range(L)] for _ in range(N)]
    for i in range(N):
        p0 = 0.1 + rng.random(q)
        phi[i][0] = p0 / p0.sum()
    if c["hard_obs"]:
        # hard one-hot observations: the centre is observed at two times, rank of the trains drops
        for t in (T // 2, min(T, T // 2 + 7)):
            o = np.zeros(q)
            o[int(rng.integers(q))] = 1.0
            phi[0][t] = o
    bonds = bond_profile(T, d, q)
    nmsg = 1 if c["infinite"] else z
    msgs = [random_message(bonds, q, q, c["decay"], rng) for _ in range(nmsg)]
    und = [] if c["infinite"] else [(0, k) for k in range(1, z + 1)]
    return dict(und=und, N=N, q=[q] * N, phi=phi, msgs=msgs, bonds=bonds)
