"""Oracle-free checks of an adjustment at sizes no dense reference fits (verification helpers for tests and bench.py;
nothing here is on the product path).

``normal_identity_residual``: for a GNSS-only network the normal matrix is known in closed form from the records
(N_ss = constraint + sum of V^-1 over the station's baselines, N_st = -sum of V^-1 over the baselines s-t), so the
rigorous variances the engine returns can be checked against it station by station:

    sum_j N[s, j] * Z[j, s] = I_3        over j = s and the neighbours of s

uses exactly the blocks of Z = N^-1 the engine stores along the measured pairs (the pattern the reference reads for
its statistics, ADJ:7784-8060) — every diagonal block and every pair block of the selected inverse enters some
station's identity, so a wrong tile of the inverse anywhere shows up.  Extended precision for the sums.
"""
import numpy as np


def gnss_normal_blocks(stn, msr, fixed_sd=1e-6, free_sd=10.0):
    """Per GNSS baseline its V^-1 (m, 3, 3) and end stations; per station its constraint block — from the records as the
    engine left them (variances already scaled on the first run)."""
    rec = msr.reshape(-1, 3)
    assert (rec["measType"] == b"G").all(), "GNSS-only networks"
    m = len(rec)
    V = np.zeros((m, 3, 3))
    V[:, 0, 0] = rec["term2"][:, 0]
    V[:, 0, 1] = V[:, 1, 0] = rec["term2"][:, 1]
    V[:, 1, 1] = rec["term3"][:, 1]
    V[:, 0, 2] = V[:, 2, 0] = rec["term2"][:, 2]
    V[:, 1, 2] = V[:, 2, 1] = rec["term3"][:, 2]
    V[:, 2, 2] = rec["term4"][:, 2]
    Vinv = np.linalg.inv(V)
    s1 = rec["station1"][:, 0].astype(np.int64)
    s2 = rec["station2"][:, 0].astype(np.int64)
    const = stn["stationConst"]
    cb = np.zeros((len(stn), 3, 3))
    for code, sd in ((b"CCC", fixed_sd), (b"FFF", free_sd)):
        sel = const == code
        cb[sel] = np.eye(3) / (sd * sd)
    assert ((const == b"CCC") | (const == b"FFF")).all(), "CCC / FFF constraints only"
    return Vinv, s1, s2, cb


def normal_identity_residual(adj, stn, msr, sample=None, seed=0):
    """max over the (sampled) stations of | sum_j N_sj Z_js - I |, and the station where it occurs."""
    Vinv, s1, s2, cb = gnss_normal_blocks(stn, msr)
    n = len(stn)
    ld = np.longdouble
    Zd = adj.station_vcvs().astype(ld)                       # Z_ss
    Ze = adj.pair_vcvs(s1, s2).astype(ld)                    # Z[s1, s2] of every baseline's pair
    Vl = Vinv.astype(ld)
    # N_ss
    Nd = cb.astype(ld)
    np.add.at(Nd, s1, Vl)
    np.add.at(Nd, s2, Vl)
    acc = np.einsum("sij,sjk->sik", Nd, Zd)
    # N[s1, s2] = -V^-1 (per baseline; several baselines on a pair add up on their own), Z[s2, s1] = Z[s1, s2]^T
    np.add.at(acc, s1, -np.einsum("mij,mkj->mik", Vl, Ze))   # station s1: N[s1,s2] Z[s2,s1] = -Vinv * Ze^T
    np.add.at(acc, s2, -np.einsum("mij,mjk->mik", Vl, Ze))   # station s2: N[s2,s1] Z[s1,s2] = -Vinv * Ze
    res = np.abs(acc - np.eye(3, dtype=ld)).reshape(n, -1).max(axis=1)
    if sample is not None and sample < n:
        idx = np.random.default_rng(seed).choice(n, size=sample, replace=False)
        res = res[idx]
        worst = int(idx[np.argmax(res)])
    else:
        worst = int(np.argmax(res))
    return float(res.max()), worst
