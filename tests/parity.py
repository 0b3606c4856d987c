"""Shared parity checks: the engine (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): adjusted coordinates 1e-9 m, sigma-zero 1e-12 (relative);
variances 2e-8 relative to the largest variance: the reference's explicit dpotri inverse is itself only
good to ~cond(N)*eps (3e-9 observed against a scaled NumPy inverse on cond(N) ~ 2e9), so a tighter bound
would test the oracle's rounding, not the engine's."""
import numpy as np

from dynadjust_b200 import engine, synth

TOL_XYZ = 1e-9
TOL_SIGMA0 = 1e-12
TOL_VCV_REL = 2e-8


def run_engine(lib_path, stn, msr, blocks=None, **opts):
    adj = engine.Adjustment(stn, msr, lib_path=lib_path, **opts)
    if blocks is not None:
        adj.set_blocks(blocks)
    info = adj.prepare()
    last = adj.adjust()
    stats = adj.statistics(write_back=True)
    return adj, info, last, stats


def check_against_oracle(oracle, lib_path, n_stations, n_baselines, seed, blocks=None, mutate=None, n_distances=0,
                         n_levels=0, tol_sigma0=TOL_SIGMA0, terrestrial=None, **opts):
    if terrestrial is not None:
        from dynadjust_b200 import synth_terrestrial
        stn, msr, truth, edges = synth_terrestrial.terrestrial_network(n_stations, n_baselines, seed, **terrestrial)
    elif n_distances or n_levels:
        stn, msr, truth, edges = synth.mixed_network(n_stations, n_baselines, seed, n_distances=n_distances, n_levels=n_levels)
    else:
        stn, msr, truth, edges = synth.gnss_network(n_stations, n_baselines, seed)
    if mutate:
        mutate(stn, msr, truth)
    stn_o, msr_o = stn.copy(), msr.copy()
    ref = oracle.adjust_simultaneous(stn_o, msr_o, want_vcv=True)
    rr = ref["res"]
    adj, info, last, stats = run_engine(lib_path, stn, msr, blocks=blocks(n_stations) if callable(blocks) else blocks, **opts)
    est = adj.estimates()
    assert last.iteration == rr.iterations, (last.iteration, rr.iterations)
    assert np.abs(est - ref["est"]).max() < TOL_XYZ
    assert stats.dof == rr.dof and stats.measurement_params == rr.measurement_params
    assert stats.unknown_params == rr.unknown_params
    assert abs(stats.sigma_zero - rr.sigma_zero) < tol_sigma0 * max(1.0, rr.sigma_zero)
    assert abs(stats.chi_squared - rr.chi_squared) < max(1e-10, 10 * tol_sigma0) * rr.chi_squared
    assert stats.outliers == rr.outliers
    assert abs(stats.global_pelzer - rr.global_pelzer) < 1e-9
    V = ref["vcv"]
    S = n_stations
    vscale = np.abs(np.diag(V)).max()
    q = adj.station_vcvs()
    qd = np.stack([V[3 * s:3 * s + 3, 3 * s:3 * s + 3] for s in range(S)])
    assert np.abs(q - qd).max() < TOL_VCV_REL * vscale
    rec = msr[:3 * n_baselines].reshape(-1, 3)
    step = max(1, len(rec) // 64)
    for b in range(0, len(rec), step):
        s1, s2 = int(rec["station1"][b, 0]), int(rec["station2"][b, 0])
        blk = adj.vcv_block(s1, s2)
        assert np.abs(blk - V[3 * s1:3 * s1 + 3, 3 * s2:3 * s2 + 3]).max() < TOL_VCV_REL * vscale
        assert np.abs(adj.vcv_block(s2, s1) - blk.T).max() == 0.0
    # statistics written back into the measurement records (ADJ:8187-8298)
    loose = tol_sigma0 > TOL_SIGMA0      # terrestrial rows: ~1e-9 m of libm-level noise in the computed heights
    red = 1e-9 if (loose and terrestrial is not None) else 0.0   # first-run reductions go through sin/cos/atan as well
    for f, tol in [("measCorr", 5e-9 if loose else 1e-9), ("measAdj", 5e-9 if loose else 1e-9),
                   ("measAdjPrec", 4 * TOL_VCV_REL * vscale), ("residualPrec", 4 * TOL_VCV_REL * vscale),
                   ("NStat", 1e-4 if loose else 1e-6), ("PelzerRel", 1e-6), ("preAdjCorr", red), ("term1", red),
                   ("preAdjMeas", 0.0), ("scale1", red), ("scale2", None), ("scale3", None), ("term2", None),
                   ("term3", None), ("term4", None)]:
        if tol is None:     # variances (scaled / derived on the first run): relative
            tol = 1e-12 * np.abs(msr_o[f]).max()
        assert np.abs(msr[f] - msr_o[f]).max() <= tol, f
    assert np.abs(stn["currentLatitude"] - stn_o["currentLatitude"]).max() < 1e-15
    assert np.abs(stn["currentHeight"] - stn_o["currentHeight"]).max() < 1e-8
    adj.close()
    return info


def chain_blocks(n_stations, width):
    """A .seg-like chain: consecutive runs of station indices (grid rows) as blocks."""
    return [list(range(b, min(n_stations, b + width))) for b in range(0, n_stations, width)]


def check_normals(oracle, lib_path, n_stations, n_baselines, seed, mutate=None, **opts):
    """Assembled N (constraints included) and w of the first iteration, block by block."""
    stn, msr, _, _ = synth.gnss_network(n_stations, n_baselines, seed)
    if mutate:
        mutate(stn, msr)
    ref = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_normals=True)
    adj = engine.Adjustment(stn, msr, lib_path=lib_path, **opts)
    adj.prepare()
    adj.iterate(normals=True)
    N, w = ref["normals"], ref["rhs"]
    scale = np.abs(N).max()
    assert np.abs(adj.rhs().ravel() - w).max() < 1e-12 * np.abs(w).max()
    rec = msr.reshape(-1, 3)
    for s in range(n_stations):
        blk = adj.normals_block(s, s)
        ref_blk = N[3 * s:3 * s + 3, 3 * s:3 * s + 3]
        assert np.abs(blk - ref_blk).max() <= 1e-13 * max(np.abs(ref_blk).max(), 1.0)
    for b in range(len(rec)):
        s1, s2 = int(rec["station1"][b, 0]), int(rec["station2"][b, 0])
        blk = adj.normals_block(s1, s2)
        ref_blk = N[3 * s1:3 * s1 + 3, 3 * s2:3 * s2 + 3]
        assert np.abs(blk - ref_blk).max() <= 1e-13 * max(np.abs(ref_blk).max(), 1.0)
    first = adj.corrections().ravel()
    assert np.abs(first - ref["first_corr"]).max() < 1e-7  # before the second iteration polishes it
    adj.close()
    return scale
