"""CPU tests of the engine's host logic (symbolic analysis, planner, control flow, C-ABI surface).

These run the product's host code against plain-loop stand-ins for the CUDA kernels
(tests/hostsim) — they validate indices, maps and the algorithm, not the kernels; the kernels
themselves are checked by the `-m gpu` tests."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from dynadjust_b200 import engine, synth
from tests import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,m,seed,leaf", [(100, 300, 1235, 16), (400, 1200, 5, 8), (1000, 3000, 7, 24)])
def test_nested_dissection_matches_oracle(oracle, hostsim_path, n, m, seed, leaf):
    info = parity.check_against_oracle(oracle, hostsim_path, n, m, seed, leaf_stations=leaf)
    assert info.nfronts > 1 and info.nlevels > 1


def test_dense_front_matches_oracle(oracle, hostsim_path):
    # one dense front wider than one pivot tile (k = 180 > NB = 128)
    info = parity.check_against_oracle(oracle, hostsim_path, 60, 170, 21, ordering=engine.ORDER_DENSE)
    assert info.nfronts == 1


def test_dense_front_block_doubling(oracle, hostsim_path):
    # k = 750 = 5 full pivot tiles + 110: the triangular inverse joins blocks of 128, 256 and 512 with short
    # second halves (pairs (4,5), then (0-3, 4-5)) — every shape of the block-doubling step
    info = parity.check_against_oracle(oracle, hostsim_path, 250, 750, 23, ordering=engine.ORDER_DENSE)
    assert info.nfronts == 1


@pytest.mark.parametrize("tile", [64, 128])
def test_gemm_tile_shapes_plan(oracle, hostsim_path, tile):
    """The planner's tile lists for either shape of the tile GEMM (the stand-in kernel walks exactly the listed tiles with
    the launch's shape): dissection with multi-tile fronts, a chain, a dense front wider than several pivot tiles — incl.
    the in-place pivot-panel product, which must stay in one column tile."""
    parity.check_against_oracle(oracle, hostsim_path, 1000, 3000, 7, leaf_stations=24, gemm_tile=tile)
    parity.check_against_oracle(oracle, hostsim_path, 300, 900, 9, blocks=lambda n: parity.chain_blocks(n, 40), gemm_tile=tile)
    parity.check_against_oracle(oracle, hostsim_path, 250, 750, 23, ordering=engine.ORDER_DENSE, gemm_tile=tile)


def test_gemm_tile_option_is_validated(hostsim_path):
    stn, msr, _, _ = synth.gnss_network(40, 120, 3)
    adj = engine.Adjustment(stn, msr, lib_path=hostsim_path, gemm_tile=96)
    with pytest.raises(engine.AdjustmentError) as e:
        adj.prepare()
    assert "gemm_tile" in str(e.value)


def test_gemm_flags_reference_semantics(hostsim_path):
    """gadj_test_gemm_ex on the stand-in kernel: the flag semantics the GPU test of the same name holds the CUDA kernel to."""
    ACCUM, NEG, LOWER, KLO_ROW, KHI_ROW, DUAL = 1, 2, 4, 16, 64, 128
    rng = np.random.default_rng(5)
    adj = engine.Adjustment(lib_path=hostsim_path)
    for tile in (64, 128):
        A, B, C0 = rng.standard_normal((150, 70)), rng.standard_normal((150, 70)), rng.standard_normal((150, 150))
        C, _, _ = adj.test_gemm_ex(A, B, C0, flags=ACCUM | NEG | LOWER, tile=tile)
        ref = C0 - A @ B.T
        ref[np.triu_indices(150, 1)] = C0[np.triu_indices(150, 1)]
        assert np.abs(C - ref).max() < 1e-12
        A = np.triu(rng.standard_normal((130, 130)))
        B = rng.standard_normal((90, 130))
        C, Ct, _ = adj.test_gemm_ex(A, B, flags=KLO_ROW | DUAL, tile=tile)
        assert np.abs(C - A @ B.T).max() < 1e-12 and np.abs(Ct - C.T).max() == 0.0
        C, _, _ = adj.test_gemm_ex(np.tril(A.T), B, flags=KHI_ROW, tile=tile)
        assert np.abs(C - np.tril(A.T) @ B.T).max() < 1e-12
    adj.close()


SLOW = pytest.mark.skipif(not os.environ.get("GADJ_SLOW_TESTS"), reason="minutes of dense CPU oracle: set GADJ_SLOW_TESTS=1")


@SLOW
def test_dense_oracle_at_5000_mixed_stations(oracle, hostsim_path):
    """The largest size the dense oracle (explicit inverse of the 15 000 x 15 000 normals) finishes in minutes: 5 000
    stations, 15 000 GNSS baselines + 4 000 slope distances + 3 000 levelled height differences, 169 fronts on 8 levels;
    north_star tolerances (1e-9 m; sigma-zero 1e-7 with local-frame rows, see test_gpu_parity.py).  3 minutes on 8 cores;
    the same check on the device is tests/test_gpu_parity.py::test_dense_oracle_at_5000_mixed_stations."""
    info = parity.check_against_oracle(oracle, hostsim_path, 5000, 15000, 101, n_distances=4000, n_levels=3000,
                                       leaf_stations=64, tol_sigma0=1e-7)
    assert info.nfronts > 100


def test_chain_blocks_match_oracle(oracle, hostsim_path):
    # a .seg-style chain of blocks (phased adjustment) is rigorous: same answer as simultaneous
    info = parity.check_against_oracle(oracle, hostsim_path, 300, 900, 9, blocks=lambda n: parity.chain_blocks(n, 40))
    assert info.nfronts == 8 and info.nlevels == 8


def test_degree_20_network(oracle, hostsim_path):
    # the C4 edge recipe (10 neighbour offsets per station) at a size the oracle finishes in seconds
    parity.check_against_oracle(oracle, hostsim_path, 400, 3600, 31, leaf_stations=32)


def test_mixed_constraints(oracle, hostsim_path):
    def mutate(stn, msr, truth):
        # partially constrained stations (held at their true position in the constrained components)
        for s, code in ((5, b"CCF"), (17, b"FFC"), (23, b"CFC")):
            stn["stationConst"][s] = code
            lat, lon, h = synth.cart_to_geo(truth[s])
            stn["currentLatitude"][s], stn["currentLongitude"][s], stn["currentHeight"][s] = lat, lon, h
    parity.check_against_oracle(oracle, hostsim_path, 120, 360, 77, mutate=mutate, leaf_stations=16)


def test_variance_scaling(oracle, hostsim_path):
    def mutate(stn, msr, truth):
        msr["scale4"] = 2.5                    # whole-matrix scalar
        rec = msr.reshape(-1, 3)
        rec["scale1"][::3] = 1.5               # phi
        rec["scale3"][::5] = 3.0               # height
    parity.check_against_oracle(oracle, hostsim_path, 100, 300, 8, mutate=mutate, leaf_stations=16)


def test_mixed_terrestrial_rows(oracle, hostsim_path):
    """GNSS baselines + slope distances 'S' + levelled height differences 'L' (geoid-reduced on the first run):
    the partials move with the estimates, so the normals are rebuilt and refactorised on every iteration."""
    parity.check_against_oracle(oracle, hostsim_path, 200, 600, 17, n_distances=150, n_levels=120, leaf_stations=16)
    parity.check_against_oracle(oracle, hostsim_path, 300, 500, 23, n_distances=400, n_levels=300, leaf_stations=24)


@pytest.mark.parametrize("kind", list("ABKCEMSVZLHRIJPQ"))
def test_every_scalar_type(oracle, hostsim_path, kind):
    """One-, two- and three-station rows of each terrestrial type with deflections of the vertical
    on half of the stations: first-run reductions, per-iteration re-linearisation, statistics written to the records."""
    parity.check_against_oracle(oracle, hostsim_path, 120, 260, 31 + ord(kind), terrestrial=dict(scalars={kind: 260}),
                                leaf_stations=16)


def test_direction_sets(oracle, hostsim_path):
    """D sets: derived angles, tridiagonal angle VCV inverted on the device stand-in, ignored directions inside a set."""
    parity.check_against_oracle(oracle, hostsim_path, 150, 300, 41, terrestrial=dict(n_dir_sets=90), leaf_stations=16)
    parity.check_against_oracle(oracle, hostsim_path, 150, 300, 42, terrestrial=dict(n_dir_sets=90, ignore_some=True),
                                leaf_stations=24)


def test_gnss_clusters(oracle, hostsim_path):
    """X baseline clusters and Y point clusters with full VCVs (covariance records), with and without the v-scale."""
    parity.check_against_oracle(oracle, hostsim_path, 150, 300, 43, terrestrial=dict(n_x=50, n_y=30), leaf_stations=16)
    parity.check_against_oracle(oracle, hostsim_path, 150, 300, 44, terrestrial=dict(n_x=40, n_y=40, v_scale=2.5),
                                leaf_stations=16)


def _y_clusters_as_llh(stn, msr):
    """Rewrite every Cartesian Y cluster of `msr` in latitude / longitude / orthometric-height form (the layout of the
    reference's urban sample): values through CartToGeo, variance matrix through the inverse Jacobians at the points."""
    i = 0
    while i < len(msr):
        if msr["measType"][i] != b"Y" or msr["measStart"][i] != 0:
            i += 1
            continue
        count = int(msr["vectorCount1"][i])
        rec, j = [], i
        for _ in range(count):
            rec.append(j)
            j += 3 + 3 * int(msr["vectorCount2"][j])
        n = 3 * count
        V = np.zeros((n, n))
        for k, r in enumerate(rec):
            v = 3 * k
            V[v, v], V[v, v + 1], V[v + 1, v + 1] = msr["term2"][r], msr["term2"][r + 1], msr["term3"][r + 1]
            V[v, v + 2], V[v + 1, v + 2], V[v + 2, v + 2] = msr["term2"][r + 2], msr["term3"][r + 2], msr["term4"][r + 2]
            for q in range(int(msr["vectorCount2"][r])):
                for x in range(3):
                    cv = r + 3 + 3 * q + x
                    V[v + x, v + 3 + 3 * q:v + 6 + 3 * q] = msr["term1"][cv], msr["term2"][cv], msr["term3"][cv]
        V = np.triu(V) + np.triu(V, 1).T
        J = np.zeros((n, n))
        for k, r in enumerate(rec):
            s = int(msr["station1"][r])
            lat, lon, h = synth.cart_to_geo(np.array([[msr["term1"][r], msr["term1"][r + 1], msr["term1"][r + 2]]]))
            # Jacobian at the station's a-priori position, as the engine / reference will use
            la, lo, hh = stn["currentLatitude"][s], stn["currentLongitude"][s], stn["currentHeight"][s]
            eps = 1e-7
            f = lambda a, b, c: synth.geo_to_cart(np.array([a]), np.array([b]), np.array([c]))[0]
            Jk = np.stack([(f(la + eps, lo, hh) - f(la - eps, lo, hh)) / (2 * eps), (f(la, lo + eps, hh) - f(la, lo - eps, hh)) / (2 * eps),
                           (f(la, lo, hh + 1.0) - f(la, lo, hh - 1.0)) / 2.0], axis=1)
            J[3 * k:3 * k + 3, 3 * k:3 * k + 3] = Jk
            msr["term1"][r], msr["term1"][r + 1], msr["term1"][r + 2] = lat[0], lon[0], h[0] - float(stn["geoidSep"][s])
            msr["coordType"][r:r + 3] = b"LLH"
        Ji = np.linalg.inv(J)
        G = Ji @ V @ Ji.T
        for k, r in enumerate(rec):
            v = 3 * k
            msr["term2"][r], msr["term2"][r + 1], msr["term3"][r + 1] = G[v, v], G[v, v + 1], G[v + 1, v + 1]
            msr["term2"][r + 2], msr["term3"][r + 2], msr["term4"][r + 2] = G[v, v + 2], G[v + 1, v + 2], G[v + 2, v + 2]
            for q in range(int(msr["vectorCount2"][r])):
                for x in range(3):
                    cv = r + 3 + 3 * q + x
                    msr["term1"][cv], msr["term2"][cv], msr["term3"][cv] = G[v + x, v + 3 + 3 * q:v + 6 + 3 * q]
        i = j


def test_point_clusters_in_geographic_form(oracle, hostsim_path):
    """Y clusters supplied as latitude / longitude / orthometric height (ADJ:6281-6325, 4563-4644): converted to
    Cartesian on the first run with geoid reduction and variance propagation, written back into the records.  The
    engine must agree with the oracle, and both with the same network given in Cartesian form."""
    from dynadjust_b200 import synth_terrestrial as st
    stn, msr, truth, _ = st.terrestrial_network(120, 300, 47, n_y=25, deflections=False)
    ref_xyz = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_vcv=False)
    msr_llh = msr.copy()
    _y_clusters_as_llh(stn, msr_llh)
    assert (msr_llh["coordType"][msr_llh["measType"] == b"Y"] == b"LLH").any()
    s_o, m_o = stn.copy(), msr_llh.copy()
    ref = oracle.adjust_simultaneous(s_o, m_o, want_vcv=False)
    assert np.abs(ref["est"] - ref_xyz["est"]).max() < 1e-5 and abs(ref["res"].sigma_zero - ref_xyz["res"].sigma_zero) < 1e-5
    s_e, m_e = stn.copy(), msr_llh.copy()
    adj, info, last, stats = parity.run_engine(hostsim_path, s_e, m_e, leaf_stations=16)
    assert np.abs(adj.estimates() - ref["est"]).max() < parity.TOL_XYZ
    assert abs(stats.sigma_zero - ref["res"].sigma_zero) < 1e-11 and stats.dof == ref["res"].dof
    y = m_e["measType"] == b"Y"
    for f in ("term1", "term2", "term3", "term4", "preAdjMeas", "preAdjCorr", "measAdj", "measCorr"):
        assert np.abs(m_e[f][y] - m_o[f][y]).max() <= 1e-9 * max(1.0, np.abs(m_o[f][y]).max()), f
    assert (m_e["coordType"][y] == b"XYZ").all() and (m_e["station3"][y & (m_e["measStart"] <= 2)] == 2).all()
    adj.close()


def _cluster_matrix(msr, first):
    count = int(msr["vectorCount1"][first])
    rec, j = [], first
    for _ in range(count):
        rec.append(j)
        j += 3 + 3 * int(msr["vectorCount2"][j])
    n = 3 * count
    V = np.zeros((n, n))
    for k, r in enumerate(rec):
        v = 3 * k
        V[v, v], V[v, v + 1], V[v + 1, v + 1] = msr["term2"][r], msr["term2"][r + 1], msr["term3"][r + 1]
        V[v, v + 2], V[v + 1, v + 2], V[v + 2, v + 2] = msr["term2"][r + 2], msr["term3"][r + 2], msr["term4"][r + 2]
        for q in range(int(msr["vectorCount2"][r])):
            for x in range(3):
                cv = r + 3 + 3 * q + x
                V[v + x, v + 3 + 3 * q:v + 6 + 3 * q] = msr["term1"][cv], msr["term2"][cv], msr["term3"][cv]
    return np.triu(V) + np.triu(V, 1).T, rec


def test_cluster_partial_variance_scalars(oracle, hostsim_path):
    """phi / lambda / height variance scalars on X and Y clusters (ScaleGPSVCV_Cluster, MFN:401-438), alone and together
    with the whole-matrix scalar — for X clusters the reference then applies the whole-matrix scalar twice (while
    loading, ADJ:4358, and inside the partial scalars, ADJ:4484-4490); restated as is.  The matrix written back into
    the records is checked against a NumPy evaluation; the engine against the oracle."""
    from dynadjust_b200 import synth_terrestrial as st
    stn, msr, truth, _ = st.terrestrial_network(120, 300, 48, n_x=12, n_y=12, deflections=False)
    firsts = [i for i in range(len(msr)) if msr["measType"][i] in (b"X", b"Y") and msr["measStart"][i] == 0 and
              (i == 0 or msr["clusterID"][i] != msr["clusterID"][i - 1])]
    for n_, i in enumerate(firsts):
        cid = msr["clusterID"][i]
        sel = (msr["clusterID"] == cid) & np.isin(msr["measType"], (b"X", b"Y"))
        msr["scale1"][sel], msr["scale2"][sel], msr["scale3"][sel] = 1.7, 0.6, 2.5
        msr["scale4"][sel] = 3.0 if n_ % 2 else 1.0
    raw = msr.copy()
    s_o, m_o = stn.copy(), msr.copy()
    ref = oracle.adjust_simultaneous(s_o, m_o, want_vcv=False)
    for i in firsts:
        V, rec = _cluster_matrix(raw, i)
        kind, vs = raw["measType"][i], float(raw["scale4"][i])
        both = abs(vs - 1.0) > 1e-5
        if kind == b"X" and both:
            V = V * vs
        sc = np.sqrt(np.array([1.7, 0.6, 2.5]) * (vs if both else 1.0))
        M = np.zeros_like(V)
        for k, r in enumerate(rec):
            s = int(raw["station1"][r])
            la, lo, hh = stn["currentLatitude"][s], stn["currentLongitude"][s], stn["currentHeight"][s]
            f = lambda a, b, c: synth.geo_to_cart(np.array([a]), np.array([b]), np.array([c]))[0]
            eps = 1e-7
            J = np.stack([(f(la + eps, lo, hh) - f(la - eps, lo, hh)) / (2 * eps), (f(la, lo + eps, hh) - f(la, lo - eps, hh)) / (2 * eps),
                          (f(la, lo, hh + 1.0) - f(la, lo, hh - 1.0)) / 2.0], axis=1)
            M[3 * k:3 * k + 3, 3 * k:3 * k + 3] = J @ np.diag(sc) @ np.linalg.inv(J)
        want = M @ V @ M.T
        got, _ = _cluster_matrix(m_o, i)
        assert np.abs(got - want).max() < 2e-7 * np.abs(want).max(), (kind, vs)
    s_e, m_e = stn.copy(), msr.copy()
    adj, info, last, stats = parity.run_engine(hostsim_path, s_e, m_e, leaf_stations=16)
    assert np.abs(adj.estimates() - ref["est"]).max() < parity.TOL_XYZ
    assert abs(stats.sigma_zero - ref["res"].sigma_zero) < 1e-11
    c = np.isin(m_e["measType"], (b"X", b"Y"))
    for fld in ("term1", "term2", "term3", "term4"):
        assert np.abs(m_e[fld][c] - m_o[fld][c]).max() <= 1e-12 * max(1.0, np.abs(m_o[fld][c]).max()), fld
    adj.close()


def test_all_types_together(oracle, hostsim_path):
    """BASELINE config C3's mix and more: every type in one network, nested dissection and a chain of blocks."""
    mix = dict(scalars={k: 50 for k in "ABKCEMSVZLHRIJPQ"}, n_dir_sets=40, n_x=20, n_y=20, ignore_some=True)
    parity.check_against_oracle(oracle, hostsim_path, 300, 800, 45, terrestrial=mix, leaf_stations=24)
    parity.check_against_oracle(oracle, hostsim_path, 300, 800, 46, terrestrial=mix, blocks=lambda n: parity.chain_blocks(n, 50))


def test_normals_and_rhs(oracle, hostsim_path):
    parity.check_normals(oracle, hostsim_path, 80, 240, 4, leaf_stations=12)


def test_repeated_station_pairs(oracle, hostsim_path):
    # several baselines over the same station pair, in both directions: their off-diagonal block is accumulated
    # (atomic adds), while pairs observed once are stored — both paths must give the reference's normals
    def mutate(stn, msr):
        rec = msr.reshape(-1, 3)
        for dst, src, flip in ((5, 0, False), (9, 0, True), (30, 12, True)):
            rec["station1"][dst] = rec["station2"][src] if flip else rec["station1"][src]
            rec["station2"][dst] = rec["station1"][src] if flip else rec["station2"][src]
            rec["term1"][dst] = -rec["term1"][src] if flip else rec["term1"][src]
    parity.check_normals(oracle, hostsim_path, 80, 240, 6, mutate=mutate, leaf_stations=12)


def _check_block_vcvs(oracle, lib, **kw):
    """gadj_get_block_vcv: the dense variance matrix of every block (inner + junction stations) against the same rows /
    columns of the oracle's full inverse."""
    blocks = kw.pop("blocks", None)
    stn, msr, _, _ = synth.gnss_network(kw.pop("n", 200), kw.pop("m", 600), kw.pop("seed", 15))
    ref = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_vcv=True)
    V = ref["vcv"]
    adj, info, last, stats = parity.run_engine(lib, stn, msr, blocks=blocks(len(stn)) if blocks else None, **kw)
    seen = set()
    for b in range(info.nfronts):
        st, Q = adj.block_vcv(b)
        idx = np.concatenate([[3 * s, 3 * s + 1, 3 * s + 2] for s in st]).astype(int)
        want = V[np.ix_(idx, idx)]
        assert Q.shape == want.shape and np.abs(Q - want).max() <= 2e-8 * np.abs(want).max(), b
        seen.update(int(s) for s in st)
    assert seen == set(range(len(stn)))
    adj.close()
    return info


def test_block_variance_matrices(oracle, hostsim_path):
    info = _check_block_vcvs(oracle, hostsim_path, leaf_stations=16)
    assert info.nfronts > 8
    info = _check_block_vcvs(oracle, hostsim_path, blocks=lambda n: parity.chain_blocks(n, 40))      # .seg chain: block = front
    assert info.nfronts == 5
    info = _check_block_vcvs(oracle, hostsim_path, n=60, m=170, seed=21, ordering=engine.ORDER_DENSE)  # the whole network
    assert info.nfronts == 1


def test_small_workspace_forces_chunks(oracle, hostsim_path):
    # a tight inverse workspace makes every level run in several chunks; results must not change
    parity.check_against_oracle(oracle, hostsim_path, 400, 1200, 5, leaf_stations=8, workspace_gb=2.0e-4)


def test_singular_network_reports_reference_message(hostsim_path):
    stn, msr, _, _ = synth.gnss_network(30, 80, 3)
    msr["term2"] = 0.0  # zero variances -> V is not positive definite
    adj = engine.Adjustment(stn, msr, lib_path=hostsim_path)
    adj.prepare()
    with pytest.raises(engine.AdjustmentError) as e:
        adj.adjust()
    assert "Invalid variance matrix" in str(e.value) or "singular" in str(e.value)


def test_error_paths(hostsim_path):
    stn, msr, _, _ = synth.gnss_network(30, 80, 3)
    adj = engine.Adjustment(lib_path=hostsim_path)
    with pytest.raises(engine.AdjustmentError):
        adj.prepare()                           # nothing set
    adj.set_stations(stn)
    bad = msr.copy()
    bad["station2"][0:3] = 1000                 # beyond the station list
    adj.set_measurements(bad)
    with pytest.raises(engine.AdjustmentError, match="beyond the station list"):
        adj.prepare()
    other = msr.copy()
    other["measType"][0:3] = b"?"
    adj.set_measurements(other)
    with pytest.raises(engine.AdjustmentError, match="not a DynAdjust measurement type"):
        adj.prepare()
    from dynadjust_b200 import synth_terrestrial
    s2, m2, _, _ = synth_terrestrial.terrestrial_network(40, 100, 3, n_y=4)
    m2["coordType"][m2["measType"] == b"Y"] = b"UTM"   # a coordinate form the reference's Y clusters do not have: refused
    a2 = engine.Adjustment(s2, m2, lib_path=hostsim_path)
    with pytest.raises(engine.AdjustmentError, match="XYZ, LLH or LLh"):
        a2.prepare()
    s3, m3, _, _ = synth_terrestrial.terrestrial_network(40, 100, 3, n_y=4)
    ycl = np.where(m3["measType"] == b"Y")[0]
    m3["term2"][ycl[0]] = -1.0                      # indefinite cluster VCV -> the reference's message
    a3 = engine.Adjustment(s3, m3, lib_path=hostsim_path)
    with pytest.raises(engine.AdjustmentError, match="Invalid variance matrix"):
        a3.prepare()
    ign = msr.copy()
    ign["ignore"][0:3] = 1                      # ignored measurements are skipped like the reference's CML
    adj.set_measurements(ign)
    info = adj.prepare()
    assert info.nbaselines == len(msr) // 3 - 1
    with pytest.raises(engine.AdjustmentError):
        adj.station_vcvs()                      # inverse not formed yet


def test_header_declares_every_exported_symbol(hostsim_path):
    """include/gadj.h, the Python mirror and the library agree on the entry points."""
    hdr = open(os.path.join(ROOT, "include", "gadj.h")).read()
    declared = set(re.findall(r"\b(gadj_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(engine.EXPORTS)
    lib = ctypes.CDLL(hostsim_path)
    for name in declared:
        assert hasattr(lib, name), name


def test_product_library_exports(tmp_path):
    """The CUDA product library loads on a CPU-only box and exports the full C-ABI (no compute calls)."""
    path = engine.LIB_PATH
    if not os.path.exists(path):
        pytest.skip("libgadj.so not built yet (run __graft_entry__.build())")
    lib = ctypes.CDLL(path)
    for name in engine.EXPORTS:
        assert hasattr(lib, name), name
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    assert "DMMA" in out and "UTMALDG" in out and "UBLKCP" in out


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(engine.LIB_PATH):
        pytest.skip("libgadj.so not built yet")
    with pytest.raises(engine.AdjustmentError, match="no CPU fallback"):
        engine.Adjustment()
