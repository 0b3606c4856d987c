"""The reporting entry points of the C-ABI (bulk pair variances, ignored / computed measurements evaluated at the
estimates, re-adjustment of files an earlier run has reduced).  Host logic on the CPU stand-in; the same checks on the
CUDA library under `-m gpu`."""
import numpy as np
import pytest

from dynadjust_b200 import engine, synth
from dynadjust_b200 import synth_terrestrial as st
from tests import parity
from tests.test_cli import cli_hostsim  # noqa: F401  (fixture)

SEC = np.radians(1 / 3600.0)


def _slope(T, xyz, s1, s2, ih, th):
    """slope distance instrument -> target; both heights act along the vertical of station 1 (the adjustment's model)"""
    return np.linalg.norm(xyz[s2] - xyz[s1] + st._up(T.lat, T.lon, s1) * (th - ih)[:, None], axis=1)


def _bulk_pairs(lib, oracle):
    stn, msr, _, _ = synth.gnss_network(80, 240, 17)
    ref = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_vcv=True)
    V = ref["vcv"]
    adj, _, _, _ = parity.run_engine(lib, stn, msr, leaf_stations=10)
    first = msr[msr["measStart"] == 0]
    si, sj = first["station1"].astype(np.uint32), first["station2"].astype(np.uint32)
    si, sj = np.concatenate([si, sj[:5], si[:5]]), np.concatenate([sj, si[:5], si[:5]])     # + transposed pairs, + diagonal blocks
    Q = adj.pair_vcvs(si, sj)
    for p in range(len(si)):
        want = V[3 * si[p]:3 * si[p] + 3, 3 * sj[p]:3 * sj[p] + 3]
        assert np.abs(Q[p] - want).max() <= 2e-8 * np.abs(V).max(), p
        if p % 40 == 0:
            assert np.array_equal(Q[p], adj.vcv_block(int(si[p]), int(sj[p])))
    far = next(j for j in range(1, 80) if not np.any(((si == 0) & (sj == j)) | ((si == j) & (sj == 0))))
    with pytest.raises(engine.AdjustmentError, match="outside the stored pattern"):
        adj.pair_vcvs([0], [far])
    adj.close()


def test_bulk_pair_variances(hostsim_path, oracle):
    _bulk_pairs(hostsim_path, oracle)


@pytest.mark.gpu
def test_bulk_pair_variances_gpu(gpu_lib, oracle):
    _bulk_pairs(gpu_lib, oracle)


def _ignored(lib):
    """Ignored measurements a posteriori (ADJ:8750-9980): computed from the adjusted coordinates in the domain they were
    observed in — baselines as coordinate differences, slope distances incl. instrument / target heights, level differences
    and orthometric heights with the geoid separations removed, directions as derived angles."""
    stn, msr, truth, _ = st.terrestrial_network(60, 170, 29, scalars={"S": 12, "L": 10, "H": 6, "R": 4, "E": 6, "M": 6, "C": 4}, n_dir_sets=3, n_x=1, n_y=1,
                                                deflections=False)
    pick = {}
    for t in "SLHREMC":
        pick[t] = int(np.where(msr["measType"] == t.encode())[0][1])
        msr["ignore"][pick[t]] = 1
    g = int(np.where(msr["measType"] == b"G")[0][3])
    msr["ignore"][g:g + 3] = 1
    d0 = int(np.where((msr["measType"] == b"D") & (msr["vectorCount1"] > 0))[0][0])
    nd = int(msr["vectorCount1"][d0])
    msr["ignore"][d0:d0 + nd] = 1
    raw = msr.copy()
    adj, _, _, _ = parity.run_engine(lib, stn, msr, leaf_stations=12)
    adj.update_ignored_measurements()
    est = adj.estimates()
    llh = np.stack(synth.cart_to_geo(est), axis=1)
    N = stn["geoidSep"].astype(float)
    a, b = raw["station1"], raw["station2"]
    # a baseline: station 2 - station 1, component by component
    for q in range(3):
        assert abs(msr["measAdj"][g + q] - (est[b[g], q] - est[a[g], q])) < 1e-9 and msr["preAdjMeas"][g + q] == raw["term1"][g + q]
        assert abs(msr["measCorr"][g + q] - (msr["measAdj"][g + q] - raw["term1"][g + q])) < 1e-12
    T = st.Truth(stn, est)
    i = pick["C"]
    assert abs(msr["measAdj"][i] - T.chord(a[[i]], b[[i]])[0]) < 1e-6                      # chord between the ellipsoid feet
    i = pick["S"]
    want = _slope(T, est, a[[i]], b[[i]], raw["term3"][[i]], raw["term4"][[i]])[0]
    assert abs(msr["measAdj"][i] - want) < 1e-7 and abs(msr["measCorr"][i] - (want - raw["term1"][i])) < 1e-7
    i = pick["R"]
    assert abs(msr["measAdj"][i] - llh[a[i], 2]) < 1e-6 and msr["preAdjCorr"][i] == 0.0
    i = pick["H"]
    assert abs(msr["measAdj"][i] - (llh[a[i], 2] - N[a[i]])) < 1e-6 and abs(msr["preAdjCorr"][i] - N[a[i]]) < 1e-6
    i = pick["L"]
    want = (llh[b[i], 2] - N[b[i]]) - (llh[a[i], 2] - N[a[i]])
    assert abs(msr["measAdj"][i] - want) < 1e-6 and abs(msr["preAdjCorr"][i] - (N[b[i]] - N[a[i]])) < 1e-6
    i = pick["E"]
    assert abs(msr["measAdj"][i] - T.ell_arc(a[[i]], b[[i]])[0]) < 1e-5 and abs(msr["measCorr"][i]) < 0.05
    i = pick["M"]
    assert abs(msr["measAdj"][i] - T.msl_arc(a[[i]], b[[i]])[0]) < 1e-5 and abs(msr["measCorr"][i]) < 0.05
    # a direction set: derived angles between successive directions, corrected like an observed angle
    for k in range(1, nd):
        r = d0 + k
        az1 = T.azimuth(a[[r - 1]], b[[r - 1]])[0]
        az2 = T.azimuth(a[[r]], b[[r]])[0]
        ang = (az2 - az1) % (2 * np.pi)
        assert abs(msr["scale1"][r] - (raw["term1"][r] - raw["term1"][r - 1]) % (2 * np.pi)) < 1e-12
        assert abs((msr["measCorr"][r] - (ang - msr["scale1"][r]) + np.pi) % (2 * np.pi) - np.pi) < 2e-9, k
        assert abs(msr["scale2"][r] - (raw["term2"][r] + raw["term2"][r - 1])) < 1e-20
    # the records that take part were not touched by the call, the raw values of the ignored ones are kept
    assert np.array_equal(msr["term1"][msr["ignore"] == 1], raw["term1"][raw["ignore"] == 1])
    adj.close()


def test_ignored_measurements(hostsim_path):
    _ignored(hostsim_path)


@pytest.mark.gpu
def test_ignored_measurements_gpu(gpu_lib):
    _ignored(gpu_lib)


def test_computed_measurements_before_the_first_iteration(hostsim_path):
    """gadj_compute_measurements at the a-priori coordinates: computed - measured is minus the right-hand-side residual
    the first iteration starts from (PrintCompMeasurements "a-priori", ADJ:2443-2445)."""
    stn, msr, truth, _ = st.terrestrial_network(40, 110, 31, scalars={"S": 10, "L": 6, "A": 6}, deflections=False)
    adj = engine.Adjustment(stn, msr, lib_path=hostsim_path, leaf_stations=12)
    adj.prepare()
    adj._check(adj.L.gadj_compute_measurements(adj.h))
    x0 = synth.geo_to_cart(stn["initialLatitude"], stn["initialLongitude"], stn["initialHeight"])
    g = np.where((msr["measType"] == b"G") & (msr["measStart"] == 0))[0]
    for q in range(3):
        assert np.abs(msr["measAdj"][g + q] - (x0[msr["station2"][g], q] - x0[msr["station1"][g], q])).max() < 1e-8
    s = np.where(msr["measType"] == b"S")[0]
    T = st.Truth(stn, x0)
    want = _slope(T, x0, msr["station1"][s], msr["station2"][s], msr["term3"][s], msr["term4"][s])
    assert np.abs(msr["measAdj"][s] - want).max() < 1e-7 and np.abs(msr["measCorr"][s] - (want - msr["term1"][s])).max() < 1e-7
    assert np.abs(msr["measCorr"][s]).max() > 0.05          # half-metre a-priori errors show up as differences
    last = adj.adjust()                                     # and the adjustment is not disturbed by the call
    stn2, msr2, _, _ = st.terrestrial_network(40, 110, 31, scalars={"S": 10, "L": 6, "A": 6}, deflections=False)
    adj2, _, last2, _ = parity.run_engine(hostsim_path, stn2, msr2, leaf_stations=12)
    assert np.abs(adj.estimates() - adj2.estimates()).max() < 1e-9 and last.iteration == last2.iteration
    adj.close(), adj2.close()


def _readjust(lib, tol):
    """A second adjustment on records an earlier run reduced and updated (metadata `reduced`, ADJ:296, 3913-3935): measured
    values come back from preAdjMeas, geoid reductions and variance scalars are not applied a second time."""
    kw = dict(scalars={"S": 14, "L": 10, "H": 6, "E": 5, "M": 5}, n_x=2, n_y=1, v_scale=1.6, deflections=False)
    stn, msr, _, _ = st.terrestrial_network(50, 150, 37, **kw)
    g = np.where(msr["measType"] == b"G")[0]
    msr["scale4"][g[:45]] = 3.0
    adj, _, _, s1 = parity.run_engine(lib, stn, msr, leaf_stations=12)
    e1 = adj.estimates().copy()
    adj.close()
    assert np.abs(msr["preAdjCorr"][msr["measType"] == b"L"]).max() > 1e-3          # the levelled differences were reduced
    # same records, now as an updated file would hand them over: stations at the adjusted coordinates
    again = engine.Adjustment(lib_path=lib, leaf_stations=12)
    again.set_stations(stn)
    again.set_measurements(msr, reduced=True)
    again.prepare()
    last = again.adjust()
    s2 = again.statistics(write_back=True)
    assert last.iteration == 1                                                           # already at the solution
    assert np.abs(again.estimates() - e1).max() < tol and abs(s2.sigma_zero - s1.sigma_zero) < 1e-5 * s1.sigma_zero
    again.close()
    # without the flag the reductions and scalars would be applied twice: a different answer
    stn3, msr3 = stn.copy(), msr.copy()
    wrong = engine.Adjustment(stn3, msr3, lib_path=lib, leaf_stations=12)
    wrong.prepare()
    wrong.adjust()
    s3 = wrong.statistics(write_back=False)
    assert abs(s3.sigma_zero - s1.sigma_zero) > 0.05 * s1.sigma_zero
    wrong.close()


def test_readjustment_of_reduced_records(hostsim_path):
    _readjust(hostsim_path, 1e-5)      # one more Gauss-Newton step than the first run took


@pytest.mark.gpu
def test_readjustment_of_reduced_records_gpu(gpu_lib):
    _readjust(gpu_lib, 1e-5)


def test_station_without_measurements_is_not_an_unknown(hostsim_path, oracle, cli_hostsim, tmp_path):
    """A station no measurement (that takes part) touches is not in the reference's station lists (network_data_loader.cpp:
    286-300): it does not count as three unknowns, and the reports leave it out."""
    import os
    import re
    from dynadjust_b200 import dnafiles
    from tests.test_cli import _run, _write_network
    stn, msr, _, _ = synth.gnss_network(30, 85, 13)
    lone = 17
    touching = np.isin(msr["station1"], [lone]) | np.isin(msr["station2"], [lone])
    for i in np.where(touching & (msr["measStart"] == 0))[0]:
        msr["ignore"][i:i + 3] = 1                       # every baseline at that station is flagged as ignored
    ref = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_vcv=False)
    adj, _, _, st = parity.run_engine(hostsim_path, stn.copy(), msr.copy(), leaf_stations=8)
    assert st.unknown_params == ref["res"].unknown_params == 3 * 29 - 9 and st.dof == ref["res"].dof
    assert abs(st.sigma_zero - ref["res"].sigma_zero) < 1e-10
    adj.close()
    _write_network(tmp_path, "un", stn, msr)
    r = _run(cli_hostsim, tmp_path, "un", "--output-pos-uncertainty", "--output-corrections-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    name = stn["stationName"][lone].decode()
    for ext in ("adj", "xyz", "apu", "cor"):
        text = open(os.path.join(tmp_path, "un.simult." + ext)).read()
        assert not re.search(r"^" + name + r"\s", text, re.M), ext
        assert re.search(r"^" + stn["stationName"][3].decode() + r"\s", text, re.M), ext
