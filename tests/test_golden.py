"""The reference's own end-to-end golden outputs pin the oracle and the engine (GNSS sample below; the urban sample
with every terrestrial type at the end of the file).

tests/golden/gnss_sample.npz (made by tests/golden/make_gnss_sample.py) holds the reference's sample GNSS network
(43 stations; 129 G baselines, one X cluster of 4 baselines, one Y cluster of 6 points; 417 measurement rows) as binary
records, and the numbers of sampleData/gnss.simult.adj.expected — the file the reference's CI compares its own dnaadjust
output with (dnadiff, tolerance 0.001).  Here the comparison is at the printed resolution — half a unit of the last
printed decimal (5e-5 m on the 4-decimal columns, 0.005 on the 2-decimal statistics): every printed digit of the
reference's expected solution block, 417 adjusted-measurement rows and 43 adjusted stations is reproduced.  (H(Ortho)
is not compared: it needs the geoid grid of the upstream dnageoid step.)"""
import os

import numpy as np
import pytest

from dynadjust_b200 import engine
from dynadjust_b200.records import MSR_DTYPE, STN_DTYPE
from tests import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "gnss_sample.npz"))
    g = dict(stn=np.ascontiguousarray(z["stn"].astype(STN_DTYPE)), msr=np.ascontiguousarray(z["msr"].astype(MSR_DTYPE)),
             sol=dict(zip(z["solution_keys"].tolist(), z["solution"].tolist())), msr_keys=z["msr_keys"].tolist(),
             msr_rows=z["msr_rows"], stn_names=z["stn_names"].tolist(), stn_rows=z["stn_rows"])
    assert g["stn"].dtype == STN_DTYPE and g["msr"].dtype == MSR_DTYPE
    return g


def _dms_to_rad(v):
    a = np.abs(v)
    d = np.floor(a + 1e-12)
    m = np.floor((a - d) * 100.0 + 1e-9)
    s = ((a - d) * 100.0 - m) * 100.0
    return np.sign(v) * np.radians(d + m / 60.0 + s / 3600.0)


def _check(g, stn, msr, est, vcv_of, stats, iterations):
    sol = g["sol"]
    assert iterations == sol["iterations"]
    assert stats["unknowns"] == sol["unknowns"] and stats["measurements"] == sol["measurements"] and stats["dof"] == sol["dof"]
    assert abs(stats["chi_squared"] - sol["chi_squared"]) < 0.0051
    assert abs(stats["sigma_zero"] - sol["sigma_zero"]) < 0.00051
    assert abs(stats["pelzer"] - sol["pelzer"]) < 0.00051
    assert stats["outliers"] == sol["outliers"]
    # adjusted coordinates: X Y Z, ellipsoidal height, latitude / longitude, SD(e, n, up)
    names = [n.decode() for n in stn["stationName"]]
    for name, row in zip(g["stn_names"], g["stn_rows"]):
        i = names.index(name)
        assert np.abs(est[i] - row[4:7]).max() < 0.51e-4, name
        assert abs(stn["currentHeight"][i] - row[3]) < 0.51e-4
        assert abs(stn["currentLatitude"][i] - _dms_to_rad(row[0])) < 4e-10 and abs(stn["currentLongitude"][i] - _dms_to_rad(row[1])) < 4e-10
        lat, lon = stn["currentLatitude"][i], stn["currentLongitude"][i]
        sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
        R = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])
        sd = np.sqrt(np.abs(np.diag(R.T @ vcv_of(i) @ R)))
        assert np.abs(sd - row[7:10]).max() < 0.51e-4, name
    # adjusted measurements, row by row in file order (covariance records carry no row)
    rows = msr[(msr["measStart"] <= 2) & (msr["ignore"] == 0)]
    assert len(rows) == len(g["msr_rows"]) == 417
    for m, key, want in zip(rows, g["msr_keys"], g["msr_rows"]):
        f = key.split()
        assert f[0] == chr(m["measType"][0]) and f[1] == names[m["station1"]] and f[-1] == "XYZ"[m["measStart"]], key
        var = (m["term2"], m["term3"], m["term4"])[m["measStart"]]
        got = [m["preAdjMeas"], m["measAdj"], m["measCorr"], np.sqrt(var), np.sqrt(abs(m["measAdjPrec"])), np.sqrt(m["residualPrec"])]
        assert np.abs(np.array(got) - want[:6]).max() < 0.51e-4, (key, got, want)
        assert abs(m["NStat"] - want[6]) < 0.0051 and abs(m["PelzerRel"] - want[7]) < 0.0051, key


def test_oracle_reproduces_reference_expected_output(oracle, golden):
    stn, msr = golden["stn"].copy(), golden["msr"].copy()
    ref = oracle.adjust_simultaneous(stn, msr, want_vcv=True)
    r = ref["res"]
    V = ref["vcv"]
    _check(golden, stn, msr, ref["est"].reshape(-1, 3), lambda i: V[3 * i:3 * i + 3, 3 * i:3 * i + 3],
           dict(unknowns=r.unknown_params, measurements=r.measurement_params, dof=r.dof, chi_squared=r.chi_squared,
                sigma_zero=r.sigma_zero, pelzer=r.global_pelzer, outliers=r.outliers), r.iterations)


def _engine(golden, lib, **kw):
    stn, msr = golden["stn"].copy(), golden["msr"].copy()
    adj, info, last, st = parity.run_engine(lib, stn, msr, **kw)
    q = adj.station_vcvs().reshape(-1, 3, 3)
    _check(golden, stn, msr, adj.estimates().reshape(-1, 3), lambda i: q[i],
           dict(unknowns=st.unknown_params, measurements=st.measurement_params, dof=st.dof, chi_squared=st.chi_squared,
                sigma_zero=st.sigma_zero, pelzer=st.global_pelzer, outliers=st.outliers), last.iteration)
    adj.close()


def test_engine_host_logic_reproduces_reference_expected_output(hostsim_path, golden):
    _engine(golden, hostsim_path, leaf_stations=8)
    _engine(golden, hostsim_path, ordering=engine.ORDER_DENSE)


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_expected_output(gpu_lib, golden):
    _engine(golden, gpu_lib, leaf_stations=8)
    _engine(golden, gpu_lib, ordering=engine.ORDER_DENSE)


# ---- the reference's urban sample: terrestrial types against its expected phased adjustment ---------------------
@pytest.fixture(scope="module")
def urban():
    z = np.load(os.path.join(ROOT, "tests", "golden", "urban_sample.npz"))
    return dict(stn=np.ascontiguousarray(z["stn"].astype(STN_DTYPE)), msr=np.ascontiguousarray(z["msr"].astype(MSR_DTYPE)),
                sol=dict(zip(z["solution_keys"].tolist(), z["solution"].tolist())), msr_keys=z["msr_keys"].tolist(),
                msr_rows=z["msr_rows"], stn_names=z["stn_names"].tolist(), stn_rows=z["stn_rows"])


SEC = np.radians(1.0 / 3600.0)


def _check_urban(g, stn, msr, est, vcv_of, stats):
    """tests/golden/urban_sample.npz: 149 stations (UTM, mixed constraints CCC / CCF / CFF / FFC), 1182 rows of types
    A B G H K L M S V Y Z, the Y cluster in latitude / longitude / orthometric height.  Every standard deviation the
    reference prints (measurement, adjusted measurement, correction; SD e/n/up of the stations) is reproduced at print
    resolution; stations and linear measurements agree to 0.1-0.2 mm (the geoid separations of the fixture are good to
    0.05 mm), horizontal angles to 0.005"; zenith angles over 10-100 m lines to 0.05" (their deflection corrections
    differ on a few lines — probably the phased expected run reducing a block at the coordinates carried in from earlier
    blocks; in the re-adjusted urban_mt run the same column agrees to 0.001")."""
    sol = g["sol"]
    assert stats["unknowns"] == sol["unknowns"] and stats["measurements"] == sol["measurements"] and stats["dof"] == sol["dof"]
    assert stats["outliers"] == sol["outliers"]
    assert abs(stats["chi_squared"] - sol["chi_squared"]) < 0.2            # 0.03 %
    assert abs(stats["sigma_zero"] - sol["sigma_zero"]) < 0.0011 and abs(stats["pelzer"] - sol["pelzer"]) < 0.0011
    names = [n.decode() for n in stn["stationName"]]
    for name, row in zip(g["stn_names"], g["stn_rows"]):
        i = names.index(name)
        assert np.abs(est[i] - row[4:7]).max() < 1.5e-4 and abs(stn["currentHeight"][i] - row[3]) < 1.5e-4, name
        lat, lon = stn["currentLatitude"][i], stn["currentLongitude"][i]
        sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
        R = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])
        sd = np.sqrt(np.abs(np.diag(R.T @ vcv_of(i) @ R)))
        assert np.abs(sd - row[7:10]).max() < 0.51e-4, name
    rows = msr[(msr["measStart"] <= 2) & (msr["ignore"] == 0)]
    assert len(rows) == len(g["msr_rows"]) == 1182
    seen = set()
    for m, key, w in zip(rows, g["msr_keys"], g["msr_rows"]):
        t = key[0]
        assert t == chr(m["measType"][0]) and key.split()[1] == names[m["station1"]], key
        seen.add(t)
        if t == "Y":
            continue   # printed back in geographic form (PrintAdjMeasurements_YLLH)
        ang = t in "ABKVZ"
        unit = SEC if ang else 1.0
        var = (m["term2"], m["term3"], m["term4"])[m["measStart"]] if t == "G" else m["term2"]
        sds = np.array([np.sqrt(var), np.sqrt(abs(m["measAdjPrec"])), np.sqrt(m["residualPrec"])]) / unit
        assert np.abs(sds - w[3:6]).max() < 0.51e-4 + (1.5e-4 if ang else 0.0), (key, sds, w[3:6])
        tol_adj = (0.07 if t in "VZ" else 0.01) if ang else 2e-4
        assert abs(m["measAdj"] - w[1]) / unit < tol_adj and abs(m["measCorr"] / unit - w[2]) < tol_adj, key
        assert abs(m["NStat"] - w[6]) < 0.015 and abs(m["PelzerRel"] - w[7]) < 0.011, key
        assert abs(m["preAdjCorr"] / unit - w[8]) < (0.05 if ang else 2e-4), key
    assert seen == set("ABGHKLMSVYZ")


def test_oracle_reproduces_reference_urban_output(oracle, urban):
    stn, msr = urban["stn"].copy(), urban["msr"].copy()
    ref = oracle.adjust_simultaneous(stn, msr, want_vcv=True)
    r, V = ref["res"], ref["vcv"]
    _check_urban(urban, stn, msr, ref["est"].reshape(-1, 3), lambda i: V[3 * i:3 * i + 3, 3 * i:3 * i + 3],
                 dict(unknowns=r.unknown_params, measurements=r.measurement_params, dof=r.dof, chi_squared=r.chi_squared,
                      sigma_zero=r.sigma_zero, pelzer=r.global_pelzer, outliers=r.outliers))


def _engine_urban(urban, oracle, lib, tol_sigma0, **kw):
    stn, msr = urban["stn"].copy(), urban["msr"].copy()
    adj, info, last, st = parity.run_engine(lib, stn, msr, **kw)
    q = adj.station_vcvs().reshape(-1, 3, 3)
    _check_urban(urban, stn, msr, adj.estimates().reshape(-1, 3), lambda i: q[i],
                 dict(unknowns=st.unknown_params, measurements=st.measurement_params, dof=st.dof, chi_squared=st.chi_squared,
                      sigma_zero=st.sigma_zero, pelzer=st.global_pelzer, outliers=st.outliers))
    # and against the oracle at the parity bar: every terrestrial type on the reference's own network
    s_o, m_o = urban["stn"].copy(), urban["msr"].copy()
    ref = oracle.adjust_simultaneous(s_o, m_o, want_vcv=False)
    assert np.abs(adj.estimates() - ref["est"]).max() < parity.TOL_XYZ
    assert abs(st.sigma_zero - ref["res"].sigma_zero) < tol_sigma0 and last.iteration == ref["res"].iterations
    adj.close()


def test_engine_host_logic_reproduces_reference_urban_output(oracle, hostsim_path, urban):
    _engine_urban(urban, oracle, hostsim_path, 1e-11, leaf_stations=12)


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_urban_output(oracle, gpu_lib, urban):
    # trigonometric rows: CUDA's and glibc's libm differ in the last ulp (see tests/test_gpu_parity.py)
    _engine_urban(urban, oracle, gpu_lib, 1e-7, leaf_stations=12)


# ---- the reference's urban_mt chain: data moved GDA94 -> GDA2020, adjusted, then adjusted again from the updated files ----
@pytest.fixture(scope="module")
def urban_mt():
    z = np.load(os.path.join(ROOT, "tests", "golden", "urban_mt_sample.npz"))
    return dict(stn=np.ascontiguousarray(z["stn"].astype(STN_DTYPE)), msr=np.ascontiguousarray(z["msr"].astype(MSR_DTYPE)),
                sol=dict(zip(z["solution_keys"].tolist(), z["solution"].tolist())), msr_keys=z["msr_keys"].tolist(),
                msr_rows=z["msr_rows"], stn_names=z["stn_names"].tolist(), stn_rows=z["stn_rows"])


def _engine_urban_mt(g, lib):
    """tests/golden/urban_mt_sample.npz (made by make_urban_mt_sample.py): sampleData/urban_mt.phased-mt.adj.expected is the
    output of the SECOND dnaadjust of the reference's CI chain (CMakeLists.txt:1081-1083) — the first runs with
    --free-stn-sd 4.0 --fixed-stn-sd 0.000001 --max-iterations 20 and updates the binary files, the second starts from them
    with the defaults.  Reproducing it needs the re-adjustment semantics of reduced files: measured values restored from
    preAdjMeas, reductions and scalars not applied twice, the converted point cluster left as it is."""
    stn, msr = g["stn"].copy(), g["msr"].copy()
    first = engine.Adjustment(stn, msr, lib_path=lib, leaf_stations=12, free_std_dev=4.0, fixed_std_dev=1.0e-6, max_iterations=20)
    first.prepare()
    first.adjust()
    first.statistics(write_back=True)          # adjusted coordinates and statistics into the records, as UpdateBinaryFiles leaves them
    first.close()
    second = engine.Adjustment(lib_path=lib, leaf_stations=12)
    second.set_stations(stn)
    second.set_measurements(msr, reduced=True)
    second.prepare()
    last = second.adjust()
    st = second.statistics(write_back=True)
    assert last.iteration == 1                 # the expected file shows one iteration with corrections of 1e-8 m
    q = second.station_vcvs().reshape(-1, 3, 3)
    _check_urban(g, stn, msr, second.estimates().reshape(-1, 3), lambda i: q[i],
                 dict(unknowns=st.unknown_params, measurements=st.measurement_params, dof=st.dof, chi_squared=st.chi_squared,
                      sigma_zero=st.sigma_zero, pelzer=st.global_pelzer, outliers=st.outliers))
    second.close()


def test_engine_host_logic_reproduces_reference_readjusted_urban_output(hostsim_path, urban_mt):
    _engine_urban_mt(urban_mt, hostsim_path)


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_readjusted_urban_output(gpu_lib, urban_mt):
    _engine_urban_mt(urban_mt, gpu_lib)
