"""`dnaadjust <network> [options]` — the reference's process boundary (dnaadjustwrapper.cpp:799-1467).

Network files (.bst/.bms/.seg in the reference's binary/ASCII layouts) are written by dynadjust_b200.dnafiles,
the command line is run, and the text/binary outputs are compared with the CPU oracle at dnadiff's tolerance
(0.001 on numeric fields; the tables print 4 decimals).  CPU: the command line linked against tests/hostsim;
`-m gpu`: the product binary dynadjust_b200/bin/dnaadjust on the device."""
import os
import re
import subprocess

import numpy as np
import pytest

from dynadjust_b200 import dnafiles, synth
from dynadjust_b200.records import MSR_DTYPE, STN_DTYPE
from tests import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def cli_hostsim(hostsim_path):
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.run(["make", "-s", "-C", d, "all"], check=True)
    return os.path.join(d, "_build", "dnaadjust_hostsim")


@pytest.fixture(scope="session")
def cli_gpu(gpu_lib):
    exe = os.path.join(ROOT, "dynadjust_b200", "bin", "dnaadjust")
    if not os.path.exists(exe):
        raise RuntimeError("dynadjust_b200/bin/dnaadjust is missing on a GPU box: run __graft_entry__.build()")
    return exe


def _write_network(tmp, name, stn, msr):
    dnafiles.write_bst(os.path.join(tmp, name + ".bst"), stn)
    dnafiles.write_bms(os.path.join(tmp, name + ".bms"), msr)


def _run(exe, tmp, *args):
    return subprocess.run([exe, *args, "--input-folder", str(tmp), "--output-folder", str(tmp)], capture_output=True, text=True,
                          timeout=600)


def _solution_block(text):
    def grab(label, cast=float):
        m = re.search(r"^" + re.escape(label) + r"\s+(\S+)", text, re.M)
        assert m, label
        return cast(m.group(1))
    return dict(unknowns=grab("Number of unknown parameters", int), measurements=grab("Number of measurements", int),
                dof=grab("Degrees of freedom", int), chi=grab("Chi squared"), sigma0=grab("Rigorous Sigma Zero"),
                pelzer=grab("Global (Pelzer) Reliability"), iterations=len(re.findall(r"^ITERATION\s+\d+", text, re.M)),
                outliers=int((re.search(r"\((\d+) potential outliers?\)", text) or [0, 0])[1]),
                converged=bool(re.search(r"^SOLUTION\s+Converged", text, re.M)))


def _station_table(text):
    body = text.split("Adjusted Coordinates")[1]
    rows = {}
    for line in body.splitlines():
        f = line.split()
        if len(f) >= 12 and re.fullmatch(r"[CF]{3}", f[1]):
            rows[f[0]] = [float(x) for x in f[2:12]]
    return rows


def _check_outputs(oracle, tmp, name, suffix, stn, msr, with_msr_table):
    stn_o, msr_o = stn.copy(), msr.copy()
    ref = oracle.adjust_simultaneous(stn_o, msr_o, want_vcv=True)
    rr = ref["res"]
    adj = open(os.path.join(tmp, f"{name}.{suffix}.adj")).read()
    xyz = open(os.path.join(tmp, f"{name}.{suffix}.xyz")).read()
    sol = _solution_block(adj)
    assert sol["converged"] and sol["iterations"] == rr.iterations
    assert sol["unknowns"] == rr.unknown_params and sol["measurements"] == rr.measurement_params and sol["dof"] == rr.dof
    assert abs(sol["chi"] - rr.chi_squared) < 0.006 and abs(sol["sigma0"] - rr.sigma_zero) < 0.0006
    assert abs(sol["pelzer"] - rr.global_pelzer) < 0.0006 and sol["outliers"] == rr.outliers
    # adjusted coordinate table (.adj and .xyz carry the same table): X Y Z, h, SDs to the printed 4 decimals
    V = ref["vcv"]
    for text in (adj, xyz):
        rows = _station_table(text)
        assert len(rows) == len(stn)
        for i in range(len(stn)):
            r = rows[stn["stationName"][i].decode()]
            assert np.abs(np.array(r[4:7]) - ref["est"].reshape(-1, 3)[i]).max() < 1e-3
            assert abs(r[3] - stn_o["currentHeight"][i]) < 1e-3
            lat, lon = stn_o["currentLatitude"][i], stn_o["currentLongitude"][i]
            sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
            R = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])
            q = R.T @ V[3 * i:3 * i + 3, 3 * i:3 * i + 3] @ R
            sd = np.sqrt(np.abs(np.diag(q)) + np.array([0, 0, float(stn["geoidSepUnc"][i]) ** 2]))
            assert np.abs(np.array(r[7:10]) - sd).max() < 1e-3
    if with_msr_table:
        body = adj.split("Adjusted Measurements")[1].split("Adjusted Coordinates")[0]
        lines = [l for l in body.splitlines() if re.match(r"^[A-Z] \S", l) and not l.startswith("M Station 1")]
        live = msr_o[msr_o["ignore"] == 0]
        assert len(lines) == len(live)
        for l, m in list(zip(lines, live))[::7]:
            f = l.split()
            nums = [float(x) for x in f if re.fullmatch(r"-?\d+\.\d+", x)]
            assert abs(nums[1] - m["measAdj"]) < 1e-3 and abs(nums[2] - m["measCorr"]) < 1e-3
    return stn_o, msr_o


def _check_binary_update(tmp, name, stn_o, msr_o):
    stn2, meta_s = dnafiles.read_binary(os.path.join(tmp, name + ".bst"), STN_DTYPE)
    msr2, meta_m = dnafiles.read_binary(os.path.join(tmp, name + ".bms"), MSR_DTYPE)
    assert meta_s["reduced"] and meta_m["reduced"]          # ADJ:445-470
    assert np.abs(stn2["currentLatitude"] - stn_o["currentLatitude"]).max() < 1e-14
    assert np.abs(stn2["currentHeight"] - stn_o["currentHeight"]).max() < 1e-7
    for f in ("measAdj", "measCorr"):
        assert np.abs(msr2[f] - msr_o[f]).max() < 5e-9


def _simultaneous(exe, oracle, tmp_path):
    stn, msr, _, _ = synth.config_network("C1")
    _write_network(tmp_path, "c1", stn, msr)
    r = _run(exe, tmp_path, "c1", "--output-adj-msr")
    assert r.returncode == 0, r.stderr
    stn_o, msr_o = _check_outputs(oracle, tmp_path, "c1", "simult", stn, msr, True)
    _check_binary_update(tmp_path, "c1", stn_o, msr_o)


def _phased(exe, oracle, tmp_path):
    stn, msr, _, _ = synth.gnss_network(300, 900, 9)
    _write_network(tmp_path, "net", stn, msr)
    isl = parity.chain_blocks(300, 40)
    dnafiles.write_seg(os.path.join(tmp_path, "net.seg"), isl, [[] for _ in isl], [[] for _ in isl])
    r = _run(exe, tmp_path, "net", "--phased", "--no-binary-update")     # unambiguous prefix, as the reference's CI uses
    assert r.returncode == 0, r.stderr
    _check_outputs(oracle, tmp_path, "net", "phased", stn, msr, False)
    stn2, meta = dnafiles.read_binary(os.path.join(tmp_path, "net.bst"), STN_DTYPE)
    assert not meta["reduced"]


def _dms_to_rad(v):
    """ddd.mmssss (the reference's printed angle format) -> radians"""
    a = abs(v)
    d = np.floor(a + 1e-9)
    m = np.floor((a - d) * 100 + 1e-7)
    sec = ((a - d) * 100 - m) * 100
    return np.sign(v) * np.radians(d + m / 60 + sec / 3600)


def _apu_cor(exe, oracle, tmp_path):
    """--output-pos-uncertainty / --output-corrections-file: the .apu and .cor tables against the
    same quantities computed here from the oracle's VCV and estimates (PRN:4326-4432, PRN:4146-4230)."""
    stn, msr, _, _ = synth.gnss_network(60, 170, 33)
    _write_network(tmp_path, "pu", stn, msr)
    r = _run(exe, tmp_path, "pu", "--output-pos-uncertainty", "--output-corrections-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    stn_o, msr_o = stn.copy(), msr.copy()
    ref = oracle.adjust_simultaneous(stn_o, msr_o, want_vcv=True)
    V, est = ref["vcv"], ref["est"].reshape(-1, 3)
    apu = open(os.path.join(tmp_path, "pu.simult.apu")).read()
    assert "DYNADJUST POSITIONAL UNCERTAINTY OUTPUT FILE" in apu and "Variance matrix units:" in apu
    body = apu.split("Positional uncertainty of adjusted station coordinates")[1].splitlines()
    rows = [l.split() for l in body if l[:1].isalnum() and not l.startswith("Station")]
    assert len(rows) == len(stn)
    x0 = synth.geo_to_cart(stn["currentLatitude"], stn["currentLongitude"], stn["currentHeight"])
    for i, f in enumerate(rows):
        assert f[0] == stn["stationName"][i].decode()
        lat, lon = stn_o["currentLatitude"][i], stn_o["currentLongitude"][i]
        assert abs(_dms_to_rad(float(f[1])) - lat) < 1e-9 and abs(_dms_to_rad(float(f[2])) - lon) < 1e-9
        sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
        R = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])
        q = V[3 * i:3 * i + 3, 3 * i:3 * i + 3]
        ql = R.T @ q @ R
        ev = np.linalg.eigvalsh(ql[:2, :2])
        smaj, smin = np.sqrt(ev[1]), np.sqrt(ev[0])
        c = smin / smaj
        hz = smaj * (1.96079 + 0.004071 * c + 0.114276 * c * c + 0.371625 * c ** 3)
        vt = 1.96 * np.sqrt(ql[2, 2])
        got = [float(x) for x in f[3:7]]
        assert np.abs(np.array(got) - [hz, vt, smaj, smin]).max() < 1e-4
        # orientation of the semi-major axis: the eigenvector of the e/n block, as an azimuth from north
        w, vec = np.linalg.eigh(ql[:2, :2])
        az = np.arctan2(vec[0, 1], vec[1, 1]) % np.pi
        assert min(abs(_dms_to_rad(float(f[7])) % np.pi - az), np.pi - abs(_dms_to_rad(float(f[7])) % np.pi - az)) < 2e-3 or c > 0.98
        assert np.abs(np.array([float(x) for x in f[8:11]]) - q[0]).max() <= 1e-8 * np.abs(q).max()
    cor = open(os.path.join(tmp_path, "pu.simult.cor")).read()
    assert "DYNADJUST CORRECTIONS OUTPUT FILE" in cor
    crows = [l.split() for l in cor.split("Corrections to stations")[1].splitlines() if l[:1].isalnum() and not l.startswith("Station")]
    assert len(crows) == len(stn)
    for i, f in enumerate(crows):
        lat, lon = stn_o["currentLatitude"][i], stn_o["currentLongitude"][i]
        sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
        R = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])
        enu = R.T @ (est[i] - x0[i])
        assert np.abs(np.array([float(x) for x in f[-3:]]) - enu).max() < 1e-4        # east north up
        assert abs(float(f[-5]) - np.linalg.norm(enu)) < 1e-4                          # slope distance
        assert abs(float(f[-4]) - np.hypot(enu[0], enu[1])) < 1e-4                     # horizontal distance
        if np.hypot(enu[0], enu[1]) > 1e-3:
            az = np.degrees(np.arctan2(enu[0], enu[1]) % (2 * np.pi))
            d, m, sec = (float(x) for x in f[1:4])                                     # "ddd mm ss"
            assert abs((d + m / 60 + sec / 3600) - az) < 1.5 / 3600 or abs((d + m / 60 + sec / 3600) - az) > 359.99
    # thresholds drop the small corrections (PRN:4172-4181)
    r = _run(exe, tmp_path, "pu", "--output-corrections-file", "--hz-corr-threshold", "1000", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    cor = open(os.path.join(tmp_path, "pu.simult.cor")).read()
    assert not [l for l in cor.split("Corrections to stations")[1].splitlines() if l[:1].isalnum() and not l.startswith("Station")]
    # --output-all-covariances (PRN:4438-4484): after every station its covariance blocks with the stations that follow
    r = _run(exe, tmp_path, "pu", "--output-pos-uncertainty", "--output-all-covariances", "--output-apu-vcv-units", "ENU", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    apu = open(os.path.join(tmp_path, "pu.simult.apu")).read()
    assert re.search(r"Full covariance matrix:\s+Yes", apu)
    body = apu.split("Positional uncertainty of adjusted station coordinates")[1].splitlines()
    n, cur, seen = len(stn), -1, 0
    names = [s.decode() for s in stn["stationName"]]
    for k, l in enumerate(body):
        f = l.split()
        if len(f) == 11 and f[0] in names:                   # station row
            cur = names.index(f[0])
        elif len(f) == 4 and f[0] in names and cur >= 0:     # first row of a covariance block, then two more rows
            j = names.index(f[0])
            assert j > cur
            C = np.array([[float(x) for x in f[1:4]], [float(x) for x in body[k + 1].split()], [float(x) for x in body[k + 2].split()]])
            lat, lon = stn_o["currentLatitude"][cur], stn_o["currentLongitude"][cur]
            sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
            R = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])
            want = R.T @ V[3 * cur:3 * cur + 3, 3 * j:3 * j + 3] @ R
            assert np.abs(C - want).max() <= 1e-8 * np.abs(V).max() + 1e-9 * np.abs(want).max()
            seen += 1
    assert seen == n * (n - 1) // 2


def _constraints(exe, oracle, tmp_path):
    """--constraints "name,CCC,..." (network_data_loader.cpp:211-263): the override reaches the solve; bad names and
    bad codes are errors with the reference's wording."""
    stn, msr, _, _ = synth.gnss_network(60, 170, 35)
    _write_network(tmp_path, "cn", stn, msr)
    n1, n2 = stn["stationName"][7].decode(), stn["stationName"][11].decode()
    r = _run(exe, tmp_path, "cn", "--constraints", f"{n1},ccc,{n2},FFC", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    stn_c = stn.copy()
    stn_c["stationConst"][7] = b"CCC"
    stn_c["stationConst"][11] = b"FFC"
    _check_outputs(oracle, tmp_path, "cn", "simult", stn_c, msr, False)
    adj = open(os.path.join(tmp_path, "cn.simult.adj")).read()
    assert "Station constraints:" in adj
    rows = _station_table(adj)
    r = _run(exe, tmp_path, "cn", "--constraints", "NOSUCH,CCC")
    assert r.returncode == 1 and "is not in the stations map" in r.stderr
    r = _run(exe, tmp_path, "cn", "--constraints", f"{n1},CXC")
    assert r.returncode == 1 and "Invalid station constraint" in r.stderr
    # discontinuity sites (LDR:314-359): dnaimport renamed two stations of a discontinuity file; a constraint given for the
    # original name reaches the renamed sites, and the original name itself (no longer a station) is passed over
    stn_d = stn.copy()
    stn_d["stationNameOrig"][20] = stn_d["stationNameOrig"][21] = b"SITE"
    stn_d["stationName"][20], stn_d["stationName"][21] = b"SITE_20100101", b"SITE_20150101"
    _write_network(tmp_path, "cd", stn_d, msr)
    r = _run(exe, tmp_path, "cd", "--constraints", f"SITE,CCF,{n1},CCC", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    stn_c = stn_d.copy()
    stn_c["stationConst"][[20, 21]] = b"CCF"
    stn_c["stationConst"][7] = b"CCC"
    _check_outputs(oracle, tmp_path, "cd", "simult", stn_c, msr, False)


def test_cli_constraints_hostsim(cli_hostsim, oracle, tmp_path):
    _constraints(cli_hostsim, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_constraints_gpu(cli_gpu, oracle, tmp_path):
    _constraints(cli_gpu, oracle, tmp_path)


def _block_outputs(exe, oracle, tmp_path):
    """--output-stn-blocks (one station table per .seg block) and --block1-phased (only block 1 is reported, no global
    test; PRN:535-595, ADJ:7140-7147).  Every block is rigorous here, so the printed values are the simultaneous ones."""
    stn, msr, _, _ = synth.gnss_network(120, 360, 12)
    _write_network(tmp_path, "blk", stn, msr)
    isl = parity.chain_blocks(120, 30)
    jsl = [list(range(b[-1] + 1, min(b[-1] + 4, 120))) for b in isl]      # a few stations of the next block as junctions
    dnafiles.write_seg(os.path.join(tmp_path, "blk.seg"), isl, jsl, [[] for _ in isl])
    ref = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_vcv=True)
    est = ref["est"].reshape(-1, 3)
    r = _run(exe, tmp_path, "blk", "--phased", "--output-stn-blocks", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    adj = open(os.path.join(tmp_path, "blk.phased.adj")).read()
    parts = re.split(r"^Block (\d+)$", adj.split("SOLUTION")[1], flags=re.M)
    assert len(parts) == 2 * len(isl) + 1
    for b in range(len(isl)):
        assert int(parts[2 * b + 1]) == b + 1
        rows = {}
        for line in parts[2 * b + 2].splitlines():
            f = line.split()
            if len(f) >= 12 and re.fullmatch(r"[CF]{3}", f[1]):
                rows[f[0]] = [float(x) for x in f[2:12]]
        want = sorted(set(isl[b]) | set(jsl[b]))
        assert sorted(rows) == sorted(stn["stationName"][i].decode() for i in want)
        for i in want:
            assert np.abs(np.array(rows[stn["stationName"][i].decode()][4:7]) - est[i]).max() < 1e-3
    # --output-msr-blocks: every measurement appears once, under the block whose CML lists it
    rec = msr.reshape(-1, 3)
    blk_of = lambda s: next(b for b, l in enumerate(isl) if s in l)
    cml = [[] for _ in isl]
    for b_ in range(len(rec)):
        cml[max(blk_of(int(rec["station1"][b_, 0])), blk_of(int(rec["station2"][b_, 0])))].append(3 * b_)
    dnafiles.write_seg(os.path.join(tmp_path, "blk.seg"), isl, jsl, cml)
    r = _run(exe, tmp_path, "blk", "--phased", "--output-adj-msr", "--output-msr-blocks", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    adj = open(os.path.join(tmp_path, "blk.phased.adj")).read()
    parts = re.split(r"^Block (\d+)$", adj.split("SOLUTION")[1].split("Adjusted Coordinates")[0], flags=re.M)
    counts = [len([l for l in parts[2 * b_ + 2].splitlines() if l.startswith("G ")]) for b_ in range(len(isl))]
    assert counts == [3 * len(c) for c in cml] and sum(counts) == len(msr)
    r = _run(exe, tmp_path, "blk", "--block1-phased", "--output-adj-msr", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    adj = open(os.path.join(tmp_path, "blk.phased-block1.adj")).read()
    assert len([l for l in adj.splitlines() if l.startswith("G ")]) == 3 * len(cml[0])
    r = _run(exe, tmp_path, "blk", "--block1-phased", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    adj = open(os.path.join(tmp_path, "blk.phased-block1.adj")).read()
    assert "Chi-Square test" not in adj and "Rigorous Sigma Zero" in adj
    rows = _station_table(adj)
    assert sorted(rows) == sorted(stn["stationName"][i].decode() for i in set(isl[0]) | set(jsl[0]))


def test_cli_block_outputs_hostsim(cli_hostsim, oracle, tmp_path):
    _block_outputs(cli_hostsim, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_block_outputs_gpu(cli_gpu, oracle, tmp_path):
    _block_outputs(cli_gpu, oracle, tmp_path)


def _golden_gnss_text(exe, tmp_path):
    """The reference's CI check, on our command line: run `dnaadjust gnss --output-adj-msr` on the reference's sample
    GNSS network (tests/golden/gnss_sample.npz) and compare the .adj text with the reference's expected file
    (dnadiff compares numeric fields at 0.001; here 1.1e-4 = print rounding on both sides)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "gnss_sample.npz"))
    stn, msr = np.ascontiguousarray(z["stn"].astype(STN_DTYPE)), np.ascontiguousarray(z["msr"].astype(MSR_DTYPE))
    _write_network(tmp_path, "gnss", stn, msr)
    r = _run(exe, tmp_path, "gnss", "--output-adj-msr", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    adj = open(os.path.join(tmp_path, "gnss.simult.adj")).read()
    sol, want = _solution_block(adj), dict(zip(z["solution_keys"].tolist(), z["solution"].tolist()))
    assert sol["unknowns"] == want["unknowns"] and sol["measurements"] == want["measurements"] and sol["dof"] == want["dof"]
    assert sol["chi"] == want["chi_squared"] and sol["sigma0"] == want["sigma_zero"] and sol["pelzer"] == want["pelzer"]
    assert sol["outliers"] == want["outliers"] and sol["iterations"] == want["iterations"] and sol["converged"]
    m = re.search(r"Chi-Square test \(95.0%\)\s+(\S+) < \S+ < (\S+)", adj)
    assert abs(float(m.group(1)) - want["chi_lower"]) < 0.0011 and abs(float(m.group(2)) - want["chi_upper"]) < 0.0011
    body = adj.split("Adjusted Measurements")[1].split("Adjusted Coordinates")[0]
    lines = [l for l in body.splitlines() if re.match(r"^[GXY] \S", l)]
    assert len(lines) == len(z["msr_rows"]) == 417
    for l, key, row in zip(lines, z["msr_keys"].tolist(), z["msr_rows"]):
        f = l.split()
        k = next(i for i in range(2, len(f)) if f[i] in ("X", "Y", "Z") and re.fullmatch(r"-?\d+\.\d+", f[i + 1]))
        assert " ".join([f[0]] + f[1:k] + [f[k]]) == key
        got = np.array([float(x) for x in f[k + 1:k + 10]])
        assert np.abs(got[:6] - row[:6]).max() < 1.1e-4 and np.abs(got[6:8] - row[6:8]).max() < 0.0101, (key, got, row)
        assert ("*" in f[k + 10:]) == (abs(row[6]) > 1.96)
    rows = _station_table(adj)
    for name, row in zip(z["stn_names"].tolist(), z["stn_rows"]):
        got = np.array(rows[name])
        assert np.abs(got[[0, 1]] - row[[0, 1]]).max() < 2e-9          # latitude, longitude as ddd.mmsssssss
        assert np.abs(got[3:10] - row[3:10]).max() < 1.1e-4, name       # h, X Y Z, SD e n up


def test_cli_reproduces_reference_expected_adj_hostsim(cli_hostsim, tmp_path):
    _golden_gnss_text(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_reproduces_reference_expected_adj_gpu(cli_gpu, tmp_path):
    _golden_gnss_text(cli_gpu, tmp_path)


REFERENCE_EXPECTED = "/root/reference/sampleData/gnss.simult.adj.expected"


@pytest.mark.skipif(not os.path.exists(REFERENCE_EXPECTED), reason="needs the reference checkout (its expected file is not copied into this repo)")
def test_cli_passes_the_reference_ci_check_on_the_gnss_network(cli_hostsim, tmp_path):
    """The reference's own CI test `test-gnss-network` (CMakeLists.txt:1033-1034): `dnadiff gnss.simult.adj
    gnss.simult.adj.expected --skip-headers 52 -t 0.001` — with the comparison rule of dnadiff restated in tools/dnadiff.py
    (dnadiff.cpp:40-268: line by line, token by token, numbers within the tolerance, text exactly).  Our file has the same
    line layout as the reference's, so everything after the first 52 lines is compared: the solution block, all 417
    adjusted-measurement rows and the 43 adjusted-station rows."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("dnadiff", os.path.join(ROOT, "tools", "dnadiff.py"))
    dnadiff = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dnadiff)
    z = np.load(os.path.join(ROOT, "tests", "golden", "gnss_sample.npz"))
    stn, msr = np.ascontiguousarray(z["stn"].astype(STN_DTYPE)), np.ascontiguousarray(z["msr"].astype(MSR_DTYPE))
    _write_network(tmp_path, "gnss", stn, msr)
    r = _run(cli_hostsim, tmp_path, "gnss", "--output-adj-msr", "--scale-normals-to-unity")     # the CI's adjust-gnss-network command line
    assert r.returncode == 0, r.stderr
    import io
    log = io.StringIO()
    differences = dnadiff.compare(os.path.join(tmp_path, "gnss.simult.adj"), REFERENCE_EXPECTED, tolerance=0.001, skip_headers=52, verbose=True, out=log)
    assert differences == 0, log.getvalue()[:4000]
    # the rule does bite: a changed digit is a difference
    text = open(os.path.join(tmp_path, "gnss.simult.adj")).read().replace("336.64", "336.74", 1)
    open(os.path.join(tmp_path, "tampered.adj"), "w").write(text)
    assert dnadiff.compare(os.path.join(tmp_path, "tampered.adj"), REFERENCE_EXPECTED, tolerance=0.001, skip_headers=52) == 1


def _golden_urban_text(exe, tmp_path):
    """`dnaadjust urban --output-adj-msr` on the reference's urban sample (tests/golden/urban_sample.npz: types
    A B G H K L M S V Y Z, the Y cluster given as latitude / longitude / orthometric height) against the rows of
    sampleData/urban.phased.adj.expected, parsed from our file by the parser that made the fixture.  Printed standard
    deviations agree to the last digit (2e-4" on the Y rows, whose Jacobian is taken at the adjusted point); values tied to
    station heights carry the 0.5 mm rounding of the exported geoid file (see tests/test_golden.py)."""
    from tests.golden.make_urban_sample import parse_expected
    z = np.load(os.path.join(ROOT, "tests", "golden", "urban_sample.npz"))
    stn, msr = np.ascontiguousarray(z["stn"].astype(STN_DTYPE)), np.ascontiguousarray(z["msr"].astype(MSR_DTYPE))
    _write_network(tmp_path, "urban", stn, msr)
    r = _run(exe, tmp_path, "urban", "--output-adj-msr", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    sol, keys, rows, stn_names, stn_rows = parse_expected(os.path.join(tmp_path, "urban.simult.adj"))
    want = dict(zip(z["solution_keys"].tolist(), z["solution"].tolist()))
    assert all(sol[k] == want[k] for k in ("unknowns", "measurements", "dof", "outliers"))
    assert abs(sol["chi_squared"] - want["chi_squared"]) < 0.2 and abs(sol["sigma_zero"] - want["sigma_zero"]) < 0.0011
    assert keys == z["msr_keys"].tolist() and len(keys) == 1182
    sec = np.radians(1.0 / 3600.0)
    for key, got, w in zip(keys, rows, z["msr_rows"]):
        t, comp = key[0], key.split()[-1]
        ang = t in "ABKVZ" or (t == "Y" and comp in "PL")
        unit = sec if ang else 1.0
        assert abs(got[0] - w[0]) / unit < 1.1e-4, key                                   # the measurement as given
        assert np.abs(got[3:6] - w[3:6]).max() < (2.6e-4 if ang else 1.1e-4), (key, got, w)     # the three SD columns
        tol = (0.07 if t in "VZ" else 0.01) if ang else 3e-4
        assert abs(got[1] - w[1]) / unit < tol and abs(got[2] - w[2]) < tol, (key, got, w)
        assert abs(got[6] - w[6]) < 0.021 and abs(got[7] - w[7]) < 0.011, (key, got, w)
        assert abs(got[8] - w[8]) < (0.05 if ang else 3e-4), (key, got, w)
    # dnaimport sorts the station file by name; the fixture keeps the order of the ASCII file — pair the rows by name
    assert sorted(stn_names) == sorted(z["stn_names"].tolist())
    stn_rows = stn_rows[[stn_names.index(n) for n in z["stn_names"].tolist()]]
    assert np.abs(stn_rows[:, [0, 1]] - z["stn_rows"][:, [0, 1]]).max() < 2e-9          # latitude, longitude as ddd.mmsssssss
    assert np.abs(stn_rows[:, 3:7] - z["stn_rows"][:, 3:7]).max() < 2.1e-4                # h, X Y Z
    assert np.abs(stn_rows[:, 7:10] - z["stn_rows"][:, 7:10]).max() < 1.1e-4            # SD e n up


def test_cli_reproduces_reference_expected_urban_adj_hostsim(cli_hostsim, tmp_path):
    _golden_urban_text(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_reproduces_reference_expected_urban_adj_gpu(cli_gpu, tmp_path):
    _golden_urban_text(cli_gpu, tmp_path)


def _golden_urban_mt_text(exe, tmp_path):
    """The reference's CI chain on urban_mt, with its own two command lines (CMakeLists.txt:1081, 1083): the first adjustment
    with the full set of report options updates the binary files, the second `dnaadjust urban_mt --output-adj-msr --multi`
    starts from them; its .adj is compared with sampleData/urban_mt.phased-mt.adj.expected (tests/golden/urban_mt_sample.npz:
    data moved GDA94 -> GDA2020 by the reference-frame step, 1182 rows incl. the point cluster in P / L / H form)."""
    from tests.golden.make_urban_sample import parse_expected
    z = np.load(os.path.join(ROOT, "tests", "golden", "urban_mt_sample.npz"))
    stn, msr = np.ascontiguousarray(z["stn"].astype(STN_DTYPE)), np.ascontiguousarray(z["msr"].astype(MSR_DTYPE))
    _write_network(tmp_path, "urban_mt", stn, msr)
    isl = parity.chain_blocks(len(stn), 60)
    dnafiles.write_seg(os.path.join(tmp_path, "urban_mt.seg"), isl, [[] for _ in isl], [[] for _ in isl])
    first = ("--output-adj-msr --multi --free-stn-sd 4.0 --fixed-stn-sd 0.000001 --max-iterations 20 --output-tstat-adj-msr --sort-adj-msr-field 2 "
             "--sort-stn-orig-order --stn-coord-types PLHhENz --angular-stn-type 1 --angular-msr-type 1 --precision-stn-linear 3 --precision-msr-linear 3 "
             "--precision-stn-angular 4 --precision-msr-angular 4 --output-pos-uncertainty --output-all-covariances --output-corrections-file").split()
    r = _run(exe, tmp_path, "urban_mt", *first)
    assert r.returncode == 0, r.stderr
    for e in ("adj", "xyz", "apu", "cor"):
        assert os.path.getsize(os.path.join(tmp_path, "urban_mt.phased-mt." + e)) > 1000
    r = _run(exe, tmp_path, "urban_mt", "--output-adj-msr", "--multi")
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(tmp_path, "urban_mt.phased-mt.adj")).read()
    assert len(re.findall(r"^ITERATION", text, re.M)) == 1          # already at the solution, as in the expected file
    sol, keys, rows, stn_names, stn_rows = parse_expected(os.path.join(tmp_path, "urban_mt.phased-mt.adj"))
    want = dict(zip(z["solution_keys"].tolist(), z["solution"].tolist()))
    assert all(sol[k] == want[k] for k in ("unknowns", "measurements", "dof", "outliers"))
    assert abs(sol["chi_squared"] - want["chi_squared"]) < 0.2 and abs(sol["sigma_zero"] - want["sigma_zero"]) < 0.0011
    assert keys == z["msr_keys"].tolist() and len(keys) == 1182
    sec = np.radians(1.0 / 3600.0)
    for key, got, w in zip(keys, rows, z["msr_rows"]):
        t, comp = key[0], key.split()[-1]
        ang = t in "ABKVZ" or (t == "Y" and comp in "PL")
        unit = sec if ang else 1.0
        assert abs(got[0] - w[0]) / unit < 1.1e-4, key                                   # the measurement as the file gave it
        assert np.abs(got[3:6] - w[3:6]).max() < (2.6e-4 if ang else 1.1e-4), (key, got, w)
        tol = (0.07 if t in "VZ" else 0.01) if ang else 3e-4
        assert abs(got[1] - w[1]) / unit < tol and abs(got[2] - w[2]) < tol, (key, got, w)
        assert abs(got[6] - w[6]) < 0.021 and abs(got[7] - w[7]) < 0.011, (key, got, w)
        assert abs(got[8] - w[8]) < (0.05 if ang else 3e-4), (key, got, w)
    stn_rows = stn_rows[[stn_names.index(n) for n in z["stn_names"].tolist()]]
    assert np.abs(stn_rows[:, [0, 1]] - z["stn_rows"][:, [0, 1]]).max() < 2e-9 and np.abs(stn_rows[:, 3:7] - z["stn_rows"][:, 3:7]).max() < 2.1e-4
    assert np.abs(stn_rows[:, 7:10] - z["stn_rows"][:, 7:10]).max() < 1.1e-4


def test_cli_reference_ci_chain_urban_mt_hostsim(cli_hostsim, tmp_path):
    _golden_urban_mt_text(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_reference_ci_chain_urban_mt_gpu(cli_gpu, tmp_path):
    _golden_urban_mt_text(cli_gpu, tmp_path)


def _read_snx(path):
    text = open(path).read()
    assert text.startswith("%=SNX 2.00 DNA") and text.rstrip().endswith("%ENDSNX")
    sec = lambda name: text.split("+" + name)[1].split("-" + name)[0].splitlines()[1:]
    sites = [l.split()[0] for l in sec("SITE/ID") if l.startswith(" ")]
    est = [(l.split()[1], l.split()[2], float(l.split()[8]), float(l.split()[9])) for l in sec("SOLUTION/ESTIMATE") if l.startswith(" ")]
    n = len(est)
    Q = np.zeros((n, n))
    for l in sec("SOLUTION/MATRIX_ESTIMATE L COVA"):
        if not l.startswith(" "):
            continue
        f = l.split()
        r, c0 = int(f[0]) - 1, int(f[1]) - 1
        for k, v in enumerate(f[2:]):
            Q[r, c0 + k] = Q[c0 + k, r] = float(v)
    stats = {l[1:31].strip(): float(l[31:]) for l in sec("SOLUTION/STATISTICS") if l.startswith(" ")}
    return sites, est, Q, stats


def _sinex(exe, oracle, tmp_path):
    """--export-sinex-file (PRN:2906-3010, snx_file_writer.cpp): estimates, standard deviations and the
    lower triangle of the dense variance matrix, for the whole network (simultaneous) and per block (phased)."""
    stn, msr, _, _ = synth.gnss_network(90, 260, 14)
    stn["stationName"] = np.char.add("P", np.char.zfill(np.arange(90).astype(str), 3)).astype("S31")   # 4-character SINEX codes
    _write_network(tmp_path, "sx", stn, msr)
    ref = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_vcv=True)
    V, est_ref = ref["vcv"], ref["est"].reshape(-1, 3)
    names = [n.decode() for n in stn["stationName"]]

    def check(path):
        sites, est, Q, stats = _read_snx(path)
        idx = np.concatenate([[3 * names.index(c), 3 * names.index(c) + 1, 3 * names.index(c) + 2] for c in sites])
        assert [e[1] for e in est] == [c for c in sites for _ in range(3)] and [e[0] for e in est] == ["STAX", "STAY", "STAZ"] * len(sites)
        assert np.abs(np.array([e[2] for e in est]) - est_ref.reshape(-1)[idx]).max() < 1e-8
        want = V[np.ix_(idx, idx)]
        assert np.abs(Q - want).max() <= 2e-8 * np.abs(want).max()
        assert np.abs(np.array([e[3] for e in est]) - np.sqrt(np.diag(want))).max() <= 1e-5 * np.sqrt(np.diag(want)).max()
        assert stats["NUMBER OF UNKNOWNS"] == ref["res"].unknown_params and abs(stats["VARIANCE FACTOR"] - ref["res"].sigma_zero) < 1e-6
        return sites

    r = _run(exe, tmp_path, "sx", "--export-sinex-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    assert len(check(os.path.join(tmp_path, "sx.GDA2020.snx"))) == 90
    isl = parity.chain_blocks(90, 30)
    dnafiles.write_seg(os.path.join(tmp_path, "sx.seg"), isl, [[] for _ in isl], [[] for _ in isl])
    r = _run(exe, tmp_path, "sx", "--phased", "--export-sinex-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    seen = set()
    for b in range(len(isl)):
        seen.update(check(os.path.join(tmp_path, f"sx-block{b + 1}.GDA2020.snx")))
    assert seen == set(names)


def test_cli_sinex_hostsim(cli_hostsim, oracle, tmp_path):
    _sinex(cli_hostsim, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_sinex_gpu(cli_gpu, oracle, tmp_path):
    _sinex(cli_gpu, oracle, tmp_path)


def _type_b_uncertainties(exe, oracle, tmp_path):
    """--type-b-sd-global / --type-b-sd-file (ADJ:10231-10323, PRN:4000-4029): type B variances (e, n, up) added to the
    printed station uncertainties; site-specific values from the file override the global ones."""
    stn, msr, _, _ = synth.gnss_network(40, 110, 52)
    _write_network(tmp_path, "tb", stn, msr)
    tbu = os.path.join(tmp_path, "sites.tbu")
    s7 = stn["stationName"][7].decode()
    with open(tbu, "w") as f:
        f.write("!#=DNA 1.00 TBU\n* site-specific type B uncertainties\n%-20s%-13s%-13s%-13s\nNOTINNET            1.0 1.0 1.0\n" % (s7, "0.030", "0.040", "0.050"))
    r = _run(exe, tmp_path, "tb", "--type-b-sd-global", "0.010,0.010,0.020", "--type-b-sd-file", tbu, "--output-pos-uncertainty",
             "--no-binary-update")
    assert r.returncode == 0, r.stderr
    stn_o = stn.copy()
    ref = oracle.adjust_simultaneous(stn_o, msr.copy(), want_vcv=True)
    V = ref["vcv"]
    rows = _station_table(open(os.path.join(tmp_path, "tb.simult.xyz")).read())
    for i in range(len(stn)):
        lat, lon = stn_o["currentLatitude"][i], stn_o["currentLongitude"][i]
        sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
        R = np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])
        tb = np.array([0.030, 0.040, 0.050]) if i == 7 else np.array([0.010, 0.010, 0.020])
        q = R.T @ V[3 * i:3 * i + 3, 3 * i:3 * i + 3] @ R
        sd = np.sqrt(np.diag(q) + tb ** 2 + np.array([0, 0, float(stn["geoidSepUnc"][i]) ** 2]))
        assert np.abs(np.array(rows[stn["stationName"][i].decode()][7:10]) - sd).max() < 1e-4
    assert "Type B uncertainties:" in open(os.path.join(tmp_path, "tb.simult.apu")).read()
    r = _run(exe, tmp_path, "tb", "--type-b-sd-global", "0.01,abc")
    assert r.returncode == 1 and "is not a number" in r.stderr


def test_cli_type_b_uncertainties_hostsim(cli_hostsim, oracle, tmp_path):
    _type_b_uncertainties(cli_hostsim, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_type_b_uncertainties_gpu(cli_gpu, oracle, tmp_path):
    _type_b_uncertainties(cli_gpu, oracle, tmp_path)


def test_cli_apu_cor_hostsim(cli_hostsim, oracle, tmp_path):
    _apu_cor(cli_hostsim, oracle, tmp_path)


def test_cli_simultaneous_hostsim(cli_hostsim, oracle, tmp_path):
    _simultaneous(cli_hostsim, oracle, tmp_path)


def test_cli_phased_seg_hostsim(cli_hostsim, oracle, tmp_path):
    _phased(cli_hostsim, oracle, tmp_path)


def test_cli_errors(cli_hostsim, tmp_path):
    r = _run(cli_hostsim, tmp_path, "nonet")
    assert r.returncode == 1 and "Error" in r.stderr                     # missing files: EXIT_FAILURE (WRAP:1377)
    r = _run(cli_hostsim, tmp_path, "c1", "--no-such-flag")
    assert r.returncode == 1 and "unrecognised option" in r.stderr       # WRAP:1040-1046
    r = _run(cli_hostsim, tmp_path, "c1", "--output")                   # ambiguous prefix
    assert r.returncode == 1 and "ambiguous" in r.stderr
    stn, msr, _, _ = synth.config_network("C1")
    _write_network(tmp_path, "old", stn, msr)
    with open(os.path.join(tmp_path, "old.bms"), "r+b") as f:           # file version gate (io/bms_file.cpp:151)
        f.seek(10)
        f.write(b"       1.0")
    r = _run(cli_hostsim, tmp_path, "old")
    assert r.returncode == 1 and "predates" in r.stderr


@pytest.mark.gpu
def test_cli_simultaneous_gpu(cli_gpu, oracle, tmp_path):
    _simultaneous(cli_gpu, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_phased_seg_gpu(cli_gpu, oracle, tmp_path):
    _phased(cli_gpu, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_apu_cor_gpu(cli_gpu, oracle, tmp_path):
    _apu_cor(cli_gpu, oracle, tmp_path)


def _station_map_and_asl(exe, oracle, tmp_path):
    """<net>.map and <net>.asl as dnaimport writes them (map_file.cpp:44-98, asl_file.cpp:80-93) are read when present:
    --constraints resolves names through the station map (LDR:243-244), and the validity flags of the .asl — one station
    here is in the station file without any measurement — decide with the measurement list which stations are reported."""
    stn, msr, _, _ = synth.gnss_network(60, 170, 35)
    # a 61st station nobody measures: in the files, not in the adjustment
    stn2 = np.zeros(61, dtype=STN_DTYPE)
    stn2[:60] = stn
    stn2[60] = stn[5]
    stn2["stationName"][60] = stn2["stationNameOrig"][60] = b"LONELY"
    _write_network(tmp_path, "am", stn2, msr)
    names = [s.decode() for s in stn2["stationName"]]
    dnafiles.write_map(os.path.join(tmp_path, "am.map"), names)
    validity = np.ones(61, dtype=np.uint16)
    validity[60] = 0
    dnafiles.write_asl(os.path.join(tmp_path, "am.asl"), np.ones(61), np.arange(61), validity)
    n1 = names[7]
    r = _run(exe, tmp_path, "am", "--constraints", f"{n1},CCC", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    stn_c = stn.copy()
    stn_c["stationConst"][7] = b"CCC"
    _check_outputs(oracle, tmp_path, "am", "simult", stn_c, msr, False)
    adj = open(os.path.join(tmp_path, "am.simult.adj")).read()
    assert "LONELY" not in adj.split("Adjusted Coordinates")[1]
    # a station map that does not cover the station file is not used (the names in the .bst records are)
    dnafiles.write_map(os.path.join(tmp_path, "am.map"), names[:10])
    r = _run(exe, tmp_path, "am", "--constraints", f"{names[40]},CCC", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    # truncated files fail loudly with the reference's wording
    with open(os.path.join(tmp_path, "am.asl"), "r+b") as f:
        f.truncate(100)
    r = _run(exe, tmp_path, "am")
    assert r.returncode == 1 and "An error was encountered when reading from" in r.stderr


def test_cli_station_map_and_associated_station_list_hostsim(cli_hostsim, oracle, tmp_path):
    _station_map_and_asl(cli_hostsim, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_station_map_and_associated_station_list_gpu(cli_gpu, oracle, tmp_path):
    _station_map_and_asl(cli_gpu, oracle, tmp_path)


def test_cli_gpus_option_threads(cli_hostsim, oracle, tmp_path):
    """dnaadjust --gpus N: the ranks as threads of the command line (csrc/host/gpu_group.hpp), every collective call made by
    all of them; outputs as from one device — simultaneous and phased, with the per-block SINEX files whose dense block matrices come
    from the rank holding the block."""
    stn, msr, _, _ = synth.gnss_network(900, 2700, 41)
    _write_network(tmp_path, "mg", stn, msr)
    r = _run(cli_hostsim, tmp_path, "mg", "--gpus", "3", "--output-adj-msr", "--output-pos-uncertainty", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    _check_outputs(oracle, tmp_path, "mg", "simult", stn, msr, True)
    one = _run(cli_hostsim, tmp_path, "mg", "--output-adj-msr", "--output-pos-uncertainty", "--no-binary-update", "--output-folder",
               str(tmp_path / "one"))
    # phased over a chain of blocks, block-wise station output and per-block SINEX (dense block matrices from their holders)
    blocks = parity.chain_blocks(len(stn), 150)
    isl = [list(b) for b in blocks]
    jsl, cml = [[] for _ in isl], [[] for _ in isl]
    dnafiles.write_seg(os.path.join(tmp_path, "mg.seg"), isl, jsl, cml)
    r = _run(cli_hostsim, tmp_path, "mg", "--phased", "--gpus", "2", "--export-sinex-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    _check_outputs(oracle, tmp_path, "mg", "phased", stn, msr, False)
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".snx")]) == len(blocks)
    r = _run(cli_hostsim, tmp_path, "mg", "--gpus", "9")
    assert r.returncode != 0 and "--gpus takes 1 to 8" in r.stderr
