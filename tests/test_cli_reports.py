"""Report-layout options of the dnaadjust command line (the flag set the reference's CI drives it with,
CMakeLists.txt:1070-1163): sorting and units of the adjusted-measurement table, direction-set layout, t statistics,
ignored measurements a posteriori, per-iteration reports, measurements-to-station table, coordinate types, precision
and corrections of the station table.  Hostsim stand-in on CPU; one run on the product binary on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from dynadjust_b200 import dnafiles, synth
from dynadjust_b200.records import MSR_DTYPE
from dynadjust_b200 import synth_terrestrial as st
from tests import parity
from tests.golden import dna_ascii
from tests.test_cli import _run, _write_network, cli_gpu, cli_hostsim  # noqa: F401  (fixtures)

SEC = np.radians(1.0 / 3600.0)
TYPES = "ABCDEGHIJKLMPQRSVXYZ"


def _network():
    stn, msr, _, _ = st.terrestrial_network(60, 170, 77, scalars={"S": 30, "A": 10, "L": 12, "V": 8, "H": 5, "E": 4, "M": 4, "B": 4, "K": 4, "Z": 4},
                                            n_dir_sets=6, n_x=2, n_y=2, ignore_some=True)
    g0 = np.where(msr["measType"] == b"G")[0][:3]           # ignored: one baseline, one slope distance, one angle
    s0 = np.where(msr["measType"] == b"S")[0][0]
    a0 = np.where(msr["measType"] == b"A")[0][0]
    msr["ignore"][g0] = 1
    msr["ignore"][[s0, a0]] = 1
    return stn, msr, (g0[0], s0, a0)


def _tables(text, heading):
    """Bodies of every table whose heading line starts with `heading`."""
    out = []
    for part in text.split("\n" + heading)[1:]:
        lines = part.split("\n")
        dash = next(i for i, l in enumerate(lines) if l.startswith("-----") and i > 1)
        body = []
        for l in lines[dash + 1:]:
            if not l.strip():
                break
            body.append(l)
        out.append((lines[dash - 1], body))
    return out


def _num(s):
    return float(s)


def _dms(f):
    sign = -1.0 if f[0].startswith("-") else 1.0
    return sign * np.radians(abs(float(f[0])) + float(f[1]) / 60.0 + float(f[2]) / 3600.0)


def _local(lat, lon):
    sl, cl, so, co = np.sin(lat), np.cos(lat), np.sin(lon), np.cos(lon)
    return np.array([[-so, -sl * co, cl * co], [co, -sl * so, cl * so], [0, cl, sl]])   # columns e, n, up


def _reports(exe, tmp_path):
    stn, msr, (g0, s0, a0) = _network()
    _write_network(tmp_path, "net", stn, msr)
    plain = _run(exe, tmp_path, "net", "--output-adj-msr", "--no-binary-update")
    assert plain.returncode == 0, plain.stderr
    base = open(os.path.join(tmp_path, "net.simult.adj")).read()
    os.rename(os.path.join(tmp_path, "net.simult.adj"), os.path.join(tmp_path, "plain.adj"))
    r = _run(exe, tmp_path, "net", "--output-adj-msr", "--no-binary-update", "--output-ignored-msrs", "--output-msr-to-stn",
             "--sort-msr-to-stn-field", "3", "--output-iter-adj-stat", "--output-iter-adj-msr", "--output-iter-adj-stn", "--output-iter-cmp-msr",
             "--output-tstat-adj-msr", "--stn-corrections", "--sort-adj-msr-field", "7", "--stn-coord-types", "PLHhENzXYZ",
             "--angular-stn-type", "1", "--precision-stn-linear", "3", "--precision-stn-angular", "4", "--precision-msr-linear", "5",
             "--precision-msr-angular", "3", "--comments", "report test", "--verbose-level", "1")
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(tmp_path, "net.simult.adj")).read()
    grab = lambda t, label: re.findall(r"^" + re.escape(label) + r"\s+(\S+)", t, re.M)
    # the per-iteration reports do not disturb the solution
    assert grab(text, "Chi squared")[-1] == grab(base, "Chi squared")[-1]
    assert grab(text, "Rigorous Sigma Zero")[-1] == grab(base, "Rigorous Sigma Zero")[-1]
    iterations = len(re.findall(r"^ITERATION\s+\d+", text, re.M))
    assert iterations >= 2
    assert len(_tables(text, "Computed Measurements (a-priori)")) == iterations
    adj_tables = _tables(text, "Adjusted Measurements")
    assert len(adj_tables) == iterations                      # one per non-final iteration + the final table
    assert len(_tables(text, "Adjusted Coordinates")) == iterations
    assert len(grab(text, "Chi squared")) == iterations
    sigma0 = float(grab(text, "Rigorous Sigma Zero")[-1])

    # ---- final adjusted measurements: columns, precision, t statistic, n-stat ordering, direction-set layout
    head, body = adj_tables[-1]
    assert head.split()[-8:] == ["N-stat", "T-stat", "Pelzer", "Rel", "Pre", "Adj", "Corr", "Outlier?"]
    groups, cur, key = [], None, None
    for l in body:
        # rows of one measurement: the three components of a baseline, every member of an X / Y cluster, the angles of a set
        k = l[0] if l[0] in "XY" else l[:62]
        if l[0] != " " and k != key:
            cur = dict(type=l[0], rows=[], head=l)
            groups.append(cur)
            key = k
        if l[0] == "D" and len(l.split()) == 4:
            continue                                          # heading row of a direction set: instrument, RO, count
        f = l[67:].split()
        ang = not re.fullmatch(r"-?\d+\.\d+", f[0])           # d m s fields
        vals = f[6:] if ang else f[2:]
        cur["rows"].append((l, [float(x) for x in vals[:8]]))
    for g in groups:
        for l, v in g["rows"]:
            corr, nstat, tstat = v[0], v[4], v[5]
            assert abs(tstat - nstat / np.sqrt(sigma0)) < 0.011, l
    big = [max(abs(v[4]) for _, v in g["rows"]) for g in groups if g["rows"]]
    assert all(a >= b - 0.0051 for a, b in zip(big, big[1:]))    # --sort-adj-msr-field 7: largest |n-stat| first
    dsets = [g for g in groups if g["type"] == "D"]
    assert len(dsets) == 6
    for g in dsets:
        f = g["head"].split()
        assert len(f) == 4 and int(f[3]) == len(g["rows"])        # D inst RO n, then n derived angles against their targets
        assert all(l[:42].strip() == "" and l[42:62].strip() for l, _ in g["rows"])
    lin = next(l for g in groups if g["type"] == "S" for l, _ in g["rows"])
    assert re.search(r"\d+\.\d{5}\s+\d+\.\d{5}\s+-?\d+\.\d{5}", lin[67:])     # --precision-msr-linear 5
    angrow = next(l for g in groups if g["type"] == "A" for l, _ in g["rows"])
    assert re.search(r"\d+ \d\d \d\d\.\d{3}\s+\d+ \d\d \d\d\.\d{3}\s", angrow[67:])   # --precision-msr-angular 3

    # ---- ignored measurements, a posteriori: computed from the adjusted coordinates
    stn_o, msr_o = stn.copy(), msr.copy()
    (_, ign), = _tables(text, "Ignored Measurements (a-posteriori)")
    names = [n.decode() for n in stn["stationName"]]
    coords = {}
    for l in _tables(text, "Adjusted Coordinates")[-1][1]:
        f = l.split()
        coords[f[0]] = np.array([float(x) for x in f[9:12]])       # X Y Z (3 decimals)
    assert [l[0] for l in ign] == ["G", "G", "G", "S", "A"] and all(l[62] == "*" for l in ign)
    s1, s2 = names[msr["station1"][g0]], names[msr["station2"][g0]]
    for q, l in enumerate(ign[:3]):
        f = l[67:].split()
        assert abs(float(f[0]) - msr["term1"][g0 + q]) < 1e-5 and abs(float(f[1]) - (coords[s2] - coords[s1])[q]) < 2.1e-3
        assert abs(float(f[2]) - (float(f[1]) - float(f[0]))) < 2.1e-5
    f = ign[3][67:].split()
    d = np.linalg.norm(coords[names[msr["station2"][s0]]] - coords[names[msr["station1"][s0]]])
    assert abs(float(f[1]) - d) < 0.05 and abs(float(f[0]) - msr["term1"][s0]) < 1e-5    # instrument / target heights are small

    # ---- measurements to station: totals per type = stations touched by the measurements that take part
    (mhead, m2s), = _tables(text, "Measurements to Station")
    assert mhead.split() == ["Station"] + list(TYPES) + ["Total"]
    totals = text.split("\nTotals")[1].split("\n")[0]
    tot = {t: (int(totals[19 + 8 * k:27 + 8 * k]) if totals[19 + 8 * k:27 + 8 * k].strip() else 0) for k, t in enumerate(TYPES)}
    used = msr[msr["ignore"] == 0]
    assert tot["S"] == 2 * int((used["measType"] == b"S").sum()) and tot["A"] == 3 * int((used["measType"] == b"A").sum())
    assert tot["H"] == int((used["measType"] == b"H").sum())
    assert tot["G"] == 2 * int(((used["measType"] == b"G") & (used["measStart"] == 0)).sum())
    m2s = m2s[:next(i for i, l in enumerate(m2s) if l.startswith("-----"))]      # up to the line above "Totals"
    counts = [int(l.split()[-1]) for l in m2s]
    assert counts == sorted(counts, reverse=True) and len(m2s) == len(stn)          # --sort-msr-to-stn-field 3

    # ---- station table: coordinate types, decimal degrees, grid coordinates, corrections
    shead, rows = _tables(text, "Adjusted Coordinates")[-1]
    assert shead.split()[:12] == ["Station", "Const", "Latitude", "Longitude", "H(Ortho)", "h(Ellipse)", "Easting", "Northing", "Zone", "X", "Y", "Z"]
    assert shead.split()[12:18] == ["SD(e)", "SD(n)", "SD(up)", "Corr(e)", "Corr(n)", "Corr(up)"]
    x0 = synth.geo_to_cart(stn["initialLatitude"], stn["initialLongitude"], stn["initialHeight"])
    for l in rows:
        f = l.split()
        i = names.index(f[0])
        lat, lon = np.radians(float(f[2])), np.radians(float(f[3]))
        assert re.fullmatch(r"-?\d+\.\d{8}", f[2]) and re.fullmatch(r"-?\d+\.\d{3}", f[9])
        glat, glon = dna_ascii.grid_to_geo(float(f[6]), float(f[7]), int(f[8]))
        assert abs(glat - lat) < 5e-10 and abs(glon - lon) < 5e-10, l           # 3 mm on the ground
        R = _local(lat, lon)
        corr = R.T @ (np.array([float(x) for x in f[9:12]]) - x0[i])
        assert np.abs(corr - np.array([float(x) for x in f[15:18]])).max() < 2.1e-3, l
    return stn_o, msr_o


def test_cli_report_options_hostsim(cli_hostsim, tmp_path):
    _reports(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_report_options_gpu(cli_gpu, tmp_path):
    _reports(cli_gpu, tmp_path)


def test_cli_gnss_alternate_units(cli_hostsim, tmp_path):
    """--output-adj-gnss-units 1 / 2 / 3: baselines in east-north-up, azimuth-elevation-distance, azimuth-distance-up of the
    local frame at the first station (PRN:4717-5047), against the Cartesian rows of the same adjustment."""
    stn, msr, _, _ = synth.gnss_network(40, 110, 5)
    _write_network(tmp_path, "g", stn, msr)
    names = [n.decode() for n in stn["stationName"]]

    def rows(*flags):
        r = _run(cli_hostsim, tmp_path, "g", "--output-adj-msr", "--no-binary-update", *flags)
        assert r.returncode == 0, r.stderr
        text = open(os.path.join(tmp_path, "g.simult.adj")).read()
        body = _tables(text, "Adjusted Measurements")[-1][1]
        st_tab = {l.split()[0]: l.split() for l in _tables(text, "Adjusted Coordinates")[-1][1]}
        return body, st_tab

    xyz, st_tab = rows()
    enu, _ = rows("--output-adj-gnss-units", "1")
    aed, _ = rows("--output-adj-gnss-units", "2")
    adu, _ = rows("--output-adj-gnss-units", "3")
    assert len(xyz) == len(enu) == len(aed) == len(adu) == 330

    def hp(v):   # ddd.mmsssss -> radians
        v = float(v)
        a = abs(v)
        d = np.floor(a + 1e-12)
        m = np.floor((a - d) * 100 + 1e-9)
        s = ((a - d) * 100 - m) * 100
        return np.sign(v) * np.radians(d + m / 60 + s / 3600)

    for b in range(0, 330, 3):
        s1 = xyz[b][2:22].strip()
        lat, lon = hp(st_tab[s1][2]), hp(st_tab[s1][3])
        R = _local(lat, lon)
        meas = np.array([float(xyz[b + q][67:].split()[0]) for q in range(3)])
        adjd = np.array([float(xyz[b + q][67:].split()[1]) for q in range(3)])
        lm, la = R.T @ meas, R.T @ adjd
        assert [enu[b + q][65] for q in range(3)] == ["e", "n", "u"]
        for q in range(3):
            f = enu[b + q][67:].split()
            assert abs(float(f[0]) - lm[q]) < 2e-4 and abs(float(f[1]) - la[q]) < 2e-4 and abs(float(f[2]) - (la[q] - lm[q])) < 2.1e-4
            # statistics recomputed in that frame: n-stat = correction / corr. sd, corr. sd^2 = meas. sd^2 - adj. sd^2
            assert abs(float(f[5]) ** 2 - abs(float(f[3]) ** 2 - float(f[4]) ** 2)) < 3e-4 * max(float(f[5]), 1e-3) + 2e-8
        assert [aed[b + q][65] for q in range(3)] == ["a", "v", "s"] and [adu[b + q][65] for q in range(3)] == ["a", "s", "u"]
        az = np.arctan2(lm[0], lm[1]) % (2 * np.pi)
        el = np.arctan2(lm[2], np.hypot(lm[0], lm[1]))
        dist = np.linalg.norm(lm)
        fa, fv, fs = (aed[b + q][67:].split() for q in range(3))
        assert abs(_dms(fa[0:3]) - az) < 2e-4 / dist + 1e-9 and abs(_dms(fv[0:3]) - el) < 2e-4 / dist + 1e-9 and abs(float(fs[0]) - dist) < 2e-4
        ga, gs, gu = (adu[b + q][67:].split() for q in range(3))
        assert ga[:6] == fa[:6] and abs(float(gs[0]) - dist) < 2e-4 and abs(float(gu[0]) - lm[2]) < 2e-4
        # the standard deviation of the azimuth (seconds) times the distance is a length comparable with the e / n values
        sd_az = float(fa[7]) * SEC * dist
        sd_en = [float(enu[b + q][67:].split()[3]) for q in range(2)]
        assert 0.5 * min(sd_en) < sd_az < 2.0 * max(sd_en)


def _cluster_vcv_of(m):
    """Full variance matrix of a Y cluster held in binary records (as the engine loads it)."""
    first = [i for i in range(len(m)) if m["measStart"][i] == 0]
    n = len(first)
    V = np.zeros((3 * n, 3 * n))
    for k, i in enumerate(first):
        r = m[i:i + 3]
        blk = np.array([[r["term2"][0], r["term2"][1], r["term2"][2]], [r["term2"][1], r["term3"][1], r["term3"][2]],
                        [r["term2"][2], r["term3"][2], r["term4"][2]]])
        V[3 * k:3 * k + 3, 3 * k:3 * k + 3] = blk
        for q in range(int(r["vectorCount2"][0])):
            cv = m[i + 3 + 3 * q:i + 6 + 3 * q]
            B = np.stack([cv["term1"], cv["term2"], cv["term3"]], axis=1)
            j = k + 1 + q
            V[3 * k:3 * k + 3, 3 * j:3 * j + 3] = B
            V[3 * j:3 * j + 3, 3 * k:3 * k + 3] = B.T
    return [int(m["station1"][i]) for i in first], np.array([m["term1"][i:i + 3] for i in first]), V


def _exports(exe, oracle, tmp_path):
    """--export-dna-stn-file / --export-dna-msr-file / --export-xml-stn-file / --export-xml-msr-file (SURVEY 8f item 4;
    PRN:2775-2903, 3012-3164): the adjusted stations in the form they were supplied in, and the estimates with their
    full variance matrix as a GNSS point cluster per block.  The DNA files are read back with the DNA reader of the
    golden fixtures; the DynaML files by tag."""
    stn, msr, _, _ = synth.gnss_network(45, 130, 21)
    stn["suppliedStationType"][::3] = 0            # XYZ
    stn["suppliedStationType"][1::3] = 3           # UTM
    stn["fileOrder"] = np.arange(len(stn))[::-1]   # exported in the order of the imported file
    _write_network(tmp_path, "ex", stn, msr)
    ref = oracle.adjust_simultaneous(stn.copy(), msr.copy(), want_vcv=True)
    V, est = ref["vcv"], ref["est"].reshape(-1, 3)
    names = [n.decode() for n in stn["stationName"]]
    r = _run(exe, tmp_path, "ex", "--export-dna-stn-file", "--export-dna-msr", "--export-xml-stn-file", "--export-xml-msr-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    stem = os.path.join(tmp_path, "ex.simult.adj")
    head = open(stem + ".stn").readline().split()
    assert head[:3] == ["!#=DNA", "3.01", "STN"] and head[-1] == "45" and head[4] == "GDA2020"
    back = dna_ascii.read_stations(stem + ".stn")
    assert [n.decode() for n in back["stationName"]] == names[::-1]
    xyz = synth.geo_to_cart(back["currentLatitude"], back["currentLongitude"], back["currentHeight"] +
                            np.where(back["suppliedStationType"] == 0, 0.0, stn["geoidSep"][::-1]))   # LLH / UTM rows carry H
    assert np.abs(xyz - est[::-1]).max() < 4e-4
    assert sorted(set(l[24:27] for l in open(stem + ".stn") if l[0] not in "!*")) == ["LLH", "UTM", "XYZ"]
    cl = dna_ascii.read_measurements(stem + ".msr", stn, reftran=False)
    who, val, Q = _cluster_vcv_of(cl)
    assert who == list(range(45)) and np.abs(val - est).max() < 1e-4
    assert np.abs(Q - V).max() <= 2e-8 * np.abs(V).max()
    x = open(stem + ".msr.xml").read()
    assert x.count("<Clusterpoint>") == 45 and x.count("<PointCovariance>") == 45 * 44 // 2 and "<Total>45</Total>" in x
    sxx = [float(v) for v in re.findall(r"<SigmaXX>(\S+)</SigmaXX>", x)]
    assert np.abs(np.array(sxx) - np.diag(V)[0::3]).max() <= 2e-8 * np.abs(V).max()
    m13 = [float(v) for v in re.findall(r"<m13>(\S+)</m13>", x)]
    assert abs(m13[0] - V[0, 5]) <= 2e-8 * np.abs(V).max()
    sx = open(stem + ".stn.xml").read()
    assert sx.count("<DnaStation>") == 45 and sx.count("<HemisphereZone>") == 15 and sx.rstrip().endswith("</DnaXmlFormat>")
    # phased: one cluster per block, inner and junction stations
    isl = parity.chain_blocks(45, 15)
    dnafiles.write_seg(os.path.join(tmp_path, "ex.seg"), isl, [[] for _ in isl], [[] for _ in isl])
    r = _run(exe, tmp_path, "ex", "--phased", "--export-dna-msr-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    cl = dna_ascii.read_measurements(os.path.join(tmp_path, "ex.phased.adj.msr"), stn, reftran=False)
    seen = set()
    for cid in np.unique(cl["clusterID"]):
        who, val, Q = _cluster_vcv_of(cl[cl["clusterID"] == cid])
        idx = np.concatenate([[3 * s, 3 * s + 1, 3 * s + 2] for s in who])
        assert np.abs(Q - V[np.ix_(idx, idx)]).max() <= 2e-8 * np.abs(V).max() and np.abs(val - est[who]).max() < 1e-4
        seen.update(who)
    assert seen == set(range(45)) and len(np.unique(cl["clusterID"])) == len(isl)


def test_cli_exports_hostsim(cli_hostsim, oracle, tmp_path):
    _exports(cli_hostsim, oracle, tmp_path)


@pytest.mark.gpu
def test_cli_exports_gpu(cli_gpu, oracle, tmp_path):
    _exports(cli_gpu, oracle, tmp_path)


def test_cli_project_file_and_short_options(cli_hostsim, tmp_path):
    """`dnaadjust -p <net>.dnaproj` (all other options ignored, WRAP:487-505; file syntax dnaprojectfile.cpp:127-310), the short
    forms -n -i -o, file-name overrides, and the output names of the execution-strategy flags (WRAP:659-734)."""
    stn, msr, _, _ = synth.gnss_network(40, 110, 8)
    _write_network(tmp_path, "pj", stn, msr)
    isl = parity.chain_blocks(40, 14)
    dnafiles.write_seg(os.path.join(tmp_path, "pj.seg"), isl, [[] for _ in isl], [[] for _ in isl])
    out = os.path.join(tmp_path, "out")
    os.mkdir(out)
    rec = lambda k, v: f"{k:<35}{v}\n"
    with open(os.path.join(tmp_path, "pj.dnaproj"), "w") as f:
        f.write("# DynAdjust project file\n\n#general" + " " * 40 + "\n" + "-" * 80 + "\n")
        f.write(rec("network-name", "pj") + rec("input-folder", str(tmp_path)) + rec("output-folder", out) + rec("quiet", "yes"))
        f.write("\n#import" + " " * 40 + "\n" + "-" * 80 + "\n" + rec("max-iterations", "1") + rec("reference-frame", "GDA2020"))
        f.write("\n#adjust" + " " * 40 + "\n" + "-" * 80 + "\n")
        f.write(rec("adjustment-mode", "phased-adjustment") + rec("multi-thread", "yes") + rec("staged-adjustment", "no") + rec("max-iterations", "7"))
        f.write(rec("output-adj-msr", "yes") + rec("output-pos-uncertainty", "no") + rec("free-stn-sd", "5.0") + rec("stn-coord-types", "ENzPLh"))
        f.write(rec("no-binary-update", "yes"))
    r = subprocess.run([cli_hostsim, "--max-iterations", "1", "-p", os.path.join(tmp_path, "pj.dnaproj"), "--output-corrections-file"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout == "", r.stderr + r.stdout
    assert sorted(os.listdir(out)) == ["pj-pam.mtx", "pj-rva.mtx", "pj.phased-mt.adj", "pj.phased-mt.xyz"]
    text = open(os.path.join(out, "pj.phased-mt.adj")).read()
    assert "Adjusted Measurements" in text and re.search(r"^SOLUTION\s+Converged", text, re.M)
    head = _tables(text, "Adjusted Coordinates")[-1][0].split()
    assert head[2:8] == ["Easting", "Northing", "Zone", "Latitude", "Longitude", "h(Ellipse)"]
    r = subprocess.run([cli_hostsim, "-p", os.path.join(tmp_path, "missing.dnaproj")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "does not exist" in r.stderr
    # short forms, file-name overrides, staged naming
    os.rename(os.path.join(tmp_path, "pj.bst"), os.path.join(tmp_path, "stations.bin"))
    r = subprocess.run([cli_hostsim, "-n", "pj", "-i", str(tmp_path), "-o", out, "--binary-stn-file", "stations.bin", "--staged", "--create-stage-files",
                        "--no-binary-update"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(os.path.join(out, "pj.phased-stage.adj"))
    r = subprocess.run([cli_hostsim, "-x"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "unrecognised option" in r.stderr


def _report_results(exe, tmp_path):
    """--report-results (WRAP:607-614; SerialiseAdjustedVarianceMatrices ADJ:6770-6799): after an adjustment has updated the
    binary files and written <net>-rva.mtx / <net>-pam.mtx, every report is printed again without solving — same numbers,
    other layout options allowed."""
    stn, msr, _, _ = st.terrestrial_network(50, 140, 61, scalars={"S": 20, "A": 8, "L": 8, "V": 6, "H": 4, "E": 3, "M": 3}, n_dir_sets=4, n_x=1, n_y=1,
                                            v_scale=1.7)
    g = np.where(msr["measType"] == b"G")[0]
    msr["scale4"][g[:30]] = 2.5                    # variance scalars on some baselines: applied once, on the first run only
    _write_network(tmp_path, "rr", stn, msr)
    flags = ["--output-adj-msr", "--output-pos-uncertainty", "--output-corrections-file", "--output-tstat-adj-msr", "--stn-corrections"]
    r = _run(exe, tmp_path, "rr", *flags)
    assert r.returncode == 0, r.stderr
    first = {e: open(os.path.join(tmp_path, "rr.simult." + e)).read() for e in ("adj", "xyz", "apu", "cor")}
    for e in first:
        os.remove(os.path.join(tmp_path, "rr.simult." + e))
    assert os.path.getsize(os.path.join(tmp_path, "rr-rva.mtx")) > 72 * 50 and os.path.exists(os.path.join(tmp_path, "rr-pam.mtx"))
    r = _run(exe, tmp_path, "rr", "--report-results", *flags)
    assert r.returncode == 0 and "Report last adjustment results" in r.stdout, r.stderr
    again = {e: open(os.path.join(tmp_path, "rr.simult." + e)).read() for e in ("adj", "xyz", "apu", "cor")}
    assert "Printing results of last adjustment only" in again["adj"] and "ITERATION" not in again["adj"]
    body = lambda t, h: [l for tab in _tables(t, h) for l in tab[1]]

    def same(a, b, loose=False):   # the estimates come back through latitude / longitude / height: one unit of the last printed digit
        assert len(a) == len(b)
        for la, lb in zip(a, b):
            fa, fb = la.split(), lb.split()
            assert len(fa) == len(fb), (la, lb)
            for x, y in zip(fa, fb):
                if x != y:
                    tol = 1.1 if "." not in x else (0.011 if len(x.split(".")[1]) == 2 else 1.1e-4)   # whole seconds of the .cor angles
                    if loose:      # deflection corrections are re-evaluated at the adjusted coordinates
                        tol *= 30
                    assert abs(float(x) - float(y)) < tol, (la, lb)
    assert body(again["adj"], "Adjusted Measurements") == body(first["adj"], "Adjusted Measurements")
    same(body(again["adj"], "Adjusted Coordinates"), body(first["adj"], "Adjusted Coordinates"))
    same(body(again["xyz"], "Adjusted Coordinates"), body(first["xyz"], "Adjusted Coordinates"))
    pick = lambda t: t.split("----------\n\n", 1)[1] if "----------\n\n" in t else t
    assert again["apu"].split("Positional uncertainty of adjusted")[1] == first["apu"].split("Positional uncertainty of adjusted")[1]
    same(again["cor"].split("Corrections to stations")[1].splitlines()[4:], first["cor"].split("Corrections to stations")[1].splitlines()[4:])
    for label in ("Chi squared", "Rigorous Sigma Zero", "Degrees of freedom", "Number of measurements"):
        assert re.findall(r"^" + label + r".*$", again["adj"], re.M) == re.findall(r"^" + label + r".*$", first["adj"], re.M)
    # a second adjustment starts from the files the first one updated (metadata `reduced`): measured values are restored
    # from preAdjMeas, geoid / deflection reductions and variance scalars are not applied twice (ADJ:296, 3913-3935)
    r = _run(exe, tmp_path, "rr", *flags)
    assert r.returncode == 0, r.stderr
    second = open(os.path.join(tmp_path, "rr.simult.adj")).read()
    sig = lambda t: float(re.findall(r"^Rigorous Sigma Zero\s+(\S+)", t, re.M)[-1])
    assert abs(sig(second) - sig(first["adj"])) < 2e-3
    same([l.replace("*", " ") for l in body(second, "Adjusted Measurements")], [l.replace("*", " ") for l in body(first["adj"], "Adjusted Measurements")],
         loose=True)
    # baselines in east / north / up need the full precision of the adjusted baselines: -pam.mtx
    r = _run(exe, tmp_path, "rr", "--output-adj-msr", "--output-adj-gnss-units", "1", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    enu = body(open(os.path.join(tmp_path, "rr.simult.adj")).read(), "Adjusted Measurements")
    r = _run(exe, tmp_path, "rr", "--max-iterations", "0", "--output-adj-msr", "--output-adj-gnss-units", "1")
    assert r.returncode == 0, r.stderr
    # (the run above started from the adjusted files: same solution to rounding)
    same([l.replace("*", " ") for l in body(open(os.path.join(tmp_path, "rr.simult.adj")).read(), "Adjusted Measurements")],
         [l.replace("*", " ") for l in enu])
    # without the files: a clear error
    os.remove(os.path.join(tmp_path, "rr-rva.mtx"))
    r = _run(exe, tmp_path, "rr", "--report-results")
    assert r.returncode == 1 and "Run an adjustment first" in r.stderr


def test_cli_report_results_hostsim(cli_hostsim, tmp_path):
    _report_results(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_report_results_gpu(cli_gpu, tmp_path):
    _report_results(cli_gpu, tmp_path)


def test_cli_output_json(cli_hostsim, tmp_path):
    """--output-json (DynAdjustJsonPrinter dnaadjust_json_printer.cpp): one JSON object per line beside the
    .adj / .xyz / .apu / .cor reports, the same numbers at full precision, keys as the reference names them."""
    import json
    stn, msr, _ = _network()
    _write_network(tmp_path, "js", stn, msr)
    r = _run(cli_hostsim, tmp_path, "js", "--output-adj-msr", "--output-json", "--output-pos-uncertainty", "--output-corrections-file",
             "--output-tstat-adj-msr")
    assert r.returncode == 0, r.stderr
    load = lambda e: [json.loads(l) for l in open(os.path.join(tmp_path, "js.simult." + e + ".jsonl"))]
    adj, xyz, apu, cor = load("adj"), load("xyz"), load("apu"), load("cor")
    for recs, kind in ((adj, "adj"), (xyz, "xyz"), (apu, "apu"), (cor, "cor")):
        h = recs[0]["DnaAdjustmentReport"]
        assert h["report"] == kind and h["type"] == "Adjustment" and h["referenceframe"] == "GDA2020"
    raw = open(os.path.join(tmp_path, "js.simult.adj.jsonl")).readline()
    assert raw.index('"epoch"') < raw.index('"referenceframe"') < raw.index('"report"') < raw.index('"type"') and ", " not in raw   # sorted keys, compact
    text = open(os.path.join(tmp_path, "js.simult.adj")).read()
    stt = adj[1]["DnaStatistics"]
    assert f'{stt["chisq"]:.2f}' == re.search(r"^Chi squared\s+(\S+)", text, re.M).group(1)
    assert f'{stt["sigma_zero"]:.3f}' == re.search(r"^Rigorous Sigma Zero\s+(\S+)", text, re.M).group(1)
    assert stt["dof"] == int(re.search(r"^Degrees of freedom\s+(\S+)", text, re.M).group(1)) and stt["chisq_test"] in ("passed", "warning", "failed")
    back = dnafiles.read_binary(os.path.join(tmp_path, "js.bms"), MSR_DTYPE)
    back = back[0] if isinstance(back, tuple) else back
    ms = [r["DnaMeasurement"] for r in adj if "DnaMeasurement" in r]
    used = back[back["ignore"] == 0]
    names = [n.decode() for n in stn["stationName"]]
    assert len(ms) == int(((used["measStart"] == 0) & np.isin(used["measType"], [b"G"])).sum()) + 4 + 6 + \
        int(np.isin(used["measType"], [b"S", b"A", b"L", b"V", b"H", b"E", b"M", b"B", b"K", b"Z"]).sum())
    s = next(m for m in ms if m["Type"] == "S")
    rec = used[(used["measType"] == b"S") & (used["station1"] == names.index(s["First"])) & (used["station2"] == names.index(s["Second"]))][0]
    assert s["Value"] == rec["term1"] and s["Adjusted"] == rec["measAdj"] and s["Correction"] == rec["measCorr"] and s["NStat"] == rec["NStat"]
    assert abs(s["TStat"] - s["NStat"] / np.sqrt(stt["sigma_zero"])) < 1e-12 and s["StdDev"] == np.sqrt(rec["term2"])
    d = next(m for m in ms if m["Type"] == "D")
    assert d["Total"] == len(d["Directions"]) and all("Target" in e for e in d["Directions"]) and abs(d["Directions"][0]["StdDev"] - 1.0) < 1e-9
    x = next(m for m in ms if m["Type"] == "X")
    assert x["Total"] == len(x["GPSBaseline"]) == len(x["Adjusted"]) and "GPSCovariance" in x["GPSBaseline"][0] and set(x["Adjusted"][0]) == {"X", "Y", "Z"}
    g = next(m for m in ms if m["Type"] == "G")
    assert set(g["Adjusted"]) == {"X", "Y", "Z"} and g["Total"] == 1                   # a single baseline collapses to one triple
    y = next(m for m in ms if m["Type"] == "Y")
    assert y["Coords"] == "XYZ" and "Second" not in y and "PointCovariance" in y["Clusterpoint"][0]
    sx = {r["DnaStation"]["Name"]: r["DnaStation"] for r in xyz[1:]}
    rows = {l.split()[0]: l.split() for l in _tables(open(os.path.join(tmp_path, "js.simult.xyz")).read(), "Adjusted Coordinates")[-1][1]}
    for name, f in rows.items():
        a, u = sx[name]["Adjusted"], sx[name]["Uncertainty"]
        assert np.abs(np.array([a["X"], a["Y"], a["Z"]]) - [float(v) for v in f[6:9]]).max() < 5.1e-5
        assert np.abs(np.array([u["SE"], u["SN"], u["SU"]]) - [float(v) for v in f[9:12]]).max() < 5.1e-5
        assert abs(a["Lat"] - float(f[2])) < 1e-9 and np.array(u["VarianceCart"]).shape == (3, 3)
    assert [r["DnaStation"]["Name"] for r in apu[1:]] == names and "HzPosU" in apu[1]["DnaStation"]["Uncertainty"]
    assert set(cor[1]["DnaStation"]["Corrections"]) == {"dE", "dN", "dUp"} and len(cor) == len(stn) + 1


def _non_convergence(exe, tmp_path):
    """An adjustment that runs out of iterations reports its iterations and "Failed to converge" only — no statistics, no
    tables (WRAP:1386-1390) — and still exits 0; a converged one lists the measurements beyond the critical n-statistic on
    the console (PrintSuspectMeasurementSummary ADJ:7652-7779)."""
    stn, msr, _, _ = synth.gnss_network(40, 110, 3)
    _write_network(tmp_path, "nc", stn, msr)
    r = _run(exe, tmp_path, "nc", "--max-iterations", "1", "--output-adj-msr")
    assert r.returncode == 0 and "failed to converge after 1 iteration" in r.stdout, r.stderr
    text = open(os.path.join(tmp_path, "nc.simult.adj")).read()
    assert re.search(r"^SOLUTION\s+Failed to converge", text, re.M) and len(re.findall(r"^ITERATION", text, re.M)) == 1
    assert "Adjusted Coordinates" not in text and "Chi squared" not in text and "Adjusted Measurements" not in text
    assert not os.path.exists(os.path.join(tmp_path, "nc-rva.mtx"))
    back = dnafiles.read_binary(os.path.join(tmp_path, "nc.bst"), stn.dtype)
    back = back[0] if isinstance(back, tuple) else back
    assert np.array_equal(back["currentLatitude"], stn["currentLatitude"])      # the binary files are left as they were
    r = _run(exe, tmp_path, "nc", "--output-adj-msr")
    assert r.returncode == 0, r.stderr
    text = open(os.path.join(tmp_path, "nc.simult.adj")).read()
    outliers = int(re.search(r"\((\d+) potential outlier", text).group(1))
    m = re.search(r"\+ Largest measurement N-statistics \((\d+) total, showing top (\d+)\):", r.stdout)
    assert m and int(m.group(1)) == outliers and int(m.group(2)) == min(outliers, 20)
    rows = re.findall(r"^  - G msr (\d+) cluster \d+ file-order \d+ (\S+) -> (\S+): N=(-?\d+\.\d+), corr=\S+, residual precision=\S+, Pelzer=\S+, exceeds critical$",
                      r.stdout, re.M)
    assert len(rows) == min(outliers, 20)
    ns = [abs(float(x[3])) for x in rows]
    assert ns == sorted(ns, reverse=True) and min(ns) > 1.95


def test_cli_non_convergence_and_suspect_summary_hostsim(cli_hostsim, tmp_path):
    _non_convergence(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_non_convergence_and_suspect_summary_gpu(cli_gpu, tmp_path):
    _non_convergence(cli_gpu, tmp_path)


def _database_ids(exe, tmp_path):
    """--output-database-ids (LoadDatabaseId ADJ:2211-2276, PrintMeasurementDatabaseID PRN:239-263): the measurement id of
    <net>.dbid beside every row, the cluster id as well for D G X Y; one list in file order whatever --output-msr-blocks says."""
    stn, msr, _ = _network()
    _write_network(tmp_path, "db", stn, msr)
    n = len(msr)
    ids = np.zeros(n, dtype=[("m", "<u4"), ("c", "<u4"), ("ms", "<u2"), ("cs", "<u2")])
    ids["m"], ids["c"], ids["ms"], ids["cs"] = 100000 + np.arange(n), 5000 + msr["clusterID"], 1, 1
    with open(os.path.join(tmp_path, "db.dbid"), "wb") as f:
        f.write(np.uint32(n).tobytes() + ids.tobytes())
    r = _run(exe, tmp_path, "db", "--output-adj-msr", "--output-database-ids", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    head, body = _tables(open(os.path.join(tmp_path, "db.simult.adj")).read(), "Adjusted Measurements")[-1]
    assert head.split()[-4:] == ["Meas.", "ID", "Clust.", "ID"]
    used = np.where((msr["ignore"] == 0) & (msr["measStart"] <= 2))[0]
    k = 0
    for l in body:
        f = l.split()
        if l[0] == "D" and len(f) == 6:                      # heading row: D inst RO count, ids of the set's first record
            rec = used[k]
            assert msr["measType"][rec] == b"D" and int(f[-2]) == 100000 + rec and int(f[-1]) == 5000 + msr["clusterID"][rec]
            k += 1
            continue
        rec = used[k]
        while msr["measType"][rec] == b"D" and msr["vectorCount1"][rec] > 0:
            k += 1
            rec = used[k]
        t = chr(msr["measType"][rec][0])
        if t in "DGXY":
            assert int(f[-2]) == 100000 + rec and int(f[-1]) == 5000 + msr["clusterID"][rec], (l, rec)
        else:
            assert int(f[-1]) == 100000 + rec and not f[-2].isdigit() or f[-2] == "*", (l, rec)
        k += 1
    assert k == len(used)
    os.remove(os.path.join(tmp_path, "db.dbid"))
    r = _run(exe, tmp_path, "db", "--output-adj-msr", "--output-database-ids", "--no-binary-update")
    assert r.returncode == 1 and "dbid" in r.stderr


def test_cli_database_ids_hostsim(cli_hostsim, tmp_path):
    _database_ids(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_database_ids_gpu(cli_gpu, tmp_path):
    _database_ids(cli_gpu, tmp_path)


def test_cli_writes_project_file(cli_hostsim, tmp_path):
    """After an adjustment <net>.dnaproj holds the settings of the run (dnaprojectfile.cpp:1703-2027; WRAP:1456-1466): the
    #adjust / #output sections are rewritten, sections of the other programs are kept, and `-p` on it repeats the run."""
    stn, msr, _, _ = synth.gnss_network(40, 110, 19)
    _write_network(tmp_path, "pw", stn, msr)
    proj = os.path.join(tmp_path, "pw.dnaproj")
    with open(proj, "w") as f:
        f.write("# pw project file. Created by dnaimport.\n\n\n" + f"{'#general (35)':<35}VALUE\n" + "-" * 80 + "\n" + f"{'network-name':<35}pw\n" +
                f"{'input-folder':<35}{tmp_path}\n" + f"{'output-folder':<35}{tmp_path}\n\n" + f"{'#import (35)':<35}VALUE\n" + "-" * 80 + "\n" +
                f"{'reference-frame':<35}GDA2020\n" + f"{'stn-msr-file':<35}pw.stn\n\n" + f"{'#plot (35)':<35}VALUE\n" + "-" * 80 + "\n\n")
    r = _run(cli_hostsim, tmp_path, "pw", "--output-adj-msr", "--sort-adj-msr-field", "7", "--free-stn-sd", "4", "--stn-coord-types", "ENzh",
             "--output-tstat-adj-msr", "--max-iterations", "12")
    assert r.returncode == 0 and "+ Open pw.simult.adj to view the adjustment details." in r.stdout, r.stderr
    text = open(proj).read()
    secs = re.findall(r"^(#\w+) \(35\)", text, re.M)
    assert secs == ["#general", "#import", "#adjust", "#output", "#plot"] and "stn-msr-file                       pw.stn" in text
    val = lambda k: re.search(r"^" + re.escape(k) + r"\s+(\S+)", text, re.M).group(1)
    assert val("adjustment-mode") == "simultaneous-adjustment" and val("max-iterations") == "12" and val("free-stn-sd") == "4.000"
    assert val("fixed-stn-sd") == "1.0000e-06" and val("output-adj-msr") == "yes" and val("sort-adj-msr-field") == "7" and val("stn-coord-types") == "ENzh"
    assert val("output-tstat-adj-msr") == "yes" and val("output-pos-uncertainty") == "no"
    first = open(os.path.join(tmp_path, "pw.simult.adj")).read()
    stn2, msr2, _, _ = synth.gnss_network(40, 110, 19)           # fresh binaries: -p repeats the same adjustment
    _write_network(tmp_path, "pw", stn2, msr2)
    r = subprocess.run([cli_hostsim, "-p", proj], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    again = open(os.path.join(tmp_path, "pw.simult.adj")).read()
    tab = lambda t, h: _tables(t, h)[-1]
    assert tab(again, "Adjusted Measurements") == tab(first, "Adjusted Measurements") and tab(again, "Adjusted Coordinates") == tab(first, "Adjusted Coordinates")
    assert re.findall(r"^(#\w+) \(35\)", open(proj).read(), re.M) == secs


def test_cli_sigint_cancels_between_iterations(cli_hostsim, tmp_path):
    """SIGINT asks for a graceful stop (WRAP:67-71; CancelAdjustment polled once per iteration, ADJ:2432): the iteration in
    flight finishes, the .adj records the iterations done and the status, the binary files stay untouched, exit code 0."""
    import signal
    stn, msr, _, _ = synth.gnss_network(4000, 12000, 41)      # an iteration on the CPU stand-in takes about a second
    _write_network(tmp_path, "sg", stn, msr)
    p = subprocess.Popen([cli_hostsim, "sg", "--input-folder", str(tmp_path), "--output-folder", str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         text=True)
    line = p.stdout.readline()                    # "+ Preparing for adjustment... " is flushed once the handler is installed
    while "Adjusting network" not in line and line:
        line = p.stdout.readline()
    p.send_signal(signal.SIGINT)                  # the first of the two iterations this network needs is under way
    out, err = p.communicate(timeout=600)
    assert p.returncode == 0, err
    assert "Adjustment cancelled by the user after 1 iteration" in out
    text = open(os.path.join(tmp_path, "sg.simult.adj")).read()
    assert re.search(r"^SOLUTION\s+Adjustment cancelled", text, re.M) and len(re.findall(r"^ITERATION", text, re.M)) == 1
    assert not os.path.exists(os.path.join(tmp_path, "sg-rva.mtx"))


def test_cli_stale_segmentation_file(cli_hostsim, tmp_path):
    """A default .seg older than station / measurement files that dnaimport has written since is refused (WRAP:1239-1265);
    naming it with --seg-file, or files last written by an adjustment, are fine."""
    import time
    stn, msr, _, _ = synth.gnss_network(40, 110, 23)
    isl = parity.chain_blocks(40, 14)
    dnafiles.write_seg(os.path.join(tmp_path, "sf.seg"), isl, [[] for _ in isl], [[] for _ in isl])
    past = time.time() - 3600
    os.utime(os.path.join(tmp_path, "sf.seg"), (past, past))
    _write_network(tmp_path, "sf", stn, msr)                      # "imported" after the segmentation
    r = _run(cli_hostsim, tmp_path, "sf", "--phased")
    assert r.returncode == 1 and "imported after" in r.stderr and "--seg-file" in r.stderr
    r = _run(cli_hostsim, tmp_path, "sf", "--phased", "--seg-file", "sf.seg")
    assert r.returncode == 0, r.stderr                             # this run updates the binaries: modified by "adjust"
    r = _run(cli_hostsim, tmp_path, "sf", "--phased")
    assert r.returncode == 0, r.stderr


def test_cli_nstat_sort_in_alternate_units(cli_hostsim, tmp_path):
    """--sort-adj-msr-field 7 with --output-adj-gnss-units 1/2/3 (the reference's CI runs these, CMakeLists.txt:1046-1060):
    baselines are ordered by the n-statistics of the frame they are printed in (PRN:1708-1712, 4526-4715)."""
    stn, msr, _, _ = synth.gnss_network(40, 110, 6)
    _write_network(tmp_path, "ns", stn, msr)
    for units in ("1", "2", "3"):
        r = _run(cli_hostsim, tmp_path, "ns", "--output-adj-msr", "--sort-adj-msr-field", "7", "--output-adj-gnss-units", units, "--scale-normals-to-unity",
                 "--no-binary-update")
        assert r.returncode == 0, r.stderr
        body = _tables(open(os.path.join(tmp_path, "ns.simult.adj")).read(), "Adjusted Measurements")[-1][1]
        assert len(body) == 330
        big = []
        for b in range(0, 330, 3):
            ns = []
            for l in body[b:b + 3]:
                f = l[67:].split()
                ns.append(abs(float(f[10] if not re.fullmatch(r"-?\d+\.\d+", f[0]) else f[6])))
            big.append(max(ns))
        assert all(a >= b - 0.0051 for a, b in zip(big, big[1:])), units


def _reference_ci_command_lines(exe, tmp_path):
    """Every option set the reference's CI drives dnaadjust with (CMakeLists.txt:1046-1963) runs to completion here on a
    mixed network; the ones its CI expects to fail, fail."""
    import shlex
    stn, msr, _ = _network()
    name = stn["stationName"][5].decode()
    isl = parity.chain_blocks(len(stn), 20)
    n = len(msr)
    ids = np.zeros(n, dtype=[("m", "<u4"), ("c", "<u4"), ("ms", "<u2"), ("cs", "<u2")])
    ids["m"], ids["c"], ids["ms"], ids["cs"] = np.arange(n), msr["clusterID"], 1, 1
    with open(os.path.join(tmp_path, "dsg.typeb"), "w") as f:
        f.write("!#=DNA 1.00 TBU\n" + f"{name:<20}0.010 0.010 0.030\n")

    def fresh():
        _write_network(tmp_path, "ci", stn.copy(), msr.copy())
        dnafiles.write_seg(os.path.join(tmp_path, "ci.seg"), isl, [[] for _ in isl], [[] for _ in isl])
        with open(os.path.join(tmp_path, "ci.dbid"), "wb") as f:
            f.write(np.uint32(n).tobytes() + ids.tobytes())
    ok = [
        "--output-adj-msr --scale-normals-to-unity",
        *[f"--output-adj-msr --sort-adj-msr-field 7 --output-adj-gnss-units {u} --scale-normals-to-unity" for u in range(4)],
        "--output-msr-to-stn", *[f"--output-msr-to-stn --sort-msr-to-stn-field {k}" for k in range(4)],
        "--block1-phased --output-adj-msr --output-adj-gnss-units 1",
        "--phased --output-adj-msr --output-adj-gnss-units 2 --output-iter-adj-msr",
        "--phased --output-adj-msr --sort-adj-msr-field 7 --output-adj-gnss-units 1",
        "--staged-adjustment --create-stage-files --output-adj-msr --output-adj-gnss-units 2 --sort-adj-msr-field 4",
        "--staged-adjustment --create-stage-files --output-adj-msr --sort-adj-msr-field 7 --output-adj-gnss-units 2",
        "--staged-adjustment --purge-stage-files --output-adj-msr --output-adj-gnss-units 3 --output-stn-blocks --output-msr-blocks --sort-adj-msr-field 5",
        "--verbose 5", "--staged --create --purge", "--staged", "--multi --output-adj-msr",
        "--output-adj-msr --phased --stn-corrections --export-sinex-file --export-xml-stn-file --export-dna-stn-file --output-pos-uncertainty "
        "--export-dna-msr --export-xml-msr",
        "--phased --staged-adjustment --create-stage-files --output-adj-msr --export-sinex-file --output-pos-uncertainty --export-xml-stn-file "
        "--export-xml-msr-file --export-dna-stn-file --export-dna-msr --output-iter-adj-stn --output-iter-adj-stat --output-iter-adj-msr "
        "--output-iter-cmp-msr --stn-corrections --output-corrections-file",
        "--output-adj-msr --free-stn-sd 5.0 --fixed-stn-sd 0.000005 --max-iterations 15 --output-tstat-adj-msr --sort-adj-msr-field 7 --sort-stn-orig-order "
        "--stn-coord-types XYZPLHhENz --angular-stn-type 1 --angular-msr-type 1 --precision-stn-linear 3 --precision-msr-linear 3 --precision-stn-angular 4 "
        "--precision-msr-angular 4 --output-pos-uncertainty --output-all-covariances --output-corrections-file",
        f'--type-b-sd-glob "0.01,0.01,0.035" --type-b-sd-file {os.path.join(tmp_path, "dsg.typeb")} --output-pos',
        '--comments "This is a comment that is quite lengthy in content and is relatively meaningless.  Feel free to delete this comment."',
        "--output-adj-msr --output-database-ids --output-ignored-msrs --sort-adj-msr-field 6 --output-msr-to-stn --sort-msr-to-stn-field 1 "
        "--output-iter-adj-stn --output-iter-adj-stat --output-iter-adj-msr --output-iter-cmp-msr",
        f'--constraint "{name},CCC"',
    ]
    for cmd in ok:
        fresh()
        r = _run(exe, tmp_path, "ci", *shlex.split(cmd))
        assert r.returncode == 0, (cmd, r.stderr)
    # the last run left updated binaries and the .mtx files: report mode on them (adjust-dbid-04 of the reference)
    r = _run(exe, tmp_path, "ci", *shlex.split(f'--report-results --output-adj-msr --output-pos-uncertainty --output-apu-vcv-units 1 --constraints "{name},CCC"'))
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "-p", os.path.join(tmp_path, "ci.dnaproj")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    for cmd in ('ci --constraints "no-name,CCC"', f'ci --constraints "{name},AAA"', "-x", "-p ./nofile.dnaproj", "--phased", "missing"):
        r = subprocess.run([exe, *shlex.split(cmd), "--input-folder", str(tmp_path), "--output-folder", str(tmp_path)], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 1 and "Error" in r.stderr, cmd
    for cmd in ("-h", "--version", "--help-module output"):
        r = subprocess.run([exe, *shlex.split(cmd)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and r.stdout, cmd


def test_cli_accepts_the_reference_ci_command_lines_hostsim(cli_hostsim, tmp_path):
    _reference_ci_command_lines(cli_hostsim, tmp_path)


@pytest.mark.gpu
def test_cli_accepts_the_reference_ci_command_lines_gpu(cli_gpu, tmp_path):
    _reference_ci_command_lines(cli_gpu, tmp_path)


