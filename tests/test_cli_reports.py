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
from dynadjust_b200 import synth_terrestrial as st
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
