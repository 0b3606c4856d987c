"""TEST INFRASTRUCTURE — reads the reference's ASCII DNA station / measurement files (format "DNA 3.01", the files
dnaimport consumes) into the binary record layouts, and applies the dnareftran step the reference's CI runs before
dnaadjust, for the measurement types the sample GNSS network holds (G baselines, X baseline clusters, Y point clusters).

It exists so that the reference's own end-to-end golden output (sampleData/gnss.simult.adj.expected, produced by
import -> geoid -> reftran -> adjust, CMakeLists.txt:1027-1034) can pin the CPU oracle and the CUDA path.  Never part of
the product.  Column layout: dnaimport's DNA parser (station 20, constraint 3, type 3, three 20-wide coordinates;
measurement: type, ignore flag, three 20-wide station fields, four 10-wide variance scalars, 20-wide frame and epoch;
value lines: 20-wide value at column 62 followed by 20-wide variance terms).  Frame transformation:
dnareftran.cpp:1740-1826 (a baseline is re-formed from its two transformed end points), parameters
ITRF2008 -> GDA2020 dnatransformationparameters.hpp:2413-2430, epoch arithmetic dnatemplatedatetimefuncs.hpp:292-327,
7-parameter transformation dnatemplatematrixfuncs.hpp:729-800."""
import numpy as np

from dynadjust_b200 import synth
from dynadjust_b200.records import GRS80_A, GRS80_INVF, LLH_TYPE, UTM_TYPE, XYZ_TYPE, new_msr, new_stn

# millimetres, ppb, milli-arc-seconds and their rates per year, reference epoch 2020.0
ITRF2008_TO_GDA2020 = (13.790, 4.550, 15.220, 2.5500, 0.2808, 0.2677, -0.4638,
                       1.420, 1.340, 0.900, 0.1090, 1.5461, 1.1820, 1.1551)
ITRF2014_TO_GDA2020 = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.50379, 1.18346, 1.20716)   # :1593-1614
GDA94_TO_GDA2020 = (61.55, -10.87, -40.19, -9.994, -39.4924, -32.7221, -32.8979, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)            # :2465-2481
PARAMETERS = {"ITRF2008": ITRF2008_TO_GDA2020, "ITRF2014": ITRF2014_TO_GDA2020, "GDA94": GDA94_TO_GDA2020}
CUMULATIVE_DAYS = ((0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334), (0, 31, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335))


def dms_to_rad(v):
    """ddd.mmssssss ("HP" notation of the DNA files) -> radians"""
    a = abs(v)
    d = np.floor(a + 1e-12)
    m = np.floor((a - d) * 100.0 + 1e-9)
    s = ((a - d) * 100.0 - m) * 100.0
    return np.sign(v) * np.radians(d + m / 60.0 + s / 3600.0)


def decimal_year(date):
    """referenceEpoch: year + (day of year - 0.5) / days in year, date = dd.mm.yyyy"""
    d, m, y = (int(x) for x in date.split("."))
    leap = 1 if (y % 400 == 0 or (y % 100 != 0 and y % 4 == 0)) else 0
    return y + (CUMULATIVE_DAYS[leap][m - 1] + d - 0.5) / (366.0 if leap else 365.0)


def helmert_to_gda2020(xyz, frame, epoch):
    """Transform_7parameter with the parameters reduced to the measurement's epoch (non-rigorous rotation matrix)."""
    p = PARAMETERS[frame.upper()]
    dt = decimal_year(epoch) - 2020.0
    t = np.array([p[0] + p[7] * dt, p[1] + p[8] * dt, p[2] + p[9] * dt]) / 1000.0
    sc = (p[3] + p[10] * dt) / 1e9
    rx, ry, rz = (np.radians((p[4 + k] + p[11 + k] * dt) / 3600.0) / 1000.0 for k in range(3))
    R = np.array([[1.0, rz, -ry], [-rz, 1.0, rx], [ry, -rx, 1.0]])
    return (R @ xyz) * (1.0 + sc) + t


def grid_to_geo(easting, northing, zone, a=GRS80_A, invf=GRS80_INVF, false_e=500000.0, false_n=10000000.0, k0=0.9996,
                lcm_z1=-177.0, zw=6.0):
    """GridToGeo (dnatemplategeodesyfuncs.hpp:435-510): Redfearn's formulae, UTM / MGA grid -> latitude, longitude (radians)"""
    f = 1.0 / invf
    b = a * (1 - f)
    e2 = 2 * f - f * f
    n = (a - b) / (a + b)
    n2, n3, n4 = n ** 2, n ** 3, n ** 4
    G = a * (1 - n) * (1 - n2) * (1 + 9 * n2 / 4 + 225 * n4 / 64) * (np.pi / 180.0)
    ep, npr = easting - false_e, northing - false_n
    m = npr / k0
    sigma = (m * np.pi) / (180 * G)
    lp = sigma + ((3 * n / 2) - (27 * n3 / 32)) * np.sin(2 * sigma)
    lp += ((21 * n2 / 16) - (55 * n4 / 32)) * np.sin(4 * sigma)
    lp += (151 * n3 / 96) * np.sin(6 * sigma)
    lp += (1097 * n4 / 512) * np.sin(8 * sigma)
    rho = a * (1 - e2) / (1 - e2 * np.sin(lp) ** 2) ** 1.5
    nu = a / (1 - e2 * np.sin(lp) ** 2) ** 0.5
    t, psi = np.tan(lp), nu / rho
    num1 = t / (k0 * rho)
    x = ep / (k0 * nu)
    t1 = num1 * x * ep / 2
    t2 = num1 * ep * x ** 3 / 24 * (-4 * psi ** 2 + 9 * psi * (1 - t ** 2) + 12 * t ** 2)
    t3 = num1 * ep * x ** 5 / 720 * (8 * psi ** 4 * (11 - 24 * t ** 2) - 12 * psi ** 3 * (21 - 71 * t ** 2) +
                                       15 * psi ** 2 * (15 - 98 * t ** 2 + 15 * t ** 4) + 180 * psi * (5 * t ** 2 - 3 * t ** 4) + 360 * t ** 4)
    t4 = num1 * ep * x ** 7 / 40320 * (1385 + 3633 * t ** 2 + 4095 * t ** 4 + 1575 * t ** 6)
    lat = lp - t1 + t2 - t3 + t4
    cm = np.radians(zone * zw + lcm_z1 - zw)
    sec = 1.0 / np.cos(lp)
    l1 = x * sec
    l2 = x ** 3 / 6 * sec * (psi + 2 * t ** 2)
    l3 = x ** 5 / 120 * sec * (-4 * psi ** 3 * (1 - 6 * t ** 2) + psi ** 2 * (9 - 68 * t ** 2) + 72 * psi * t ** 2 + 24 * t ** 4)
    l4 = x ** 7 / 5040 * sec * (61 + 662 * t ** 2 + 1320 * t ** 4 + 720 * t ** 6)
    return lat, cm + l1 - l2 + l3 - l4


def apply_geoid(stn, path, convert_heights=True):
    """dnageoid (dnageoid.cpp:165-176) from an exported DNA geoid file (name, N, meridian and prime-vertical deflections
    in seconds): geoid separation and deflections into the station records; orthometric station heights become
    ellipsoidal (--convert-stn-hts).  The exported file carries three decimals (the sample's .gsb grid does not
    reproduce the values of the expected run — they differ by 2-4 mm — so the exported file is the record of what that
    run used; its rounding, 0.5 mm in N, is the accuracy limit of the height-dependent rows of this fixture)."""
    index = {n.decode(): i for i, n in enumerate(stn["stationName"])}
    for l in open(path).read().splitlines():
        if not l or l.startswith("#"):
            continue
        f = l.split()
        i = index.get(f[0])
        if i is None:
            continue
        stn["geoidSep"][i] = float(f[1])
        stn["meridianDef"][i] = np.radians(float(f[2]) / 3600.0)
        stn["verticalDef"][i] = np.radians(float(f[3]) / 3600.0)
        if convert_heights:
            stn["currentHeight"][i] = stn["initialHeight"][i] + float(stn["geoidSep"][i])


def reftran_stations(stn, frame, epoch="01.01.2020"):
    """dnareftran on the station file: every station from `frame` to GDA2020.  Run before the geoid step, as in the
    reference's urban_mt pipeline (CMakeLists.txt:1078-1079): horizontal position through the 7-parameter transformation."""
    for i in range(len(stn)):
        x = synth.geo_to_cart(stn["initialLatitude"][i:i + 1], stn["initialLongitude"][i:i + 1], stn["initialHeight"][i:i + 1])[0]
        lat, lon, h = synth.cart_to_geo(helmert_to_gda2020(x, frame, epoch)[None, :])
        stn["initialLatitude"][i] = stn["currentLatitude"][i] = lat[0]
        stn["initialLongitude"][i] = stn["currentLongitude"][i] = lon[0]
        if stn["suppliedStationType"][i] == XYZ_TYPE:
            stn["initialHeight"][i] = stn["currentHeight"][i] = h[0]
        # else: an orthometric height does not depend on the reference frame and stays as supplied (the expected run keeps
        # the height-constrained station 1042 at its H)


def read_stations(path):
    rows = [l for l in open(path).read().splitlines() if l and not l.startswith(("!", "*"))]
    stn = new_stn(len(rows))
    for i, l in enumerate(rows):
        name, const, ctype = l[0:20].strip(), l[20:23], l[24:27]
        c = [float(l[27 + 20 * k:47 + 20 * k]) for k in range(3)]
        if ctype == "XYZ":
            lat, lon, h = synth.cart_to_geo(np.array([c]))
            lat, lon, h = float(lat[0]), float(lon[0]), float(h[0])
            stn["suppliedStationType"][i] = XYZ_TYPE
        elif ctype == "UTM":
            zone = int(l[87:90])
            lat, lon = grid_to_geo(c[0], c[1], zone)
            h = c[2]
            stn["suppliedStationType"][i] = UTM_TYPE
        else:
            lat, lon, h = dms_to_rad(c[0]), dms_to_rad(c[1]), c[2]
            stn["suppliedStationType"][i] = LLH_TYPE
        stn["stationName"][i] = stn["stationNameOrig"][i] = name.encode()
        stn["stationConst"][i] = const.encode()
        stn["stationType"][i] = ctype.encode()
        stn["initialLatitude"][i] = stn["currentLatitude"][i] = lat
        stn["initialLongitude"][i] = stn["currentLongitude"][i] = lon
        stn["initialHeight"][i] = stn["currentHeight"][i] = h
        stn["description"][i] = l[90 if ctype == "UTM" else 87:].strip().encode()[:127] if len(l) > 87 else b""
        stn["fileOrder"][i] = stn["nameOrder"][i] = i
        stn["epoch"][i] = b"01.01.2020"
    return stn


def _values(line):
    v = float(line[62:82])
    terms = [float(line[82 + 20 * k:102 + 20 * k]) for k in range(3) if line[82 + 20 * k:102 + 20 * k].strip()]
    return v, terms


def _terms(line):
    return [float(line[82 + 20 * k:102 + 20 * k]) for k in range(3)]


def read_measurements(path, stn, reftran=True):
    """Measurements of a DNA file -> binary records (dnaimport): G baselines, X / Y clusters, the scalar types A B K V Z
    (angles: radians, variances radians^2) and H L M S C E (metres), optionally followed by the reference-frame step
    (dnareftran) for GNSS vectors given in ITRF2008 / ITRF2014."""
    index = {n.decode(): i for i, n in enumerate(stn["stationName"])}
    xyz0 = synth.geo_to_cart(stn["currentLatitude"], stn["currentLongitude"], stn["currentHeight"])
    L = [l for l in open(path).read().splitlines() if l and not l.startswith(("!", "*"))]
    groups, cluster, i = [], 0, 0
    while i < len(L):
        l = L[i]
        kind, ignore = l[0], l[1] == "*"
        if kind == "G":
            s1, s2 = index[l[2:22].strip()], index[l[22:42].strip()]
            v, p, lam, h = (float(l[62 + 10 * k:72 + 10 * k]) for k in range(4))
            frame, epoch = l[102:122].strip(), l[122:142].strip()
            vals = [_values(L[i + 1 + k]) for k in range(3)]
            d = np.array([x[0] for x in vals])
            if reftran and frame.upper() != "GDA2020":
                a = xyz0[s1]
                d = helmert_to_gda2020(a + d, frame, epoch) - helmert_to_gda2020(a, frame, epoch)
            m = new_msr(3)
            m["measType"], m["measStart"], m["measurementStations"], m["coordType"] = b"G", [0, 1, 2], 2, b"XYZ"
            m["station1"], m["station2"], m["ignore"] = s1, s2, ignore
            m["vectorCount1"], m["vectorCount2"], m["clusterID"] = 1, 0, cluster
            m["term1"] = d
            m["term2"] = [vals[0][1][0], vals[1][1][0], vals[2][1][0]]
            m["term3"] = [0.0, vals[1][1][1], vals[2][1][1]]
            m["term4"] = [0.0, 0.0, vals[2][1][2]]
            m["scale1"], m["scale2"], m["scale3"], m["scale4"] = p, lam, h, v
            m["epoch"] = b"01.01.2020"
            groups.append(m)
            cluster += 1
            i += 4
        elif kind in "ABKVZ" or kind in "HLMSCE":
            f = l[62:].split()
            m = new_msr(1)
            m["measType"], m["ignore"] = kind.encode(), ignore
            m["station1"] = index[l[2:22].strip()]
            nst = 1
            if l[22:42].strip():
                m["station2"], nst = index[l[22:42].strip()], 2
            if l[42:62].strip():
                m["station3"], nst = index[l[42:62].strip()], 3
            m["measurementStations"] = nst
            if kind in "ABKVZ":          # degrees minutes seconds, standard deviation in seconds
                sign = -1.0 if f[0].startswith("-") else 1.0
                m["term1"] = sign * np.radians(abs(float(f[0])) + float(f[1]) / 60.0 + float(f[2]) / 3600.0)
                m["term2"] = np.radians(float(f[3]) / 3600.0) ** 2
                rest = f[4:]
            else:                        # metres
                m["term1"] = float(f[0])
                m["term2"] = float(f[1]) ** 2
                rest = f[2:]
            if len(rest) >= 2:
                m["term3"], m["term4"] = float(rest[0]), float(rest[1])   # instrument and target heights
            m["clusterID"] = cluster
            m["epoch"] = b"01.01.2020"
            groups.append(m)
            cluster += 1
            i += 1
        elif kind in "XY":
            count = int(l[42:62])
            v, p, lam, h = (float(l[62 + 10 * k:72 + 10 * k]) for k in range(4))
            frame, epoch = l[102:122].strip(), l[122:142].strip()
            ctype = l[22:42].strip() if kind == "Y" else "XYZ"
            if kind == "Y":
                assert ctype in ("XYZ", "LLH", "LLh"), "point cluster form not handled"
            total = sum(3 + 3 * (count - 1 - k) for k in range(count))
            m = new_msr(total)
            m["measType"], m["coordType"], m["clusterID"], m["ignore"] = kind.encode(), ctype.encode(), cluster, ignore
            m["measurementStations"] = 2 if kind == "X" else 1
            m["scale1"], m["scale2"], m["scale3"], m["scale4"] = p, lam, h, v
            m["epoch"] = b"01.01.2020"
            o = 0
            for k in range(count):
                s = index[L[i][2:22].strip()]
                vals = [_values(L[i + 1 + c]) for c in range(3)]
                r = m[o:o + 3]
                r["measStart"], r["station1"], r["vectorCount1"], r["vectorCount2"] = [0, 1, 2], s, count, count - 1 - k
                d = np.array([x[0] for x in vals])
                if ctype != "XYZ":
                    d[0], d[1] = dms_to_rad(d[0]), dms_to_rad(d[1])     # latitude and longitude arrive as ddd.mmssss
                    if reftran and kind == "Y" and frame.upper() != "GDA2020":    # the point itself goes to the new frame
                        x = synth.geo_to_cart(d[0:1], d[1:2], d[2:3])[0]
                        la, lo, hh = synth.cart_to_geo(helmert_to_gda2020(x, frame, epoch)[None, :])
                        d = np.array([la[0], lo[0], hh[0]])
                elif reftran and kind == "Y" and frame.upper() != "GDA2020":
                    d = helmert_to_gda2020(d, frame, epoch)
                if kind == "X":
                    r["station2"] = index[L[i][22:42].strip()]
                    if reftran and frame.upper() != "GDA2020":
                        d = helmert_to_gda2020(xyz0[s] + d, frame, epoch) - helmert_to_gda2020(xyz0[s], frame, epoch)
                r["term1"] = d
                r["term2"] = [vals[0][1][0], vals[1][1][0], vals[2][1][0]]
                r["term3"] = [0.0, vals[1][1][1], vals[2][1][1]]
                r["term4"] = [0.0, 0.0, vals[2][1][2]]
                o += 3
                i += 4
                for j in range(k + 1, count):
                    cv = m[o:o + 3]
                    cv["measStart"] = [3, 4, 5]
                    rows = [_terms(L[i + c]) for c in range(3)]
                    cv["term1"], cv["term2"], cv["term3"] = [x[0] for x in rows], [x[1] for x in rows], [x[2] for x in rows]
                    o += 3
                    i += 3
            # the covariance records carry the stations of the later member (dnagpsbaseline.cpp / dnagpspoint.cpp layout)
            firsts = [q for q in range(total) if m["measStart"][q] == 0]
            for a_, q in enumerate(firsts):
                pos = q + 3
                for b_ in range(a_ + 1, count):
                    m["station1"][pos:pos + 3] = m["station1"][firsts[b_]]
                    m["station2"][pos:pos + 3] = m["station2"][firsts[b_]]
                    pos += 3
            groups.append(m)
            cluster += 1
        else:
            raise ValueError(f"measurement type '{kind}' is not handled by this reader")
    msr = new_msr(sum(len(g) for g in groups))   # (np.concatenate would repack the padded record layout)
    o = 0
    for g in groups:
        msr[o:o + len(g)] = g
        o += len(g)
    msr["fileOrder"] = np.arange(len(msr), dtype=np.uint32)
    return msr
