"""Synthetic terrestrial / clustered measurements on top of a GNSS network (BASELINE config C3 and the type tests).

Every observed value is computed here from the TRUE coordinates with NumPy formulas written for this generator
(east-north-up rotation + arctan2, foot points on the ellipsoid, ...), not with the adjustment model's code, so that
an adjustment that converges to the truth at sigma-zero ~ 1 checks the model end to end.  Record layouts follow
what dnaimport writes (measurement_types/dna*.cpp ::WriteBinaryMsr).

Types: A horizontal angle, B geodetic azimuth, K astronomic azimuth, C chord, E ellipsoid arc, M mean-sea-level arc,
S slope distance, V zenith distance, Z vertical angle, L level difference, H orthometric height, R ellipsoidal height,
I / J astronomic latitude / longitude, P / Q geodetic latitude / longitude, D direction sets, X baseline clusters,
Y point clusters (Cartesian).
"""
import numpy as np

from .records import GRS80_A, GRS80_INVF, new_msr
from . import synth

SEC = np.pi / 180.0 / 3600.0


def _enu(truth, lat, lon, s1, d):
    """east, north, up components at station s1 of Cartesian vectors d."""
    sl, cl, so, co = np.sin(lat[s1]), np.cos(lat[s1]), np.sin(lon[s1]), np.cos(lon[s1])
    e = -so * d[:, 0] + co * d[:, 1]
    n = -sl * co * d[:, 0] - sl * so * d[:, 1] + cl * d[:, 2]
    u = cl * co * d[:, 0] + cl * so * d[:, 1] + sl * d[:, 2]
    return e, n, u


def _up(lat, lon, s):
    return np.stack([np.cos(lat[s]) * np.cos(lon[s]), np.cos(lat[s]) * np.sin(lon[s]), np.sin(lat[s])], axis=1)


class Truth:
    def __init__(self, stn, truth, a=GRS80_A, invf=GRS80_INVF):
        self.xyz = truth
        self.a = a
        _, _, self.e2 = synth.ellipsoid(a, invf)
        self.lat, self.lon, self.h = synth.cart_to_geo(truth, a, invf)
        self.N = stn["geoidSep"].astype(np.float64)
        self.dM = stn["meridianDef"].astype(np.float64)
        self.dV = stn["verticalDef"].astype(np.float64)

    def azimuth(self, s1, s2):
        e, n, _ = _enu(self.xyz, self.lat, self.lon, s1, self.xyz[s2] - self.xyz[s1])
        return np.mod(np.arctan2(e, n), 2 * np.pi)

    def sight(self, s1, s2, ih, th):
        """instrument (s1 + ih up) -> target (s2 + th up) in the local frame of s1: e, n, u."""
        d = self.xyz[s2] + _up(self.lat, self.lon, s2) * th[:, None] - self.xyz[s1] - _up(self.lat, self.lon, s1) * ih[:, None]
        return _enu(self.xyz, self.lat, self.lon, s1, d)

    def zenith(self, s1, s2, ih, th):
        e, n, u = self.sight(s1, s2, ih, th)
        return np.arctan2(np.hypot(e, n), u)

    def foot(self, s):
        return synth.geo_to_cart(self.lat[s], self.lon[s], np.zeros(len(s)), self.a)

    def chord(self, s1, s2):
        return np.linalg.norm(self.foot(s2) - self.foot(s1), axis=1)

    def nu_rho(self, lat):
        d = np.sqrt(1.0 - self.e2 * np.sin(lat) ** 2)
        return self.a / d, self.a * (1.0 - self.e2) / d ** 3

    def ell_arc(self, s1, s2):
        c = self.chord(s1, s2)
        nu, rho = self.nu_rho(0.5 * (self.lat[s1] + self.lat[s2]))
        az = self.azimuth(s1, s2)
        r = rho * nu / (nu * np.cos(az) ** 2 + rho * np.sin(az) ** 2)
        return 2.0 * r * np.arcsin(c / (2.0 * r))

    def msl_arc(self, s1, s2):
        c = self.chord(s1, s2)
        nu, rho = self.nu_rho(0.5 * (self.lat[s1] + self.lat[s2]))
        rm = np.sqrt(nu * rho)
        n1, n2 = self.N[s1], self.N[s2]
        msl_chord = np.sqrt(c ** 2 * (1.0 + n1 / rm) * (1.0 + n2 / rm) + (n2 - n1) ** 2)
        r = rm + 0.5 * (n1 + n2)
        return 2.0 * r * np.arcsin(msl_chord / (2.0 * r))

    def direction_deflection(self, s1, az, zen):
        return (self.dM[s1] * np.sin(az) - self.dV[s1] * np.cos(az)) / np.tan(zen)


def add_deflections(stn, rng, fraction=0.5, scale_sec=5.0):
    """Deflections of the vertical (a few arc seconds) on a share of the stations."""
    n = len(stn)
    has = rng.random(n) < fraction
    stn["meridianDef"] = np.where(has, rng.normal(0.0, scale_sec, n) * SEC, 0.0)
    stn["verticalDef"] = np.where(has, rng.normal(0.0, scale_sec, n) * SEC, 0.0)


def _pairs(n_stations, count, rng):
    pool = synth.grid_edges(n_stations, max(count, 1), rng)
    pr = pool[rng.choice(len(pool), size=count, replace=False)]
    flip = rng.random(count) < 0.5
    pr[flip] = pr[flip][:, ::-1]
    return pr[:, 0], pr[:, 1]


def _scalar(kind, s1, s2, value, sigma, rng, ih=None, th=None, s3=None, nstn=2, noise=True):
    m = new_msr(len(s1))
    m["measType"] = kind.encode()
    m["measurementStations"] = nstn
    m["station1"] = s1
    if s2 is not None:
        m["station2"] = s2
    if s3 is not None:
        m["station3"] = s3
    m["term1"] = value + (sigma * rng.standard_normal(len(s1)) if noise else 0.0)
    m["term2"] = sigma ** 2
    if ih is not None:
        m["term3"], m["term4"] = ih, th
    return m


def scalar_measurements(stn, truth, kind, count, rng, noise=True):
    """`count` records of one scalar type between grid neighbours (or on single stations)."""
    T = Truth(stn, truth)
    n = len(stn)
    ones = np.ones(count)
    if kind in "HRIJPQ":
        s1 = rng.choice(n, size=count, replace=count > n)
        if kind == "H":
            return _scalar(kind, s1, None, T.h[s1] - T.N[s1], 0.01 * ones, rng, nstn=1, noise=noise)
        if kind == "R":
            return _scalar(kind, s1, None, T.h[s1], 0.01 * ones, rng, nstn=1, noise=noise)
        sig = 0.3 * SEC * ones       # astronomic / geodetic positions: ~0.3 arc seconds (~10 m on the ground)
        if kind == "P":
            return _scalar(kind, s1, None, T.lat[s1], sig, rng, nstn=1, noise=noise)
        if kind == "Q":
            return _scalar(kind, s1, None, T.lon[s1], sig, rng, nstn=1, noise=noise)
        if kind == "I":
            return _scalar(kind, s1, None, T.lat[s1] + T.dM[s1], sig, rng, nstn=1, noise=noise)
        return _scalar(kind, s1, None, T.lon[s1] + T.dV[s1] / np.cos(T.lat[s1]), sig, rng, nstn=1, noise=noise)
    s1, s2 = _pairs(n, count, rng)
    ih, th = rng.uniform(1.2, 1.8, count), rng.uniform(1.2, 1.8, count)
    dist = np.linalg.norm(truth[s2] - truth[s1], axis=1)
    if kind == "S":
        # the adjustment model rotates both heights at station 1 (CartesianElementsFromInstrumentHeight)
        d = np.linalg.norm(truth[s2] - truth[s1] + _up(T.lat, T.lon, s1) * (th - ih)[:, None], axis=1)
        return _scalar(kind, s1, s2, d, 0.002 + 2.0e-6 * d, rng, ih, th, noise=noise)
    if kind == "C":
        return _scalar(kind, s1, s2, T.chord(s1, s2), 0.002 + 2.0e-6 * dist, rng, noise=noise)
    if kind == "E":
        return _scalar(kind, s1, s2, T.ell_arc(s1, s2), 0.002 + 2.0e-6 * dist, rng, noise=noise)
    if kind == "M":
        return _scalar(kind, s1, s2, T.msl_arc(s1, s2), 0.002 + 2.0e-6 * dist, rng, noise=noise)
    if kind == "L":
        dh = T.h[s2] - T.h[s1]
        sig = 0.001 * np.sqrt(np.maximum(dist / 1000.0, 0.05))
        return _scalar(kind, s1, s2, dh - (T.N[s2] - T.N[s1]), sig, rng, noise=noise)
    ang_sig = 1.0 * SEC * ones
    az = T.azimuth(s1, s2)
    zen = T.zenith(s1, s2, ih, th)
    if kind == "B":
        return _scalar(kind, s1, s2, az, ang_sig, rng, noise=noise)
    if kind == "K":
        lap = T.dV[s1] * np.tan(T.lat[s1]) + T.direction_deflection(s1, az, zen)
        return _scalar(kind, s1, s2, np.mod(az + lap, 2 * np.pi), ang_sig, rng, ih, th, noise=noise)
    zcorr = T.dM[s1] * np.cos(az) + T.dV[s1] * np.sin(az)
    if kind == "V":
        return _scalar(kind, s1, s2, zen - zcorr, ang_sig, rng, ih, th, noise=noise)
    if kind == "Z":
        e, nn, u = T.sight(s1, s2, ih, th)
        return _scalar(kind, s1, s2, np.arctan2(u, np.hypot(e, nn)) + zcorr, ang_sig, rng, ih, th, noise=noise)
    if kind == "A":
        # third station: another neighbour of s1
        nx, _ = synth._grid_shape(n)
        s3 = s1 + np.where((s1 % nx) + 2 < nx, 2, -2)
        s3 = np.where((s3 == s2) | (s3 < 0) | (s3 >= n), (s1 + nx) % n, s3)
        s3 = np.where((s3 == s2) | (s3 == s1), (s1 + 2 * nx + 1) % n, s3)
        az3 = T.azimuth(s1, s3)
        zen3 = T.zenith(s1, s3, ih, th)
        corr = T.direction_deflection(s1, az3, zen3) - T.direction_deflection(s1, az, zen)
        return _scalar(kind, s1, s2, np.mod(az3 - az + corr, 2 * np.pi), np.sqrt(2.0) * ang_sig, rng, ih, th, s3=s3, nstn=3,
                       noise=noise)
    raise ValueError(kind)


def direction_sets(stn, truth, n_sets, rng, targets=(4, 6), sigma_sec=1.0, ignore_some=False, noise=True):
    """Direction sets 'D': RO record + one record per further target (dnadirectionset.cpp:430-466)."""
    T = Truth(stn, truth)
    n = len(stn)
    nx, ny = synth._grid_shape(n)
    offs = [(1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (1, -1), (-1, -1), (2, 0), (0, 2), (-2, 0), (0, -2)]
    inst = rng.choice(n, size=n_sets, replace=n_sets > n)
    recs = []
    for k, i in enumerate(inst):
        ix, iy = i % nx, i // nx
        cand = []
        for dx, dy in offs:
            jx, jy = ix + dx, iy + dy
            j = jy * nx + jx
            if 0 <= jx < nx and 0 <= jy < ny and j < n:
                cand.append(j)
        nt = min(len(cand), int(rng.integers(targets[0], targets[1] + 1)))
        if nt < 2:
            continue
        tg = np.array(cand)[rng.choice(len(cand), size=nt, replace=False)]
        s1 = np.full(nt, i)
        ih, th = np.zeros(nt), np.zeros(nt)
        az = T.azimuth(s1, tg)
        zen = T.zenith(s1, tg, ih, th)
        order = np.argsort(az)                       # a round of directions is observed clockwise
        tg, az, zen = tg[order], az[order], zen[order]
        omega = rng.uniform(0, 2 * np.pi)            # unknown orientation of the circle
        sig = sigma_sec * SEC
        d = np.mod(az + T.direction_deflection(s1, az, zen) + omega + (sig * rng.standard_normal(nt) if noise else 0.0), 2 * np.pi)
        m = new_msr(nt)
        m["measType"] = b"D"
        m["measurementStations"] = 2
        m["station1"] = i
        m["station2"] = tg
        m["term1"] = d
        m["term2"] = sig ** 2
        m["measStart"] = 1
        m["measStart"][0] = 0
        m["clusterID"] = k
        ign = np.zeros(nt, bool)
        if ignore_some and nt >= 4 and k % 3 == 0:
            ign[2] = True                            # an ignored direction inside the set (ADJ:5120-5129)
        m["ignore"] = ign
        m["vectorCount1"][0] = nt
        m["vectorCount2"][0] = nt - int(ign.sum())
        recs.append(m)
    return recs


def _cluster_vcv(Vb, rho, rng):
    """Full SPD VCV of a cluster from per-member 3x3 blocks and a common correlation rho between members."""
    k = len(Vb)
    L = np.linalg.cholesky(Vb)
    V = np.zeros((3 * k, 3 * k))
    for i in range(k):
        for j in range(k):
            V[3 * i:3 * i + 3, 3 * j:3 * j + 3] = Vb[i] if i == j else rho * L[i] @ L[j].T
    return V


def _cluster_records(kind, s1, s2, obs, V, cluster_id):
    """3 records per member + 3 covariance records per later member (dnagpsbaseline.cpp:421-493)."""
    k = len(s1)
    total = sum(3 + 3 * (k - 1 - i) for i in range(k))
    m = new_msr(total)
    m["measType"] = kind.encode()
    m["coordType"] = b"XYZ"
    m["clusterID"] = cluster_id
    m["measurementStations"] = 2 if kind == "X" else 1
    o = 0
    for i in range(k):
        r = m[o:o + 3]
        r["measStart"] = [0, 1, 2]
        r["station1"] = s1[i]
        if s2 is not None:
            r["station2"] = s2[i]
        r["vectorCount1"] = k
        r["vectorCount2"] = k - 1 - i
        r["term1"] = obs[3 * i:3 * i + 3]
        B = V[3 * i:3 * i + 3, 3 * i:3 * i + 3]
        r["term2"] = [B[0, 0], B[0, 1], B[0, 2]]
        r["term3"] = [0.0, B[1, 1], B[1, 2]]
        r["term4"] = [0.0, 0.0, B[2, 2]]
        o += 3
        for j in range(i + 1, k):
            cv = m[o:o + 3]
            cv["measStart"] = [3, 4, 5]
            cv["station1"] = s1[j]
            if s2 is not None:
                cv["station2"] = s2[j]
            Cb = V[3 * i:3 * i + 3, 3 * j:3 * j + 3]
            cv["term1"], cv["term2"], cv["term3"] = Cb[:, 0], Cb[:, 1], Cb[:, 2]
            o += 3
    return m


def gnss_clusters(stn, truth, kind, n_clusters, rng, members=(2, 5), rho=0.3, v_scale=1.0, noise=True):
    """'X' baseline clusters / 'Y' point clusters (Cartesian) with a full VCV."""
    T = Truth(stn, truth)
    n = len(stn)
    recs = []
    for c in range(n_clusters):
        k = int(rng.integers(members[0], members[1] + 1))
        if kind == "X":
            hub = int(rng.integers(0, n))
            s1, s2 = _pairs(n, k, rng)
            s1[: k // 2] = hub                       # some baselines share a station, as session solutions do
            s2 = np.where(s2 == s1, (s1 + 1) % n, s2)
            d = truth[s2] - truth[s1]
            sh = 0.003 + 0.5e-6 * np.linalg.norm(d, axis=1)
        else:
            s1 = rng.choice(n, size=k, replace=False)
            s2 = None
            d = truth[s1]
            sh = np.full(k, 0.004)
        sig = np.stack([sh, sh, 3.0 * sh], axis=1)
        R = synth.local_to_cart_rotation(T.lat[s1], T.lon[s1])
        Vb = np.einsum("mij,mj,mkj->mik", R, sig ** 2, R)
        V = _cluster_vcv(Vb, rho, rng)
        obs = d.reshape(-1) + (np.linalg.cholesky(V) @ rng.standard_normal(3 * k) if noise else 0.0)
        m = _cluster_records(kind, s1, s2, obs, V / v_scale, c)
        m["scale4"] = v_scale
        recs.append(m)
    return recs


def assemble(gmsr, groups):
    """Concatenate record groups behind the GNSS baselines into one .bms-ordered array."""
    total = len(gmsr) + sum(len(g) for g in groups)
    msr = new_msr(total)
    msr[:len(gmsr)] = gmsr
    o = len(gmsr)
    for g in groups:
        msr[o:o + len(g)] = g
        o += len(g)
    msr["fileOrder"] = np.arange(total, dtype=np.uint32)
    return msr


def terrestrial_network(n_stations, n_baselines, seed, scalars=None, n_dir_sets=0, n_x=0, n_y=0, deflections=True,
                        ignore_some=False, noise=True, v_scale=1.0, **kw):
    """GNSS backbone + the requested terrestrial / cluster measurements.  scalars: {'A': count, 'B': count, ...}."""
    stn, gmsr, truth, edges = synth.gnss_network(n_stations, n_baselines, seed, **kw)
    rng = np.random.default_rng(seed + 104729)
    if deflections:
        add_deflections(stn, rng)
    groups = []
    for kind, count in (scalars or {}).items():
        if count:
            groups.append(scalar_measurements(stn, truth, kind, count, rng, noise=noise))
    if n_dir_sets:
        groups += direction_sets(stn, truth, n_dir_sets, rng, ignore_some=ignore_some, noise=noise)
    if n_x:
        groups += gnss_clusters(stn, truth, "X", n_x, rng, v_scale=v_scale, noise=noise)
    if n_y:
        groups += gnss_clusters(stn, truth, "Y", n_y, rng, v_scale=v_scale, noise=noise)
    return stn, assemble(gmsr, groups), truth, edges
