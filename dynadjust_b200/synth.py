"""Seeded synthetic geodetic networks written straight into DynAdjust binary records.

``dnaimport`` cannot be built offline (Xerces-C / XSD / Boost), so the workloads
named in BASELINE.json are generated here in the record layout ``dnaimport``
would have produced: stations on a jittered grid
over an Australia-sized lat/lon box, GNSS baselines ``G`` between grid
neighbours plus optional long "CORS hub" baselines, Cartesian VCVs rotated
from a local (e,n,up) error model.
"""
import numpy as np

from .records import GRS80_A, GRS80_INVF, LLH_TYPE, new_msr, new_stn

LAT_MIN, LAT_MAX = np.radians(-44.0), np.radians(-10.0)
LON_MIN, LON_MAX = np.radians(113.0), np.radians(154.0)

# neighbour offsets ordered by length (each undirected grid edge appears once)
_OFFSETS = [(1, 0), (0, 1), (1, 1), (1, -1), (2, 0), (0, 2), (2, 1), (1, 2), (2, -1), (1, -2),
            (2, 2), (2, -2), (3, 0), (0, 3), (3, 1), (1, 3), (3, -1), (1, -3)]


def ellipsoid(a=GRS80_A, invf=GRS80_INVF):
    b = a * (1.0 - 1.0 / invf)
    e2 = (a * a - b * b) / (a * a)
    return a, b, e2


def geo_to_cart(lat, lon, h, a=GRS80_A, invf=GRS80_INVF):
    _, _, e2 = ellipsoid(a, invf)
    nu = a / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
    x = (nu + h) * np.cos(lat) * np.cos(lon)
    y = (nu + h) * np.cos(lat) * np.sin(lon)
    z = (nu * (1.0 - e2) + h) * np.sin(lat)
    return np.stack([x, y, z], axis=-1)


def cart_to_geo(xyz, a=GRS80_A, invf=GRS80_INVF):
    """Vectorised fixed-point iteration; good to < 1e-13 rad for terrestrial points."""
    _, _, e2 = ellipsoid(a, invf)
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    p = np.hypot(x, y)
    lon = np.arctan2(y, x)
    lat = np.arctan2(z, p * (1.0 - e2))
    for _ in range(8):
        nu = a / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
        lat = np.arctan2(z + e2 * nu * np.sin(lat), p)
    nu = a / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
    h = p / np.cos(lat) - nu
    return lat, lon, h


def local_to_cart_rotation(lat, lon):
    """R with columns (east, north, up) expressed in the Cartesian frame; shape (..., 3, 3)."""
    sl, cl = np.sin(lat), np.cos(lat)
    so, co = np.sin(lon), np.cos(lon)
    R = np.empty(lat.shape + (3, 3))
    R[..., 0, 0] = -so
    R[..., 0, 1] = -sl * co
    R[..., 0, 2] = cl * co
    R[..., 1, 0] = co
    R[..., 1, 1] = -sl * so
    R[..., 1, 2] = cl * so
    R[..., 2, 0] = 0.0
    R[..., 2, 1] = cl
    R[..., 2, 2] = sl
    return R


def _grid_shape(n_stations):
    aspect = (LON_MAX - LON_MIN) / (LAT_MAX - LAT_MIN)
    nx = max(1, int(np.ceil(np.sqrt(n_stations * aspect))))
    ny = int(np.ceil(n_stations / nx))
    return nx, ny


def _box(n_stations):
    """Lat/lon box: the Australia-sized box at 1M stations, shrunk about its centre for smaller
    networks so that station spacing stays what it is at 1M stations (~3.6 km).  Keeping the
    spacing fixed keeps the baseline error model (3 mm + 0.5 ppm) and hence the conditioning of
    the normals comparable across the BASELINE.json configurations."""
    f = min(1.0, np.sqrt(n_stations / 1.0e6))
    latc, lonc = 0.5 * (LAT_MIN + LAT_MAX), 0.5 * (LON_MIN + LON_MAX)
    hlat, hlon = 0.5 * f * (LAT_MAX - LAT_MIN), 0.5 * f * (LON_MAX - LON_MIN)
    return latc - hlat, latc + hlat, lonc - hlon, lonc + hlon


def grid_edges(n_stations, n_edges, rng):
    """Undirected neighbour edges on the station grid, exactly ``n_edges`` of them (shortest offsets first)."""
    nx, ny = _grid_shape(n_stations)
    idx = np.arange(n_stations, dtype=np.int64)
    ix, iy = idx % nx, idx // nx
    chunks, total = [], 0
    for dx, dy in _OFFSETS:
        jx, jy = ix + dx, iy + dy
        ok = (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny)
        j = jy * nx + jx
        ok &= j < n_stations
        e = np.stack([idx[ok], j[ok]], axis=1)
        if total + len(e) >= n_edges:
            keep = n_edges - total
            sel = np.sort(rng.choice(len(e), size=keep, replace=False))
            chunks.append(e[sel])
            total += keep
            break
        chunks.append(e)
        total += len(e)
    if total < n_edges:
        raise ValueError(f"cannot place {n_edges} grid edges on {n_stations} stations")
    return np.concatenate(chunks, axis=0)


def gnss_network(n_stations, n_baselines, seed, hub_fraction=0.0, n_hubs=0, apriori_sigma=0.5,
                 n_fixed=3, a=GRS80_A, invf=GRS80_INVF):
    """Stations + ``G`` baselines.  Returns (stn, msr, truth_xyz, edges)."""
    rng = np.random.default_rng(seed)
    nx, ny = _grid_shape(n_stations)
    idx = np.arange(n_stations, dtype=np.int64)
    ix, iy = idx % nx, idx // nx
    jit = rng.uniform(-0.3, 0.3, size=(n_stations, 2))
    lat0, lat1, lon0, lon1 = _box(n_stations)
    lat = lat0 + (iy + 0.5 + jit[:, 1]) / ny * (lat1 - lat0)
    lon = lon0 + (ix + 0.5 + jit[:, 0]) / nx * (lon1 - lon0)
    h = rng.uniform(0.0, 1000.0, size=n_stations)
    truth = geo_to_cart(lat, lon, h, a, invf)

    n_hub_edges = int(round(hub_fraction * n_baselines)) if n_hubs > 0 else 0
    edges = grid_edges(n_stations, n_baselines - n_hub_edges, rng)
    if n_hub_edges:
        hubs = np.sort(rng.choice(n_stations, size=n_hubs, replace=False))
        src = rng.integers(0, n_stations, size=n_hub_edges)
        dst = hubs[rng.integers(0, n_hubs, size=n_hub_edges)]
        clash = src == dst
        src[clash] = (src[clash] + 1) % n_stations
        edges = np.concatenate([edges, np.stack([src, dst], axis=1)], axis=0)
    flip = rng.random(len(edges)) < 0.5
    edges[flip] = edges[flip][:, ::-1]
    m = len(edges)

    # fixed stations: spread over the index range, a-priori = truth
    fixed = np.unique(np.linspace(0, n_stations - 1, n_fixed).astype(np.int64)) if n_fixed else np.array([], np.int64)
    apri = truth + rng.normal(0.0, apriori_sigma, size=truth.shape)
    apri[fixed] = truth[fixed]
    alat, alon, ah = cart_to_geo(apri, a, invf)

    stn = new_stn(n_stations)
    stn["stationName"] = np.char.add("S", np.char.zfill(idx.astype(str), 7)).astype("S31")
    stn["stationNameOrig"] = stn["stationName"]
    stn["stationConst"] = b"FFF"
    stn["stationConst"][fixed] = b"CCC"
    stn["stationType"] = b"LLH"
    stn["suppliedStationType"] = LLH_TYPE
    stn["initialLatitude"] = stn["currentLatitude"] = alat
    stn["initialLongitude"] = stn["currentLongitude"] = alon
    stn["initialHeight"] = stn["currentHeight"] = ah
    stn["suppliedHeightRefFrame"] = 1          # ELLIPSOIDAL_type_i: the initial heights above are ellipsoidal
    stn["geoidSep"] = (30.0 * np.sin(3.0 * alat) * np.cos(2.0 * alon)).astype(np.float32)
    stn["geoidSepUnc"] = 0.05
    stn["fileOrder"] = idx
    stn["nameOrder"] = idx
    stn["epoch"] = b"01.01.2020"

    s1, s2 = edges[:, 0], edges[:, 1]
    d_true = truth[s2] - truth[s1]
    length = np.linalg.norm(d_true, axis=1)
    sh = 0.003 + 0.5e-6 * length
    sig = np.stack([sh, sh, 3.0 * sh], axis=1)                      # e, n, up
    R = local_to_cart_rotation(lat[s1], lon[s1])
    V = np.einsum("mij,mj,mkj->mik", R, sig ** 2, R)
    noise = np.einsum("mij,mj->mi", R, sig * rng.standard_normal((m, 3)))
    obs = d_true + noise

    msr = new_msr(3 * m)
    rec = msr.reshape(m, 3)
    rec["measType"] = b"G"
    rec["measStart"] = np.array([0, 1, 2], dtype=np.int8)[None, :]
    rec["measurementStations"] = 2
    rec["coordType"] = b"XYZ"
    rec["epoch"] = b"01.01.2020"
    rec["station1"] = s1[:, None]
    rec["station2"] = s2[:, None]
    rec["vectorCount1"] = 1
    rec["vectorCount2"] = 0
    rec["clusterID"] = np.arange(m, dtype=np.uint32)[:, None]
    rec["fileOrder"] = np.arange(3 * m, dtype=np.uint32).reshape(m, 3)
    rec["term1"] = obs
    rec["term2"][:, 0] = V[:, 0, 0]
    rec["term2"][:, 1] = V[:, 0, 1]
    rec["term3"][:, 1] = V[:, 1, 1]
    rec["term2"][:, 2] = V[:, 0, 2]
    rec["term3"][:, 2] = V[:, 1, 2]
    rec["term4"][:, 2] = V[:, 2, 2]
    return stn, msr, truth, edges


def mixed_network(n_stations, n_baselines, seed, n_distances=0, n_levels=0, **kw):
    """GNSS network plus terrestrial rows between grid neighbours: slope distances 'S' (sigma 2 mm + 2 ppm,
    instrument / target heights ~1.5 m) and levelled height differences 'L' (sigma 1 mm * sqrt(km), orthometric,
    reduced by adjust with the stations' geoid separations).  The terrestrial records follow the baselines."""
    stn, gmsr, truth, edges = gnss_network(n_stations, n_baselines, seed, **kw)
    rng = np.random.default_rng(seed + 7919)
    a, invf = kw.get("a", GRS80_A), kw.get("invf", GRS80_INVF)
    _, _, e2 = ellipsoid(a, invf)
    tlat, tlon, th = cart_to_geo(truth, a, invf)
    pairs = grid_edges(n_stations, max(n_distances, n_levels, 1), rng)
    recs = []
    if n_distances:
        pr = pairs[rng.choice(len(pairs), size=n_distances, replace=False)]
        s1, s2 = pr[:, 0], pr[:, 1]
        ih = rng.uniform(1.2, 1.8, n_distances)
        tg = rng.uniform(1.2, 1.8, n_distances)
        # the adjustment model rotates both heights at station 1 (CartesianElementsFromInstrumentHeight)
        up1 = np.stack([np.cos(tlat[s1]) * np.cos(tlon[s1]), np.cos(tlat[s1]) * np.sin(tlon[s1]), np.sin(tlat[s1])], axis=1)
        d = np.linalg.norm(truth[s2] - truth[s1] + up1 * (tg - ih)[:, None], axis=1)
        sig = 0.002 + 2.0e-6 * d
        m = new_msr(n_distances)
        m["measType"] = b"S"
        m["measurementStations"] = 2
        m["station1"], m["station2"] = s1, s2
        m["term1"] = d + sig * rng.standard_normal(n_distances)
        m["term2"] = sig ** 2
        m["term3"], m["term4"] = ih, tg
        recs.append(m)
    if n_levels:
        pr = pairs[rng.choice(len(pairs), size=n_levels, replace=False)]
        s1, s2 = pr[:, 0], pr[:, 1]

        def ell_height(i):
            nu = a / np.sqrt(1.0 - e2 * np.sin(tlat[i]) ** 2)
            zn = e2 * nu * np.sin(tlat[i])
            return np.sqrt(truth[i, 0] ** 2 + truth[i, 1] ** 2 + (truth[i, 2] + zn) ** 2) - nu
        dh = ell_height(s2) - ell_height(s1)
        dist_km = np.linalg.norm(truth[s2] - truth[s1], axis=1) / 1000.0
        sig = 0.001 * np.sqrt(np.maximum(dist_km, 0.05))
        geoid = stn["geoidSep"].astype(np.float64)      # float32 in the record: use exactly what adjust will read
        m = new_msr(n_levels)
        m["measType"] = b"L"
        m["measurementStations"] = 2
        m["station1"], m["station2"] = s1, s2
        m["term1"] = dh - (geoid[s2] - geoid[s1]) + sig * rng.standard_normal(n_levels)
        m["term2"] = sig ** 2
        recs.append(m)
    # (np.concatenate would repack the padded record dtype: copy into a fresh array of the exact layout instead)
    msr = new_msr(len(gmsr) + sum(len(r) for r in recs))
    msr[:len(gmsr)] = gmsr
    o = len(gmsr)
    for r in recs:
        msr[o:o + len(r)] = r
        o += len(r)
    msr["fileOrder"] = np.arange(len(msr), dtype=np.uint32)
    return stn, msr, truth, edges


# BASELINE.json configurations: seeds 1234 + config index
CONFIGS = {
    "C1": dict(n_stations=100, n_baselines=300, seed=1235),
    "C2": dict(n_stations=10_000, n_baselines=30_000, seed=1236),
    "C3g": dict(n_stations=100_000, n_baselines=300_000, seed=1237),
    # BASELINE config C3: ~3 GNSS baselines per station, direction sets of 4-6 targets on 30 % of the
    # stations, slope distances and levelled height differences along neighbour lines
    "C3": dict(n_stations=100_000, n_baselines=300_000, seed=1237,
               terrestrial=dict(scalars={"S": 120_000, "L": 100_000}, n_dir_sets=30_000)),
    "C4": dict(n_stations=1_000_000, n_baselines=10_000_000, seed=1238, hub_fraction=0.02, n_hubs=200),
    "C5": dict(n_stations=100_000, n_baselines=300_000, seed=1239),
}


def config_network(name):
    cfg = dict(CONFIGS[name])
    terr = cfg.pop("terrestrial", None)
    if terr is not None:
        from . import synth_terrestrial
        return synth_terrestrial.terrestrial_network(cfg.pop("n_stations"), cfg.pop("n_baselines"), cfg.pop("seed"), **terr, **cfg)
    return gnss_network(**cfg)
