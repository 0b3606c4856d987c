"""GEMM tile shapes on one workload: 128 x 128 everywhere, 64 x 64 wherever allowed, the planner's per-launch choice.

For each mode: prepare, warm up, one profiled iteration (per-launch CSV: index, kind, flops, tiles, ms, tag, level),
timed iterations, and the results against the 128 x 128 run (largest coordinate difference, largest relative
difference of the station variance blocks) plus the normal-equation identity on every station.

    python tools/tile_experiment.py [C4] [leaf] [modes, e.g. 128,64,0,0:256]   -> gpurun_out/tile_<mode>.csv, tile_experiment.json
A mode "0:<steps>" runs the planner's choice with GADJ_TILE_LONG_K=<steps> (16-deep K steps from which 128-wide tiles are kept).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from dynadjust_b200 import checks, engine, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["128", "64", "0"]
steps = int(os.environ.get("TILE_EXP_STEPS", "3"))
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)

stn0, msr0, _, _ = synth.config_network(cfg)
results = {}
ref = None
for mode in modes:
    tile, _, long_k = mode.partition(":")
    os.environ.pop("GADJ_TILE_LONG_K", None)
    if long_k:
        os.environ["GADJ_TILE_LONG_K"] = long_k
    stn, msr = stn0.copy(), msr0.copy()
    adj = engine.Adjustment(stn, msr, leaf_stations=leaf, gemm_tile=int(tile))
    t = time.time()
    info = adj.prepare()
    prepare_s = time.time() - t
    for _ in range(2):
        adj.reset_estimates()
        adj.iterate(normals=True, inverse=True)
    csv = os.path.join(out_dir, "tile_%s.csv" % mode.replace(":", "_"))
    if os.path.exists(csv):
        os.remove(csv)
    adj.profile_enable(True)
    adj.profile_read(reset=True)
    adj.reset_estimates()
    adj.iterate(normals=True, inverse=True)
    os.environ["GADJ_PROFILE_DUMP"] = csv
    p = adj.profile_read(reset=True)
    os.environ.pop("GADJ_PROFILE_DUMP")
    adj.profile_enable(False)
    ms = []
    for _ in range(steps):
        adj.reset_estimates()
        r = adj.iterate(normals=True, inverse=True)
        ms.append(r.ms_assemble + r.ms_factor + r.ms_solve + r.ms_inverse)
    est = adj.estimates()
    vcv = adj.station_vcvs()
    rec = dict(mode=mode, prepare_s=round(prepare_s, 2), ms_per_iteration=[round(x, 2) for x in ms],
               phases=dict(assemble=r.ms_assemble, factor=r.ms_factor, solve=r.ms_solve, inverse=r.ms_inverse),
               profiled=dict(gemm_ms=p.ms_gemm, diag_ms=p.ms_diag, gather_ms=p.ms_gather, transpose_ms=p.ms_transpose,
                             gemm_launches=int(p.gemm_launches), gemm_tiles=int(p.gemm_tiles), gemm_flops=p.flops_gemm))
    try:
        res, worst = checks.normal_identity_residual(adj, stn, msr)
        rec["normal_identity_max"] = res
    except AssertionError as e:   # not a GNSS-only network
        rec["normal_identity_max"] = None
    if ref is None:
        ref = (est, vcv)
    else:
        scale = np.abs(ref[1]).reshape(len(vcv), -1).max(axis=1)
        rec["vs_first_mode"] = dict(max_abs_dx=float(np.abs(est - ref[0]).max()),
                                    max_rel_dvcv=float((np.abs(vcv - ref[1]).reshape(len(vcv), -1).max(axis=1) / scale).max()))
    results[mode] = rec
    print(json.dumps(rec), flush=True)
    adj.close()
    del adj
with open(os.path.join(out_dir, "tile_experiment.json"), "w") as f:
    json.dump(dict(workload=cfg, leaf=leaf, results=results), f, indent=1)
