"""Iteration time of one workload for several nested-dissection leaf sizes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynadjust_b200 import engine, synth
cfg = sys.argv[1]
leaves = [int(x) for x in sys.argv[2].split(",")]
stn, msr, _, _ = synth.config_network(cfg)
out = {}
for leaf in leaves:
    adj = engine.Adjustment(stn.copy(), msr, leaf_stations=leaf)
    t = time.time(); info = adj.prepare(); prep = time.time() - t
    best = None
    for _ in range(3):
        adj.reset_estimates()
        r = adj.iterate(normals=True, inverse=True)
        tot = r.ms_assemble + r.ms_factor + r.ms_solve + r.ms_inverse
        if best is None or tot < best[0]:
            best = (tot, r.ms_assemble, r.ms_factor, r.ms_solve, r.ms_inverse)
    out[leaf] = dict(total_ms=best[0], assemble=best[1], factor=best[2], solve=best[3], inverse=best[4], prepare_s=prep,
                     fronts=info.nfronts, panel_gb=info.panel_bytes / 1e9, pool_gb=info.pool_bytes / 1e9,
                     factor_tf=info.factor_flops / 1e12, inverse_tf=info.inverse_flops / 1e12)
    print(leaf, json.dumps(out[leaf]), flush=True)
    adj.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"leaf_sweep_{cfg}.json"), "w"), indent=1)
