"""Per-launch device timings of one full iteration (kind, algorithmic flops, tiles, ms) -> gpurun_out/<name>.csv"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynadjust_b200 import engine, synth
name = sys.argv[1] if len(sys.argv) > 1 else "prof"
cfg = sys.argv[2] if len(sys.argv) > 2 else "C4"
leaf = int(sys.argv[3]) if len(sys.argv) > 3 else 256
out = os.path.join(ROOT, "gpurun_out", name + ".csv")
if os.path.exists(out):
    os.remove(out)
stn, msr, _, _ = synth.config_network(cfg)
adj = engine.Adjustment(stn, msr, leaf_stations=leaf)
info = adj.prepare()
for _ in range(2):
    adj.reset_estimates(); adj.iterate(normals=True, inverse=True)
adj.profile_enable(True); adj.profile_read(reset=True)
adj.reset_estimates()
t = time.time(); r = adj.iterate(normals=True, inverse=True); wall = time.time() - t
os.environ["GADJ_PROFILE_DUMP"] = out
p = adj.profile_read(reset=True)
print("wall_ms", wall * 1e3, "phases", r.ms_assemble, r.ms_factor, r.ms_solve, r.ms_inverse, "gemm_ms", p.ms_gemm, "diag", p.ms_diag)
