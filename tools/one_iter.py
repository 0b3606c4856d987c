"""One full iteration (assemble + factorise + solve + selected inverse) of a workload — the command ncu wraps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynadjust_b200 import engine, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 128
stn, msr, _, _ = synth.config_network(cfg)
adj = engine.Adjustment(stn, msr, leaf_stations=leaf)
adj.prepare()
r = adj.iterate(normals=True, inverse=True)
print("phases", r.ms_assemble, r.ms_factor, r.ms_solve, r.ms_inverse)
