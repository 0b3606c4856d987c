"""Small chain + nested-dissection adjustment for compute-sanitizer runs (racecheck / synccheck / memcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynadjust_b200 import engine, synth  # noqa: E402
from tests import parity  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 400
stn, msr, truth, _ = synth.gnss_network(n, 3 * n, 77)
out = []
for kw, blocks in ((dict(leaf_stations=48), None), (dict(), parity.chain_blocks(n, width))):
    s, m = stn.copy(), msr.copy()
    adj, info, last, stats = parity.run_engine(engine.LIB_PATH, s, m, blocks=blocks, **kw)
    out.append((adj.estimates(), adj.station_vcvs(), stats.sigma_zero))
    print("fronts", info.nfronts, "levels", info.nlevels, "iters", last.iteration, "sigma0", stats.sigma_zero, flush=True)
    adj.close()
print("d_est", np.abs(out[0][0] - out[1][0]).max(), "d_vcv_rel", np.abs(out[0][1] - out[1][1]).max() / np.abs(out[0][1]).max())
