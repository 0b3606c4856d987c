"""Full CPU runs of the reference's path for the BASELINE configs that fit a host run: C2 (10k stations, dense normals
n = 30 000, simultaneous: assembly + packed Cholesky inverse through the compiled reference matrix_2d + solve) and C3
(100k stations, mixed types, phased over a chain of 1000-station blocks: forward, reverse and combination passes with
dense per-block inverses).  Written once per machine to profiles/ with the log; bench.py quotes these figures beside
its own bounded sample."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynadjust_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import parity  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "C2"
threads = os.cpu_count() or 1
out = dict(config=which, cores=threads, kind="reference" if pyoracle.ref_loaded() else "port", host=os.uname().nodename,
           cpu=open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else "?")
stn, msr, truth, _ = synth.config_network(which)
o = pyoracle.default_opts(threads=threads)
t0 = time.time()
if which == "C2":
    res = pyoracle.adjust_simultaneous(stn, msr, opts=o)
    r = res["res"]
    out.update(mode="simultaneous, dense normals n = %d" % (3 * len(stn)), iterations=int(r.iterations), wall_s=time.time() - t0,
               seconds_prepare=r.seconds_prepare, seconds_solve=r.seconds_solve, seconds_inverse=r.seconds_inverse,
               sigma_zero=r.sigma_zero, rms_vs_truth=float(np.sqrt(((res["est"] - truth) ** 2).mean())),
               # one iteration = assembly + (inverse + solve) of the first iteration; later iterations of a GNSS-only network reuse the inverse
               ms_per_iteration=1e3 * (r.seconds_prepare + r.seconds_solve))
else:
    blocks = parity.chain_blocks(len(stn), 1000)
    res = pyoracle.adjust_phased(stn, msr, blocks, opts=o)
    r = res["res"]
    out.update(mode="phased, %d blocks of 1000 inner stations" % len(blocks), iterations=int(r.iterations), wall_s=time.time() - t0,
               seconds_prepare=r.seconds_prepare, seconds_solve=r.seconds_solve, seconds_inverse=r.seconds_inverse,
               sigma_zero=r.sigma_zero, rms_vs_truth=float(np.sqrt(((res["est"] - truth) ** 2).mean())),
               ms_per_iteration=1e3 * r.seconds_solve / max(1, r.iterations))
    np.savez_compressed(f"/tmp/phased_{which}.npz", est=res["est"], vcv=res["vcv"])
    # golden fixture for the GPU parity test at this size: every 20th station's adjusted coordinates and rigorous 3x3 block
    idx = np.arange(0, len(stn), 20)
    np.savez_compressed(os.path.join("tests", "golden", f"{which.lower()}_phased_oracle_sample.npz"), stations=idx,
                        est=res["est"][idx], vcv=res["vcv"][idx], sigma_zero=r.sigma_zero, chi_squared=r.chi_squared, dof=r.dof,
                        iterations=r.iterations, outliers=r.outliers, block_width=1000)
print(json.dumps(out), flush=True)
json.dump(out, open(f"profiles/r2_cpu_baseline_{which.lower()}.json", "w"), indent=1)
