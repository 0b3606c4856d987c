"""Stress test for run-to-run reproducibility of the factorisation + selected inverse.

Prepares the C3 network once per ordering (nested dissection, 1000-station chain) and repeats the converged adjustment
many times in one process; every repetition's station variances are compared with the first repetition of the same
ordering.  Any difference beyond FP64 reordering noise is reported with the stations it touches."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynadjust_b200 import engine, synth  # noqa: E402
from tests import parity  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tag = sys.argv[3] if len(sys.argv) > 3 else "a"
only = sys.argv[4] if len(sys.argv) > 4 else "both"      # "nd", "chain" or "both"
stn, msr, truth, _ = synth.config_network(name)
thr = float(np.float32(0.0005))
bad = []
for label, kw, blocks in (("nd", dict(leaf_stations=96), None), ("chain", dict(), parity.chain_blocks(len(stn), 1000))):
    if only not in ("both", label):
        continue
    s, m = stn.copy(), msr.copy()
    adj = engine.Adjustment(s, m, **kw)
    if blocks is not None:
        adj.set_blocks(blocks)
    adj.prepare()
    ref = None
    t0 = time.time()
    for rep in range(reps):
        adj.reset_estimates()
        for it in range(10):
            r = adj.iterate(normals=True)
            if abs(r.max_corr) <= thr:
                break
        adj.form_inverse()
        q = adj.station_vcvs()
        e = adj.estimates()
        if ref is None:
            ref = (q, e)
            continue
        dq = np.abs(q - ref[0]).reshape(len(stn), -1).max(axis=1)
        scale = np.abs(ref[0]).max()
        if dq.max() > 1e-10 * scale or np.abs(e - ref[1]).max() > 1e-8:
            off = np.nonzero(dq > 1e-11 * scale)[0]
            rec = dict(tag=tag, ordering=label, rep=rep, dq_max=float(dq.max()), scale=float(scale), n_off=int(len(off)),
                       first=off[:40].tolist(), last=off[-10:].tolist(), d_est=float(np.abs(e - ref[1]).max()))
            print("MISMATCH", json.dumps(rec), flush=True)
            bad.append(rec)
    print(f"{tag} {label}: {reps} repetitions in {time.time() - t0:.1f} s, {len(bad)} mismatches so far", flush=True)
    adj.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(bad, open(f"gpurun_out/stress_{name}_{tag}.json", "w"), indent=1)
