"""Diagnostic: is the C3 adjustment reproducible run to run, and where do two elimination orders differ?

Runs the C3 network under nested dissection and under a 1000-station chain, twice each, prints the largest correction of
every iteration at full precision, and compares estimates / station variances between all pairs of runs."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynadjust_b200 import engine, synth  # noqa: E402
from tests import parity  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
stn, msr, truth, _ = synth.config_network(name)
thr = float(np.float32(0.0005))
runs = []
log = []
for label, kw, blocks in (("nd", dict(leaf_stations=96), None), ("chain", dict(), parity.chain_blocks(len(stn), 1000))):
    for rep in range(reps):
        s, m = stn.copy(), msr.copy()
        adj = engine.Adjustment(s, m, **kw)
        if blocks is not None:
            adj.set_blocks(blocks)
        t0 = time.time()
        info = adj.prepare()
        tp = time.time() - t0
        corr = []
        for it in range(10):
            r = adj.iterate(normals=(it == 0))
            corr.append(r.max_corr)
            if abs(r.max_corr) <= thr:
                break
        adj.form_inverse()
        st = adj.statistics(write_back=False)
        e, q = adj.estimates(), adj.station_vcvs()
        # a second inverse from a fresh factorisation at the converged estimates: same linearisation point for every run
        r2 = adj.iterate(normals=True, inverse=True)
        q2 = adj.station_vcvs()
        e2 = adj.estimates()
        runs.append((f"{label}{rep}", e, q, q2, e2, st.sigma_zero))
        rec = dict(run=f"{label}{rep}", prepare_s=tp, fronts=int(info.nfronts), levels=int(info.nlevels), iters=len(corr),
                   max_corr=[repr(c) for c in corr], extra_corr=repr(r2.max_corr), sigma0=repr(st.sigma_zero))
        print(json.dumps(rec), flush=True)
        log.append(rec)
        adj.close()

qmax = np.abs(runs[0][2]).max()
for i in range(len(runs)):
    for j in range(i + 1, len(runs)):
        a, b = runs[i], runs[j]
        de = np.abs(a[1] - b[1]).max()
        dq = np.abs(a[2] - b[2]).reshape(len(stn), -1).max(axis=1)
        dq2 = np.abs(a[3] - b[3]).reshape(len(stn), -1).max(axis=1)
        rel = dq / np.maximum(np.abs(a[2]).reshape(len(stn), -1).max(axis=1), 1e-300)
        rel2 = dq2 / np.maximum(np.abs(a[3]).reshape(len(stn), -1).max(axis=1), 1e-300)
        worst = np.argsort(dq)[-5:][::-1]
        rec = dict(pair=f"{a[0]}-{b[0]}", d_est=de, d_est2=float(np.abs(a[4] - b[4]).max()), d_sigma0=abs(a[5] - b[5]),
                   dq_max=float(dq.max()), dq_rel_to_qmax=float(dq.max() / qmax),
                   dq_rel_max=float(rel.max()), n_rel_gt_1e9=int((rel > 1e-9).sum()), n_rel_gt_1e6=int((rel > 1e-6).sum()),
                   worst_stations=[int(w) for w in worst], worst_dq=[float(dq[w]) for w in worst],
                   dq2_max=float(dq2.max()), dq2_rel_to_qmax=float(dq2.max() / qmax), dq2_rel_max=float(rel2.max()),
                   n2_rel_gt_1e9=int((rel2 > 1e-9).sum()))
        print(json.dumps(rec), flush=True)
        log.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(log, open(f"gpurun_out/diag_{name}.json", "w"), indent=1)
