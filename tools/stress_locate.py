"""Reproducibility stress test, second stage: repeat the converged nested-dissection adjustment in one process and, when a
repetition's station variances differ from the first one beyond rounding noise, say which buffers differ and where
(front, level, rows, columns) — W / Wt pivot-block inverses (left by the factorisation) and the inverse panels."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynadjust_b200 import engine, synth  # noqa: E402
from dynadjust_b200.multigpu import _CudaView  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
tag = sys.argv[3] if len(sys.argv) > 3 else "a"
leaf = int(sys.argv[4]) if len(sys.argv) > 4 else 96
stn, msr, truth, _ = synth.config_network(name)
thr = float(np.float32(0.0005))
dump = f"/tmp/fronts_{tag}.csv"
os.environ["GADJ_DUMP_FRONTS"] = dump
adj = engine.Adjustment(stn.copy(), msr.copy(), leaf_stations=leaf)
adj.prepare()
fr = np.loadtxt(dump, delimiter=",")
lvl, k, r, poff, ldk = fr[:, 0].astype(int), fr[:, 1].astype(np.int64), fr[:, 2].astype(np.int64), fr[:, 4].astype(np.int64), fr[:, 5].astype(np.int64)
ldw = k + (k & 1)
wblock = (k * ldw + 15) & ~15
woff = np.concatenate([[0], np.cumsum(2 * wblock)])[:-1]


def buf(which):
    ptr, cnt = C.c_void_p(), C.c_uint64()
    adj._check(adj.L.gadj_mg_buffer(adj.h, which, C.byref(ptr), C.byref(cnt)))
    return torch.as_tensor(_CudaView(ptr.value, cnt.value, "<f8"), device="cuda").cpu().numpy().copy()


def front_report(p, pr):
    """Per front: largest difference of the Z11 block (lower triangle) and of the Z21 block relative to the block's largest
    entry; fronts beyond 1e-10 are reported with the rows / columns of the differing entries."""
    out = []
    for fi in range(len(k)):
        kk, rr, ld = int(k[fi]), int(r[fi]), int(ldk[fi])
        blk = p[poff[fi]:poff[fi] + (kk + rr) * ld].reshape(kk + rr, ld)[:, :kk]
        ref_blk = pr[poff[fi]:poff[fi] + (kk + rr) * ld].reshape(kk + rr, ld)[:, :kk]
        z11, z11r = np.tril(blk[:kk]), np.tril(ref_blk[:kk])
        d11 = np.abs(z11 - z11r)
        s11 = np.abs(z11r).max()
        rec = None
        if d11.max() > 1e-10 * s11:
            rows, cols = np.nonzero(d11 > 1e-11 * s11)
            rec = dict(front=fi, level=int(lvl[fi]), k=kk, r=rr, z11_rel=float(d11.max() / s11), z11_n=int(len(rows)),
                       z11_rows=[int(rows.min()), int(rows.max())], z11_cols=[int(cols.min()), int(cols.max())])
        if rr:
            d21 = np.abs(blk[kk:] - ref_blk[kk:])
            s21 = np.abs(ref_blk[kk:]).max()
            if s21 > 0 and d21.max() > 1e-10 * s21:
                rows, cols = np.nonzero(d21 > 1e-11 * s21)
                rec = rec or dict(front=fi, level=int(lvl[fi]), k=kk, r=rr)
                rec.update(z21_rel=float(d21.max() / s21), z21_n=int(len(rows)), z21_rows=[int(rows.min()), int(rows.max())],
                           z21_cols=[int(cols.min()), int(cols.max())])
        if rec:
            out.append(rec)
    return out


ref = None
events = []
t0 = time.time()
for rep in range(reps):
    adj.reset_estimates()
    for it in range(10):
        res = adj.iterate(normals=True)
        if abs(res.max_corr) <= thr:
            break
    adj.form_inverse()
    q = adj.station_vcvs()
    if ref is None:
        ref = dict(q=q, w=buf(6), p=buf(1))
        continue
    scale = np.abs(ref["q"]).max()
    dq = np.abs(q - ref["q"]).reshape(len(stn), -1).max(axis=1)
    if dq.max() > 1e-10 * scale:
        off = np.nonzero(dq > 1e-11 * scale)[0]
        w, p = buf(6), buf(1)
        ev = dict(tag=tag, rep=rep, dq_rel=float(dq.max() / scale), stations=off[:30].tolist())
        dw = np.abs(w - ref["w"])
        ev["n_bad_w"] = int((dw > 1e-9 * np.maximum(np.abs(ref["w"]), 1e-3 * np.abs(ref["w"]).max())).sum())
        fr_bad = front_report(p, ref["p"])
        ev["fronts"] = fr_bad[:12]
        ev["n_fronts"] = len(fr_bad)
        if fr_bad and len(events) < 3:
            fi = fr_bad[0]["front"]
            n = (int(k[fi]) + int(r[fi])) * int(ldk[fi])
            np.savez(f"gpurun_out/locate_{tag}_rep{rep}_front{fi}.npz", bad=p[poff[fi]:poff[fi] + n], ref=ref["p"][poff[fi]:poff[fi] + n],
                     k=int(k[fi]), r=int(r[fi]), ldk=int(ldk[fi]))
        print("MISMATCH", json.dumps(ev), flush=True)
        events.append(ev)
print(f"{tag}: {reps} repetitions in {time.time() - t0:.1f} s, {len(events)} mismatches", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(events, open(f"gpurun_out/locate_{name}_{tag}.json", "w"), indent=1)
