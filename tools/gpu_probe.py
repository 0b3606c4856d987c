"""Exploratory measurements on the GPU box (not the bench): FP64 ceilings, tile-GEMM rate, phase timings.
Writes one JSON document to gpurun_out/<name>.json; every section is independent and failure-tolerant."""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dynadjust_b200 import engine, synth  # noqa: E402

OUT = {}


def section(name):
    def deco(fn):
        def run(*a, **k):
            t = time.time()
            try:
                OUT[name] = fn(*a, **k)
            except Exception as e:  # noqa: BLE001
                OUT[name] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
            OUT[name + "_wall_s"] = round(time.time() - t, 3)
            print(name, json.dumps(OUT[name])[:600], flush=True)
        return run
    return deco


@section("cublas_fp64")
def cublas_fp64():
    import torch
    res = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[f"dgemm_{n}_tflops"] = 2.0 * n ** 3 / best / 1e9
        del a, b
    # sustained: back to back for ~3 s
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 80
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    res["dgemm_8192_sustained_tflops"] = reps * 2.0 * n ** 3 / e0.elapsed_time(e1) / 1e9
    return res


@section("tile_gemm")
def tile_gemm():
    res = {}
    adj = engine.Adjustment()
    rng = np.random.default_rng(0)
    for (M, N, K) in ((256, 256, 64), (2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 128), (8192, 8192, 1024)):
        A = rng.standard_normal((M, K))
        B = rng.standard_normal((N, K))
        C, ms = adj.test_gemm(A, B, reps=3)
        err = float(np.abs(C - A @ B.T).max()) if M <= 2048 else None
        res[f"{M}x{N}x{K}"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "max_err": err}
    adj.close()
    return res


def run_network(name, stn, msr, reps=3, **opts):
    res = {}
    t = time.time()
    adj = engine.Adjustment(stn, msr, **opts)
    info = adj.prepare()
    res["prepare_s"] = time.time() - t
    for f, _ in info._fields_:
        res[f] = getattr(info, f)
    its = []
    for i in range(reps):
        adj.reset_estimates()
        t = time.time()
        r = adj.iterate(normals=True, inverse=True)
        wall = (time.time() - t) * 1e3
        its.append(dict(ms_assemble=r.ms_assemble, ms_factor=r.ms_factor, ms_solve=r.ms_solve, ms_inverse=r.ms_inverse,
                        wall_ms=wall, max_corr=r.max_corr))
    res["iterations"] = its
    best = min(its, key=lambda d: d["wall_ms"])
    res["factor_tflops"] = info.factor_flops / best["ms_factor"] / 1e9 if best["ms_factor"] > 0 else None
    res["inverse_tflops"] = info.inverse_flops / best["ms_inverse"] / 1e9 if best["ms_inverse"] > 0 else None
    res["assemble_gbs"] = info.nbaselines * 944 / best["ms_assemble"] / 1e6 if best["ms_assemble"] > 0 else None
    # converge + statistics
    adj.reset_estimates()
    t = time.time()
    last = adj.adjust()
    st = adj.statistics(write_back=False)
    res["adjust_wall_s"] = time.time() - t
    res["adjust_iterations"] = last.iteration
    res["sigma_zero"] = st.sigma_zero
    res["dof"] = st.dof
    adj.close()
    return res


@section("c1")
def c1():
    stn, msr, _, _ = synth.config_network("C1")
    return run_network("C1", stn, msr, leaf_stations=16)


@section("c2_sparse")
def c2_sparse():
    stn, msr, _, _ = synth.config_network("C2")
    return run_network("C2", stn, msr)


@section("c2_dense")
def c2_dense():
    stn, msr, _, _ = synth.config_network("C2")
    return run_network("C2", stn, msr, reps=2, ordering=engine.ORDER_DENSE)


@section("c3g")
def c3g():
    stn, msr, _, _ = synth.config_network("C3g")
    return run_network("C3g", stn, msr)


@section("deg20_100k")
def deg20_100k():
    stn, msr, _, _ = synth.gnss_network(100_000, 1_000_000, 99, hub_fraction=0.02, n_hubs=40)
    return run_network("deg20_100k", stn, msr, leaf_stations=256)


@section("c4")
def c4():
    t = time.time()
    stn, msr, _, _ = synth.config_network("C4")
    gen = time.time() - t
    r = run_network("C4", stn, msr, reps=2, leaf_stations=256)
    r["generate_s"] = gen
    return r


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="probe")
    ap.add_argument("--sections", default="cublas_fp64,tile_gemm,c1,c2_sparse,c3g")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for s in a.sections.split(","):
        globals()[s]()
        with open(os.path.join(ROOT, "gpurun_out", a.name + ".json"), "w") as f:
            json.dump(OUT, f, indent=1, default=str)
