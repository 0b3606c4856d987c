#!/usr/bin/env python
"""bench.py — ms per least-squares iteration (assemble N,w -> Cholesky -> solve -> rigorous inverse).

    python bench.py --gpus 1 --steps K --warmup W            # this engine on the B200
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle/_ref) on the host cores

One "step" is one full Gauss-Newton iteration of the BASELINE.json workload (config C4: 1M stations /
10M GNSS baselines, synthetic, seeded) from the a-priori coordinates: assembly of the normals from the raw
208-byte measurement records, supernodal FP64 Cholesky, forward/backward solve, estimate update and the
selected (rigorous) inverse.  `value` times K steps with every input resident in HBM; `e2e` times the same
K steps through the C-ABI with the measurement records in pinned host memory (H2D inside the timed region)
and the adjusted coordinates + every station's 3x3 VCV read back (D2H inside the timed region).
Inputs (6.2 GB of records, ~30 GB of front panels) are far larger than the 126 MB L2, so no explicit flush
is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ms_per_lsq_iteration"
UNIT = "ms"
BYTES_PER_BASELINE = 944   # 3*208 records + 8 plan words + 48 station XYZ + 27*8 block updates + 48 rhs

WORKLOADS = {
    # name: (synth config, engine options)
    "C4": ("C4", dict(leaf_stations=64)),      # leaf sweep on the current kernels: profiles/r1_leaf_sweep_c4.json
    "C3g": ("C3g", dict(leaf_stations=96)),
    "C2": ("C2", dict(leaf_stations=96)),
    "C1": ("C1", dict(leaf_stations=16)),
}


def workload_name(key, cfg):
    return (f"{key}: {cfg['n_stations']} stations / {cfg['n_baselines']} GNSS baselines, "
            "one Gauss-Newton iteration = assemble + factorise + solve + rigorous (selected) inverse")


def ncu_traffic(*kernels):
    """DRAM bytes per launch of the named kernel(s), summed, from the committed ncu capture of this workload
    (profiles/r1_ncu_dram_traffic_c4.json: dram__bytes_read.sum + dram__bytes_write.sum, C4 on one GPU)."""
    p = os.path.join(ROOT, "profiles", "r1_ncu_dram_traffic_c4.json")
    try:
        k = json.load(open(p))["kernels"]
        return float(sum(k[name]["dram_bytes_per_launch"] for name in kernels))
    except (OSError, KeyError, ValueError):
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = str(device_index)
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", self.idx], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak_tflops():
    """cuBLAS DGEMM 8192^3, best of 5 — MEASURED_PEAKS.json carries no FP64 figure."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e9


def cpu_sample(total_stations, baselines_per_station, blocks=1000, block_stations=150, threads=None):
    """The reference's CPU path on a bounded sample of the workload.

    The reference adjusts a network of this size block by block (phased mode) on blocks produced by dnasegment,
    whose default block size is 150 stations (min_inner_stations = max_total_stations = 150,
    include/config/dnaoptions.hpp:382).  The sample is `blocks` such blocks of the same synthetic recipe, each run
    through the oracle's dense per-block path (assembly, dpotrf+dpotri inverse through the compiled reference
    matrix_2d when oracle/_ref is present, solve); the figure is extrapolated to total_stations/150 blocks,
    forward pass only — the reference's reverse and combine passes (two more dense inversions per block,
    ADJ:3512, ADJ:3556) and the junction-station carry are NOT counted, so this is a lower bound on its time."""
    from dynadjust_b200 import synth
    from oracle import pyoracle
    threads = threads or os.cpu_count() or 1
    o = pyoracle.default_opts(threads=threads, max_iterations=1)
    nets = [synth.gnss_network(block_stations, int(block_stations * baselines_per_station), 4242 + (s % 16))[:2]
            for s in range(min(blocks, 16))]
    pyoracle.adjust_simultaneous(nets[0][0].copy(), nets[0][1].copy(), opts=o)   # library / BLAS thread start-up
    t_total = 0.0
    for s in range(blocks):
        stn, msr = nets[s % len(nets)]
        stn, msr = stn.copy(), msr.copy()
        t = time.perf_counter()
        pyoracle.adjust_simultaneous(stn, msr, opts=o)
        t_total += time.perf_counter() - t
    total_blocks = total_stations / block_stations
    kind = "reference" if pyoracle.ref_loaded() else "port"
    per_block = t_total / blocks
    return dict(value=per_block * 1e3 * total_blocks, unit=UNIT, cores=threads if kind == "reference" else 1, kind=kind,
                sample=(f"{blocks} blocks of {block_stations} stations ({3 * block_stations} unknowns, dnasegment's default block "
                        f"size, same synthetic recipe) through the per-block dense path: {t_total:.1f} s of CPU work, "
                        f"{per_block * 1e3:.2f} ms/block; extrapolated to {total_blocks:.0f} blocks, forward pass only "
                        "(reverse/combine passes and junction carry not counted: lower bound)")), t_total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dynadjust_b200 import synth
    cfg = synth.CONFIGS[WORKLOADS[args.workload][0]]
    for _ in range(min(args.warmup, 1)):
        cpu_sample(cfg["n_stations"], cfg["n_baselines"] / cfg["n_stations"], blocks=20)
    vals = []
    for _ in range(args.steps):
        base, _ = cpu_sample(cfg["n_stations"], cfg["n_baselines"] / cfg["n_stations"], blocks=args.sample_blocks)
        vals.append(base["value"])
    base["value"] = float(np.mean(vals))
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=base["value"], higher_is_better=False, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.workload, cfg),
                            note="the reference's per-block dense CPU path (oracle + compiled reference matrix_2d) on a bounded sample "
                                 "of dnasegment-size blocks, extrapolated to the network; forward pass only"),
                cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_engine(args):
    import torch
    import torch.distributed as dist
    from dynadjust_b200 import engine, synth
    from dynadjust_b200.records import MSR_DTYPE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    cfg_name, eng_opts = WORKLOADS[args.workload]
    cfg = synth.CONFIGS[cfg_name]
    stn, msr, truth, _ = synth.config_network(cfg_name)
    # measurement records in pinned host memory (the e2e leg copies them H2D every step)
    pinned = torch.empty(msr.nbytes, dtype=torch.uint8, pin_memory=True)
    msr_p = pinned.numpy().view(MSR_DTYPE)
    msr_p[:] = msr
    del msr

    if world > 1:
        from dynadjust_b200 import multigpu
        runner = multigpu.ShardedAdjustment(stn, msr_p, rank, world, multigpu.TorchExchange(torch.device("cuda", local_rank)),
                                            device=local_rank, **eng_opts)
    else:
        runner = engine.Adjustment(stn, msr_p, device=local_rank, **eng_opts)
    t = time.time()
    info = runner.prepare()
    prepare_s = time.time() - t
    peak_tf = fp64_peak_tflops()
    peaks, peaks_kind = measured_peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        runner.reset_estimates()
        return runner.iterate(normals=True, inverse=True)

    for _ in range(args.warmup):
        step()
    runner.profile_enable(True)
    runner.profile_read(reset=True)
    phase = dict(assemble=0.0, factor=0.0, solve=0.0, inverse=0.0)
    with ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank) as clk:
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = step()
            phase["assemble"] += r.ms_assemble
            phase["factor"] += r.ms_factor
            phase["solve"] += r.ms_solve
            phase["inverse"] += r.ms_inverse
        barrier()
        t1 = time.perf_counter()
    prof = runner.profile_read(reset=True)
    runner.profile_enable(False)
    clocks = clk.summary()
    ms_step = (t1 - t0) * 1e3 / args.steps
    if world > 1:
        tt = torch.tensor([ms_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step = float(tt.item())

    # ---- end to end through the C-ABI with host buffers ------------------------------
    nstn = len(stn)
    h2d = msr_p.nbytes
    d2h = nstn * 24 + nstn * 72
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        runner.upload_measurements()
        step()
        est = runner.estimates()
        vcv = runner.station_vcvs()
    barrier()
    t1 = time.perf_counter()
    e2e_ms = (t1 - t0) * 1e3 / args.steps
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())

    # sanity of what was timed: the step really solved the network
    rms = float(np.sqrt(((est - truth) ** 2).mean()))
    if not (rms < 0.05 and np.isfinite(vcv).all()):
        raise SystemExit(f"bench: adjusted coordinates are off (rms vs truth {rms})")

    if rank != 0:
        return
    alg_flops = info.factor_flops + info.inverse_flops            # whole network
    rank_flops = info.rank_factor_flops + info.rank_inverse_flops   # this rank's fronts (== alg_flops on one GPU)
    gemm_ms = prof.ms_gemm / args.steps
    roofline = dict(bound="tensor", kernel="gemm_tile_kernel (TMA-fed DMMA, all panel/Schur/inverse products)",
                    achieved=rank_flops / gemm_ms / 1e9 if gemm_ms > 0 else None, peak=peak_tf, unit="TFLOP/s",
                    frac=(rank_flops / gemm_ms / 1e9 / peak_tf) if gemm_ms > 0 else None,
                    traffic=ncu_traffic("gemm_tile_kernel") if (world == 1 and args.workload == "C4") else None,
                    traffic_note="DRAM bytes per launch (avg over the iteration's launches) from profiles/r1_ncu_dram_traffic_c4.json",
                    algorithmic_flops_per_launch=(rank_flops / (prof.gemm_launches / args.steps)) if prof.gemm_launches else None,
                    peak_source="cuBLAS DGEMM 8192^3 best-of-5 measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                    algorithmic_flops_per_step=alg_flops, rank0_algorithmic_flops_per_step=rank_flops,
                    executed_gemm_flops_per_step=prof.flops_gemm / args.steps,
                    kernel_ms_per_step=gemm_ms, kernel_share_of_step=gemm_ms / ms_step,
                    launches_per_step=prof.gemm_launches / args.steps)
    asm_ms = prof.ms_assemble / args.steps
    roofline_asm = dict(bound="hbm", kernel="init_normals_kernel + assemble_g_kernel + station_sum_kernel (the assembly pass)", achieved=info.nbaselines * BYTES_PER_BASELINE / asm_ms / 1e6,
                        peak=peaks["hbm_gbs"], unit="GB/s", peak_source=peaks_kind,
                        frac=info.nbaselines * BYTES_PER_BASELINE / asm_ms / 1e6 / peaks["hbm_gbs"],
                        bytes_per_baseline=BYTES_PER_BASELINE, kernel_ms_per_step=asm_ms,
                        traffic=ncu_traffic("init_normals_kernel", "assemble_g_kernel", "station_sum_kernel") if args.workload == "C4" else None)
    line = dict(metric=METRIC, value=ms_step, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_step, higher_is_better=False, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic",
                config=dict(workload=workload_name(args.workload, cfg),
                            ordering=f"nested dissection, leaf {eng_opts.get('leaf_stations')} stations",
                            fronts=int(info.nfronts), levels=int(info.nlevels), l2_note="inputs larger than L2",
                            sharding=(f"{world} ranks, subtrees of the dissection tree; {info.top_fronts} shared top fronts "
                                      f"replicated, their tiles shared out among the ranks; all-reduce / multicast stores / barriers in "
                                      f"our kernels over NVLink peer mappings (cut at level {info.cut_level})") if world > 1 else "none",
                            panel_gb=info.panel_bytes / 1e9, prepare_s=prepare_s),
                e2e=dict(value=e2e_ms, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
                gpu_launches=int(prof.launches), clocks=clocks, roofline=roofline, roofline_assembly=roofline_asm,
                phase_ms_per_step={k: v / args.steps for k, v in phase.items()},
                kernel_ms_per_step=dict(gemm=prof.ms_gemm / args.steps, diag=prof.ms_diag / args.steps,
                                        tri=prof.ms_tri / args.steps, gemv=prof.ms_gemv / args.steps,
                                        transpose=prof.ms_transpose / args.steps, gather=prof.ms_gather / args.steps,
                                        memset=prof.ms_zero / args.steps, assemble=prof.ms_assemble / args.steps,
                                        other=prof.ms_other / args.steps),
                gflops=alg_flops / ms_step / 1e6, rms_vs_truth_m=rms)
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_sample(cfg["n_stations"], cfg["n_baselines"] / cfg["n_stations"], blocks=args.sample_blocks)
        line["cpu_baseline"] = base
    print(json.dumps(line))
    runner.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--sample-blocks", type=int, default=2000,
                    help="150-station blocks per reference step (bounded CPU sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
