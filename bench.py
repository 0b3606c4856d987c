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

--workload selects the other BASELINE.json configurations (records of those runs are kept under profiles/):
  C2  10k stations, dense normals (one 30 000 x 30 000 front: Solve ADJ:6586 on dense normals)
  C3  100k stations, GNSS + direction sets + distances + levelling, phased: a chain of 1000-station blocks
  C5  100k stations, rigorous full inverse: every block's dense variance matrix and every pattern block read back
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
    # name: synth config, engine options, chain block width (phased), extras
    "C4": dict(cfg="C4", opts=dict(leaf_stations=64)),      # leaf sweep: profiles/r1_leaf_sweep_c4.json
    "C2": dict(cfg="C2", opts=dict(ordering=1), dense=True),
    "C3": dict(cfg="C3", opts=dict(), chain=1000),
    "C5": dict(cfg="C5", opts=dict(leaf_stations=96), full_vcv=True),
    "C3g": dict(cfg="C3g", opts=dict(leaf_stations=96)),
    "C1": dict(cfg="C1", opts=dict(leaf_stations=16)),
}


def workload_name(key, cfg):
    extra = {"C2": ", dense normals (one front)", "C3": " + direction sets, distances, levelling; phased over a chain of 1000-station blocks",
             "C5": ", rigorous full inverse: every block's dense variance matrix and every pattern block read back"}.get(key, "")
    return (f"{key}: {cfg['n_stations']} stations / {cfg['n_baselines']} GNSS baselines{extra}, "
            "one Gauss-Newton iteration = assemble + factorise + solve + rigorous (selected) inverse")


def ncu_traffic(*kernels):
    """DRAM bytes per launch of the named kernel(s), summed, from the committed ncu capture of this workload
    (dram__bytes_read.sum + dram__bytes_write.sum, C4 on one GPU)."""
    for name in ("r2b_ncu_dram_traffic_c4.json", "r2_ncu_dram_traffic_c4.json", "r1_ncu_dram_traffic_c4.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            k = json.load(open(p))["kernels"]
            return float(sum(k[kn]["dram_bytes_per_launch"] for kn in kernels)), name
        except (OSError, KeyError, ValueError):
            continue
    return None, None


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
    """cuBLAS DGEMM 8192^3: best of 5 (burst) and back to back for ~2 s (sustained) — MEASURED_PEAKS.json carries no FP64
    figure.  The GEMM kernel is timed inside a step that runs for seconds, so the roofline uses the
    sustained rate."""
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
    reps = max(8, int(2000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / reps
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / best / 1e9, 2.0 * n ** 3 / sustained / 1e9


def cached_cpu_run(key):
    """The full CPU run of this configuration made once with tools/cpu_baselines.py (C2, C3), if its record is committed."""
    p = os.path.join(ROOT, "profiles", f"r2_cpu_baseline_{key.lower()}.json")
    try:
        return json.load(open(p))
    except (OSError, ValueError):
        return None


def first_stations_subnetwork(stn, msr, nkeep):
    """The measurements among the first `nkeep` stations (whole measurements only: baselines, clusters and direction sets
    are kept or dropped as a unit) — a slice of the network with the full network's width, for bounded CPU samples."""
    keep = np.zeros(len(msr), dtype=bool)
    i, n = 0, len(msr)
    mt, st1, st2, st3 = msr["measType"], msr["station1"], msr["station2"], msr["station3"]
    vc1, vc2 = msr["vectorCount1"], msr["vectorCount2"]
    while i < n:
        t = mt[i]
        if t == b"G":
            span = 3
            ok = st1[i] < nkeep and st2[i] < nkeep
        elif t in (b"X", b"Y"):
            j = i
            ok = True
            for _ in range(int(vc1[i])):
                ok = ok and st1[j] < nkeep and (t == b"Y" or st2[j] < nkeep)
                j += 3 + 3 * int(vc2[j])
            span = j - i
        elif t == b"D":
            span = max(1, int(vc1[i]))
            ok = st1[i] < nkeep and bool((st2[i:i + span] < nkeep).all())
        else:
            span = 1
            ok = st1[i] < nkeep and (t in (b"H", b"R", b"I", b"J", b"P", b"Q") or st2[i] < nkeep) and (t != b"A" or st3[i] < nkeep)
        if ok:
            keep[i:i + span] = True
        i += span
    sub = msr[keep].copy()
    return stn[:nkeep].copy(), sub


def cpu_sample(key, args, threads=None):
    """The reference's CPU path on a bounded sample of the workload (about 10-30 s of CPU work), scaled to the workload.

    C4 / C5 / C3g: the reference adjusts a network of this size block by block (phased mode) on blocks produced by
      dnasegment, whose default block size is 150 stations (dnaoptions.hpp:382).  The sample is a number of such blocks of
      the same synthetic recipe, each run through the oracle's dense per-block path (assembly, packed Cholesky inverse
      through the compiled reference matrix_2d when oracle/_ref is present, solve), extrapolated to stations/150 blocks,
      forward pass only — reverse / combination passes and the junction carry are NOT counted: a lower bound.
    C2: the dense simultaneous path on a 2 500-station network of the same recipe (n = 7 500), its inverse + solve time
      scaled by (30 000 / 7 500)^3, its assembly by the measurement count.
    C3: the phased path (forward, reverse and combination passes) on the first blocks of the C3 network itself
      (same width, same junction sizes), time per block x 100 blocks.
    The full runs of C2 and C3 made once with tools/cpu_baselines.py are quoted beside the sample (`full_run`)."""
    from dynadjust_b200 import synth
    from oracle import pyoracle
    threads = threads or os.cpu_count() or 1
    kind = "reference" if pyoracle.ref_loaded() else "port"
    cores = threads if kind == "reference" else 1
    cfg = synth.CONFIGS[WORKLOADS[key]["cfg"]]
    full = cached_cpu_run(key)
    if key == "C2":
        n_s = 2500
        stn, msr, _, _ = synth.gnss_network(n_s, 3 * n_s, 4242)
        o = pyoracle.default_opts(threads=threads, max_iterations=1)
        t = time.perf_counter()
        r = pyoracle.adjust_simultaneous(stn, msr, opts=o)["res"]
        wall = time.perf_counter() - t
        scale3 = (cfg["n_stations"] / n_s) ** 3
        value = 1e3 * (r.seconds_solve * scale3 + r.seconds_prepare * cfg["n_baselines"] / (3 * n_s))
        sample = (f"dense simultaneous path on {n_s} stations (n = {3 * n_s}): {wall:.1f} s of CPU work, inverse + solve "
                  f"{r.seconds_solve:.2f} s scaled by (30000/{3 * n_s})^3")
    elif key == "C3":
        nblocks = args.sample_blocks_c3
        stn, msr, _, _ = synth.config_network("C3")
        sub_stn, sub_msr = first_stations_subnetwork(stn, msr, 1000 * nblocks)
        from tests import parity
        o = pyoracle.default_opts(threads=threads, max_iterations=1)
        t = time.perf_counter()
        r = pyoracle.adjust_phased(sub_stn, sub_msr, parity.chain_blocks(len(sub_stn), 1000), opts=o)["res"]
        wall = time.perf_counter() - t
        value = 1e3 * r.seconds_solve / nblocks * (cfg["n_stations"] / 1000)
        sample = (f"phased path (forward + reverse + combination, one iteration) on the first {nblocks} blocks of the C3 network: "
                  f"{wall:.1f} s of CPU work, {r.seconds_solve / nblocks:.2f} s per block x {cfg['n_stations'] // 1000} blocks "
                  "(a lower bound: the last block of the sample carries no junction stations and the two end blocks need no "
                  "combination pass; the full run is quoted beside it)")
    else:
        blocks, block_stations = args.sample_blocks, 150
        bps = cfg["n_baselines"] / cfg["n_stations"]
        o = pyoracle.default_opts(threads=threads, max_iterations=1)
        nets = [synth.gnss_network(block_stations, int(block_stations * bps), 4242 + (s % 16))[:2] for s in range(min(blocks, 16))]
        pyoracle.adjust_simultaneous(nets[0][0].copy(), nets[0][1].copy(), opts=o)   # library / BLAS thread start-up
        t_total = 0.0
        for s in range(blocks):
            stn, msr = nets[s % len(nets)]
            stn, msr = stn.copy(), msr.copy()
            t = time.perf_counter()
            pyoracle.adjust_simultaneous(stn, msr, opts=o)
            t_total += time.perf_counter() - t
        total_blocks = cfg["n_stations"] / block_stations
        value = t_total / blocks * 1e3 * total_blocks
        sample = (f"{blocks} blocks of {block_stations} stations ({3 * block_stations} unknowns, dnasegment's default block size, same "
                  f"synthetic recipe) through the per-block dense path: {t_total:.1f} s of CPU work, {t_total / blocks * 1e3:.2f} ms/block; "
                  f"extrapolated to {total_blocks:.0f} blocks, forward pass only (reverse / combination passes and junction carry not "
                  "counted: lower bound)")
    base = dict(value=value, unit=UNIT, cores=cores, kind=kind, sample=sample)
    if full:
        base["full_run"] = dict(ms_per_iteration=full.get("ms_per_iteration"), cores=full.get("cores"), mode=full.get("mode"),
                                iterations=full.get("iterations"), wall_s=full.get("wall_s"), host_cpu=full.get("cpu"),
                                record=f"profiles/r2_cpu_baseline_{key.lower()}.json (tools/cpu_baselines.py, run once)")
    return base


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dynadjust_b200 import synth
    cfg = synth.CONFIGS[WORKLOADS[args.workload]["cfg"]]
    vals = []
    base = None
    for i in range(min(args.warmup, 1) + args.steps):
        base = cpu_sample(args.workload, args)
        if i >= min(args.warmup, 1):
            vals.append(base["value"])
    base["value"] = float(np.mean(vals))
    line = dict(metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=base["value"], higher_is_better=False, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.workload, cfg),
                            note="the reference's CPU path (oracle + compiled reference matrix_2d) on a bounded sample of the workload, "
                                 "scaled to it (see cpu_baseline.sample)"),
                cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_engine(args):
    import torch
    import torch.distributed as dist
    from dynadjust_b200 import checks, engine, synth
    from dynadjust_b200.records import MSR_DTYPE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    wl = WORKLOADS[args.workload]
    eng_opts = dict(wl["opts"])
    cfg = synth.CONFIGS[wl["cfg"]]
    stn, msr, truth, _ = synth.config_network(wl["cfg"])
    gnss_only = bool((msr["measType"] == b"G").all())
    blocks = None
    if wl.get("chain"):
        from tests import parity
        blocks = parity.chain_blocks(len(stn), wl["chain"])
    if wl.get("dense") and world > 1:
        raise SystemExit("C2 (one dense front) does not shard: run it on one GPU")
    # measurement records in pinned host memory (the e2e leg copies them H2D every step)
    pinned = torch.empty(msr.nbytes, dtype=torch.uint8, pin_memory=True)
    msr_p = pinned.numpy().view(MSR_DTYPE)
    msr_p[:] = msr
    del msr

    def make_runner(sharded):
        if sharded:
            from dynadjust_b200 import multigpu
            r = multigpu.ShardedAdjustment(stn, msr_p, rank, world, multigpu.TorchExchange(torch.device("cuda", local_rank)),
                                           device=local_rank, **eng_opts)
        else:
            r = engine.Adjustment(stn, msr_p, device=local_rank, **eng_opts)
        if blocks is not None:
            r.set_blocks(blocks)
        return r

    runner = make_runner(world > 1)
    t = time.time()
    info = runner.prepare()
    prepare_s = time.time() - t
    peak_burst, peak_sustained = fp64_peak_tflops()
    peaks, peaks_kind = measured_peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        runner.reset_estimates()
        return runner.iterate(normals=True, inverse=True)

    def reduce_max(x):
        if world > 1:
            tt = torch.tensor([x], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return x

    t = time.perf_counter()
    barrier()
    step()
    barrier()
    first_iteration_ms = (time.perf_counter() - t) * 1e3     # cold: module load, first touch of every buffer
    for _ in range(max(0, args.warmup - 1)):
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
    per_rank = None
    if world > 1:
        # every rank's device time per phase and in its GEMM / exchange kernels: where the ranks wait for each other
        mine = torch.tensor([phase["assemble"], phase["factor"], phase["solve"], phase["inverse"], prof.ms_gemm, prof.ms_other, prof.ms_diag],
                            dtype=torch.float64, device="cuda") / args.steps
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        per_rank = [dict(zip(("assemble", "factor", "solve", "inverse", "gemm_kernels", "exchange_and_other_kernels", "diag_kernels"),
                             [round(float(v), 3) for v in p.cpu()])) for p in parts]
    ms_step = reduce_max((t1 - t0) * 1e3 / args.steps)
    device_ms = reduce_max(sum(phase.values()) / args.steps)       # CUDA events around the four phases, max over ranks

    # ---- end to end through the C-ABI with host buffers ------------------------------
    nstn = len(stn)
    h2d = msr_p.nbytes // world if world > 1 else msr_p.nbytes      # every rank sends its share over its own PCIe link
    d2h = nstn * 24 + nstn * 72
    rec = msr_p.reshape(-1, 3) if gnss_only else None
    full_vcv = bool(wl.get("full_vcv")) and world == 1
    if full_vcv:
        d2h += len(rec) * 72
    barrier()
    t0 = time.perf_counter()
    block_bytes = 0
    for _ in range(args.steps):
        runner.upload_measurements()
        step()
        est = runner.estimates()
        vcv = runner.station_vcvs()
        if full_vcv:
            # C5: every pattern block (one per baseline) and every block's dense variance matrix (v_rigorousVariances_, ADJ:3805)
            pv = runner.pair_vcvs(rec["station1"][:, 0], rec["station2"][:, 0])
            block_bytes = 0
            for b in range(int(info.nfronts)):
                bs, bv = runner.block_vcv(b)
                block_bytes += bv.shape[0] * (bv.shape[0] + 1) // 2 * 8
    barrier()
    t1 = time.perf_counter()
    e2e_ms = reduce_max((t1 - t0) * 1e3 / args.steps)
    d2h += block_bytes

    # sanity of what was timed: the step really solved the network
    rms = float(np.sqrt(((est - truth) ** 2).mean()))
    if not (rms < 0.05 and np.isfinite(vcv).all()):
        raise SystemExit(f"bench: adjusted coordinates are off (rms vs truth {rms})")
    parity_rec = dict(rms_vs_truth_m=rms)
    if gnss_only and rank == 0 and not args.no_checks:
        # oracle-free identity on every station: sum_j N[s,j] Z[j,s] = I from the records and the returned variance blocks
        res, worst = checks.normal_identity_residual(runner, stn, msr_p)
        parity_rec.update(normal_identity_max=res, normal_identity_worst_station=worst,
                          normal_identity_note="max over all stations of |sum_j N_sj Z_js - I| (dynadjust_b200/checks.py)")
    if world > 1 and not args.no_checks:
        # the same step on one GPU (rank 0's), compared with what the sharded run returned
        barrier()
        if rank == 0:
            one = make_runner(False)
            one.prepare()
            one.reset_estimates()
            one.iterate(normals=True, inverse=True)
            e1, q1 = one.estimates(), one.station_vcvs()
            one.close()
            parity_rec["parity_vs_1gpu"] = dict(max_dx_m=float(np.abs(est - e1).max()),
                                                max_rel_dvcv=float(np.abs(vcv - q1).max() / np.abs(q1).max()))
        barrier()

    if rank == 0:
        alg_flops = info.factor_flops + info.inverse_flops            # whole network (SURVEY 8d count, from the symbolic factorisation)
        rank_flops = info.rank_factor_flops + info.rank_inverse_flops   # this rank's share (== alg_flops on one GPU)
        gemm_ms = prof.ms_gemm / args.steps
        executed = prof.flops_gemm / args.steps                        # useful flops of the GEMM launches actually run (planner's count)
        tr, tr_file = ncu_traffic("gemm_tile_kernel") if (world == 1 and args.workload == "C4") else (None, None)
        roofline = dict(bound="tensor", kernel="gemm_tile_kernel (TMA-fed DMMA, all panel/Schur/inverse products)",
                        achieved=executed / gemm_ms / 1e9 if gemm_ms > 0 else None, peak=peak_sustained, unit="TFLOP/s",
                        frac=(executed / gemm_ms / 1e9 / peak_sustained) if gemm_ms > 0 else None,
                        traffic=tr, traffic_note=f"DRAM bytes per launch (avg over the iteration's launches) from profiles/{tr_file}" if tr_file else None,
                        flops_note="achieved = executed useful GEMM flops of this rank (the planner's per-launch count: 2MNK, halved for "
                                   "triangular outputs / operands) / summed GEMM-kernel time (CUDA events around every launch)",
                        peak_source="cuBLAS DGEMM 8192^3 sustained (back to back ~2 s) measured in this run; MEASURED_PEAKS.json has no FP64 entry",
                        peak_burst=peak_burst, executed_gemm_flops_per_step=executed,
                        algorithmic_flops_per_step=alg_flops, rank0_algorithmic_flops_per_step=rank_flops,
                        # the whole step against the same peak: this rank's executed useful GEMM flops over the step time
                        # (everything that is not a GEMM — pivot tiles, gather, substitutions, assembly — counts as lost time);
                        # symbolic_step_frac: the symbolic count (pivot-tile flops and the full k^2 r terms included) per rank
                        whole_step_frac=executed / (ms_step * 1e9) / peak_sustained,
                        symbolic_step_frac=alg_flops / world / (ms_step * 1e9) / peak_sustained,
                        kernel_ms_per_step=gemm_ms, kernel_share_of_step=gemm_ms / ms_step,
                        launches_per_step=prof.gemm_launches / args.steps)
        if wl.get("dense"):
            n3 = float(3 * nstn) ** 3
            roofline.update(dense_n3_flops=n3, dense_n3_tflops=n3 / (ms_step * 1e9), dense_n3_frac=n3 / (ms_step * 1e9) / peak_sustained)
        asm_ms = prof.ms_assemble / args.steps
        ta, _ = ncu_traffic("init_normals_kernel", "assemble_g_kernel", "station_sum_kernel") if args.workload == "C4" else (None, None)
        roofline_asm = dict(bound="hbm", kernel="init_normals_kernel + assemble_g_kernel + station_sum_kernel (the assembly pass)",
                            achieved=info.nbaselines * BYTES_PER_BASELINE / asm_ms / 1e6, peak=peaks["hbm_gbs"], unit="GB/s",
                            peak_source=peaks_kind, frac=info.nbaselines * BYTES_PER_BASELINE / asm_ms / 1e6 / peaks["hbm_gbs"],
                            bytes_per_baseline=BYTES_PER_BASELINE, kernel_ms_per_step=asm_ms, traffic=ta)
        ordering = ("one dense front" if wl.get("dense") else f"chain of {wl['chain']}-station blocks (phased)" if wl.get("chain")
                    else f"nested dissection, leaf {eng_opts.get('leaf_stations')} stations")
        line = dict(metric=METRIC, value=ms_step, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=False, scaling="strong", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=workload_name(args.workload, cfg), ordering=ordering,
                                fronts=int(info.nfronts), levels=int(info.nlevels), l2_note="inputs larger than L2",
                                sharding=(f"{world} ranks, subtrees of the dissection tree; {info.top_fronts} top fronts replicated, their "
                                          f"tiles shared out among the ranks; all-reduce, stores into the peers' replicas and barriers in our "
                                          f"own kernels over NVLink peer mappings (cut at level {info.cut_level})") if world > 1 else "none",
                                panel_gb=info.panel_bytes / 1e9, prepare_s=prepare_s,
                                gemm_tiles="64x64 (3 CTAs/SM) or 128x128 (1 CTA/SM), chosen per launch by the planner"),
                    e2e=dict(value=e2e_ms, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
                    e2e_first_iteration_ms=prepare_s * 1e3 + first_iteration_ms,
                    device_ms_per_step=device_ms,
                    gpu_launches=int(prof.launches), clocks=clocks, roofline=roofline, roofline_assembly=roofline_asm,
                    phase_ms_per_step={k: v / args.steps for k, v in phase.items()},
                    kernel_ms_per_step=dict(gemm=prof.ms_gemm / args.steps, diag=prof.ms_diag / args.steps,
                                            tri=prof.ms_tri / args.steps, gemv=prof.ms_gemv / args.steps,
                                            transpose=prof.ms_transpose / args.steps, gather=prof.ms_gather / args.steps,
                                            memset=prof.ms_zero / args.steps, assemble=prof.ms_assemble / args.steps,
                                            other=prof.ms_other / args.steps),
                    gflops=alg_flops / ms_step / 1e6, parity=parity_rec, rms_vs_truth_m=rms)
        if world > 1:
            line["per_rank_ms"] = per_rank
            line["nvlink"] = dict(bytes_read_per_step=info.nvlink_read_bytes, bytes_written_per_step=info.nvlink_write_bytes,
                                  barriers_per_step=int(info.barriers_per_step),
                                  note="rank 0: all-reduce slices read from / written to the peers' replicas + finished blocks pushed into the "
                                       "peers' replicas, per iteration (factor + solve + inverse), from the launch plan")
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(args.workload, args)
        print(json.dumps(line))
    barrier()
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
                    help="150-station blocks per reference step (bounded CPU sample, C4 / C5)")
    ap.add_argument("--sample-blocks-c3", type=int, default=3, help="blocks of the C3 chain in the bounded phased CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-checks", action="store_true", help="skip the identity / one-GPU parity checks after the timed region")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_engine(a)
