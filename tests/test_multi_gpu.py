"""Multi-GPU runs on real devices (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`).

The ranks as threads of one process (peer access between the devices, raw pointers — what `dnaadjust --gpus N` does)
and as one process per GPU under torchrun (cudaIpc handles — what `bench.py --gpus N` does), against the oracle and
against the one-GPU engine at a size the dense oracle cannot reach."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from dynadjust_b200 import engine, multigpu, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_threads(world, stn, msr, blocks_fn=None, **opts):
    ex = multigpu.ThreadExchange(world)
    results, errors = [None] * world, []

    def work(rank):
        try:
            s, m = stn.copy(), msr.copy()
            adj = multigpu.ShardedAdjustment(s, m, rank, world, ex.for_rank(rank), device=rank, **opts)
            info = adj.prepare()
            adj.upload_measurements()
            last = adj.adjust()
            st = adj.statistics(write_back=True)
            results[rank] = dict(est=adj.estimates(), q=adj.station_vcvs(), sigma0=st.sigma_zero, chi2=st.chi_squared, dof=st.dof,
                                 outliers=st.outliers, iters=last.iteration, top=info.top_fronts, msr=m,
                                 share=info.rank_factor_flops / info.factor_flops)
            adj.close()
        except Exception as e:
            errors.append(e)
            ex.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return results


@pytest.mark.parametrize("hubs", [False, True])
def test_threads_match_oracle(oracle, gpu_lib, hubs):
    world = min(gpu_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    n, m = 2400, 7200
    net = dict(hub_fraction=0.08, n_hubs=120) if hubs else {}
    stn, msr, _, _ = synth.gnss_network(n, m, 31, **net)
    got = run_threads(world, stn, msr, leaf_stations=24)
    ref = oracle.adjust_simultaneous(stn, msr, want_vcv=True)
    V = ref["vcv"]
    vs = np.abs(np.diag(V)).max()
    qd = np.stack([V[3 * s:3 * s + 3, 3 * s:3 * s + 3] for s in range(n)])
    for g in got:
        assert g["top"] >= 1 and 0.02 < g["share"] < 0.98
        assert g["iters"] == ref["res"].iterations
        assert np.abs(g["est"] - ref["est"]).max() < 1e-9
        assert abs(g["sigma0"] - ref["res"].sigma_zero) < 1e-12
        assert g["dof"] == ref["res"].dof and g["outliers"] == ref["res"].outliers
        assert np.abs(g["q"] - qd).max() < 2e-8 * vs
        assert np.abs(g["msr"]["measCorr"] - msr["measCorr"]).max() < 1e-9
    for g in got[1:]:
        assert np.array_equal(g["est"], got[0]["est"]) and np.array_equal(g["q"], got[0]["q"])


def test_threads_match_one_gpu_at_scale(gpu_lib):
    """BASELINE config C3 (100k stations, mixed types): the sharded run reproduces the one-GPU run — estimates, sigma-zero
    and every station's variance block — and two sharded runs reproduce each other."""
    world = min(gpu_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    stn, msr, truth, _ = synth.config_network("C3")
    s1, m1 = stn.copy(), msr.copy()
    one = engine.Adjustment(s1, m1, leaf_stations=96)
    one.prepare()
    last = one.adjust()
    st1 = one.statistics(write_back=True)
    e1, q1 = one.estimates(), one.station_vcvs()
    one.close()
    got = run_threads(world, stn, msr, leaf_stations=96)
    for g in got:
        assert g["iters"] == last.iteration
        assert np.abs(g["est"] - e1).max() < 2e-9
        assert abs(g["sigma0"] - st1.sigma_zero) < 1e-10
        assert g["dof"] == st1.dof and g["outliers"] == st1.outliers
        assert np.abs(g["q"] - q1).max() < 1e-10 * np.abs(q1).max()
        assert np.abs(g["msr"]["measCorr"] - m1["measCorr"]).max() < 1e-8
    again = run_threads(world, stn, msr, leaf_stations=96)
    assert np.abs(again[0]["q"] - got[0]["q"]).max() < 1e-11 * np.abs(q1).max()


def test_processes_match_one_gpu(gpu_lib, tmp_path):
    """One process per GPU under torchrun: the buffer handles travel through torch.distributed, the data over NVLink."""
    world = min(gpu_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    out = str(tmp_path / "mg")
    port = 29600 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mg_gpu_worker.py"), "C3g", "96", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-3000:]
    got = [np.load(f"{out}.rank{r}.npz") for r in range(world)]
    stn, msr, _, _ = synth.config_network("C3g")
    one = engine.Adjustment(stn, msr, leaf_stations=96)
    one.prepare()
    last = one.adjust()
    st1 = one.statistics(write_back=True)
    e1, q1 = one.estimates(), one.station_vcvs()
    one.close()
    for g in got:
        assert int(g["iters"]) == last.iteration
        assert np.abs(g["est"] - e1).max() < 2e-9
        assert abs(float(g["sigma0"]) - st1.sigma_zero) < 1e-11
        assert np.abs(g["q"] - q1).max() < 1e-10 * np.abs(q1).max()


def test_command_line_gpus_option(oracle, gpu_lib, tmp_path):
    """`dnaadjust <net> --gpus 2` (the product binary; ranks as threads with peer access, csrc/host/gpu_group.hpp): the reports
    equal the oracle's at print resolution, simultaneous and phased, per-block SINEX from the ranks holding the blocks."""
    from dynadjust_b200 import dnafiles
    from tests import parity
    from tests.test_cli import _check_outputs, _run, _write_network
    if gpu_count() < 2:
        pytest.skip("needs at least two GPUs")
    exe = os.path.join(ROOT, "dynadjust_b200", "bin", "dnaadjust")
    assert os.path.exists(exe), "dynadjust_b200/bin/dnaadjust is missing on a GPU box: run __graft_entry__.build()"
    stn, msr, _, _ = synth.gnss_network(900, 2700, 41)
    _write_network(tmp_path, "mg", stn, msr)
    r = _run(exe, tmp_path, "mg", "--gpus", "2", "--output-adj-msr", "--output-pos-uncertainty", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    _check_outputs(oracle, tmp_path, "mg", "simult", stn, msr, True)
    blocks = parity.chain_blocks(len(stn), 150)
    isl = [list(b) for b in blocks]
    dnafiles.write_seg(os.path.join(tmp_path, "mg.seg"), isl, [[] for _ in isl], [[] for _ in isl])
    r = _run(exe, tmp_path, "mg", "--phased", "--gpus", "2", "--export-sinex-file", "--no-binary-update")
    assert r.returncode == 0, r.stderr
    _check_outputs(oracle, tmp_path, "mg", "phased", stn, msr, False)
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".snx")]) == len(blocks)
