"""Multi-rank runs of the sharded (multi-GPU) path on CPU, with the hostsim kernel stand-ins: the subtree ownership, the
replicated top fronts with their tiles shared out among the ranks, the stores into the peers' replicas, the all-reduces
and the barriers — against the oracle.  Two set-ups, the same two the product has on GPUs:

* one process per rank (torchrun + gloo carries the buffer handles; the buffers are POSIX shared memory in the hostsim
  build, cudaIpc mappings on GPUs);
* the ranks as threads of one process (what `dnaadjust --gpus N` does)."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from dynadjust_b200 import multigpu, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mg_worker  # noqa: E402


def check(oracle, got_ranks, stn, msr, n, m, tol_sigma0=1e-12, tol_corr=1e-9):
    ref = oracle.adjust_simultaneous(stn, msr, want_vcv=True)
    rr = ref["res"]
    V = ref["vcv"]
    vs = np.abs(np.diag(V)).max()
    qd = np.stack([V[3 * s:3 * s + 3, 3 * s:3 * s + 3] for s in range(n)])
    rec = msr[:3 * m].reshape(-1, 3)
    shares = []
    for got in got_ranks:
        assert int(got["top"]) >= 1 and int(got["cut"]) >= 1          # the tree really was cut and fronts shared
        shares.append(float(got["share"]))
        assert int(got["iters"]) == rr.iterations
        assert np.abs(got["est"] - ref["est"]).max() < 1e-9
        assert abs(float(got["sigma0"]) - rr.sigma_zero) < tol_sigma0 * max(1.0, rr.sigma_zero)
        assert int(got["dof"]) == rr.dof and int(got["outliers"]) == rr.outliers
        assert np.abs(got["q"] - qd).max() < 2e-8 * vs
        for i, b in enumerate(range(0, len(rec), 7)):
            s1, s2 = int(rec["station1"][b, 0]), int(rec["station2"][b, 0])
            assert np.abs(got["blocks"][i] - V[3 * s1:3 * s1 + 3, 3 * s2:3 * s2 + 3]).max() < 2e-8 * vs
        assert np.abs(got["measCorr"] - msr["measCorr"]).max() < tol_corr
        assert np.abs(got["nstat"] - msr["NStat"]).max() < 1e-4
    assert all(0.02 < s < 0.98 for s in shares) and abs(sum(shares) - 1.0) < 1e-9   # and the work really was split
    # every rank ends with the same numbers (the all-reduce stores identical bits into every replica)
    for got in got_ranks[1:]:
        assert np.array_equal(got["est"], got_ranks[0]["est"])
        assert np.array_equal(got["q"], got_ranks[0]["q"])


@pytest.mark.parametrize("world,n,m,seed,leaf", [(2, 400, 1200, 5, 16), (3, 600, 1800, 6, 12)])
def test_sharded_processes_match_oracle(oracle, hostsim_path, tmp_path, world, n, m, seed, leaf):
    out = str(tmp_path / "mg")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mg_worker.py"), hostsim_path, str(n), str(m), str(seed),
           str(leaf), out]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    got = [np.load(f"{out}.rank{r}.npz") for r in range(world)]
    stn, msr, _, _ = synth.gnss_network(n, m, seed)
    check(oracle, got, stn, msr, n, m)


def run_threads(lib, world, n, m, seed, leaf, terrestrial=False, **net):
    ex = multigpu.ThreadExchange(world)
    results, errors = [None] * world, []

    def work(rank):
        try:
            results[rank] = mg_worker.run_rank(lib, rank, world, ex.for_rank(rank), n, m, seed, leaf, terrestrial, **net)
        except Exception as e:   # a failing rank must not leave the others waiting at the exchange
            errors.append(e)
            ex.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    return results


@pytest.mark.parametrize("world,n,m,seed,leaf", [(2, 400, 1200, 5, 16), (4, 900, 2700, 8, 10), (8, 1500, 4500, 9, 8)])
def test_sharded_threads_match_oracle(oracle, hostsim_path, world, n, m, seed, leaf):
    got = run_threads(hostsim_path, world, n, m, seed, leaf)
    stn, msr, _, _ = synth.gnss_network(n, m, seed)
    check(oracle, got, stn, msr, n, m)


def test_sharded_threads_terrestrial(oracle, hostsim_path):
    """Every measurement type through the sharded path (the normals move with the estimates: refactorised every iteration)."""
    from dynadjust_b200 import synth_terrestrial
    world, n, m, seed, leaf = 3, 500, 1200, 21, 12
    got = run_threads(hostsim_path, world, n, m, seed, leaf, terrestrial=True)
    stn, msr, _, _ = synth_terrestrial.terrestrial_network(n, m, seed, scalars={"S": n // 2, "L": n // 3}, n_dir_sets=n // 8)
    check(oracle, got, stn, msr, n, m, tol_sigma0=1e-7, tol_corr=5e-9)


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_threads_wide_top_fronts(oracle, hostsim_path, world):
    """Long-range "hub" baselines put ~150 stations into the top separators: replicated fronts of several ragged 128-wide
    pivot steps, so the distributed right-looking factorisation (pivot tile by the owner of its row tile, panel tiles
    stored into every replica, trailing tiles kept by their owners), the block-doubling inverse and the distributed
    selected inverse all run with more than one tile per rank."""
    n, m, seed, leaf = 2400, 7200, 31, 24
    net = dict(hub_fraction=0.08, n_hubs=120)
    got = run_threads(hostsim_path, world, n, m, seed, leaf, **net)
    stn, msr, _, _ = synth.gnss_network(n, m, seed, **net)
    check(oracle, got, stn, msr, n, m)
