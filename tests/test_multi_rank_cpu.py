"""world_size-2 run of the sharded (multi-GPU) path on CPU: gloo collectives + the hostsim kernel stand-ins.
Checks the subtree ownership, the staged execution and the four exchange points against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dynadjust_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,n,m,seed,leaf", [(2, 400, 1200, 5, 16), (3, 600, 1800, 6, 12)])
def test_sharded_matches_oracle(oracle, hostsim_path, tmp_path, world, n, m, seed, leaf):
    out = str(tmp_path / "mg.npz")
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mg_worker.py"), hostsim_path, str(n), str(m), str(seed),
           str(leaf), out]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    got = np.load(out)
    stn, msr, _, _ = synth.gnss_network(n, m, seed)
    ref = oracle.adjust_simultaneous(stn, msr, want_vcv=True)
    rr = ref["res"]
    assert int(got["top"]) >= 1 and int(got["cut"]) >= 1          # the tree really was cut and shared fronts exchanged
    assert 0.05 < float(got["share"]) < 0.95                      # and the work really was split
    assert int(got["iters"]) == rr.iterations
    assert np.abs(got["est"] - ref["est"]).max() < 1e-9
    assert abs(float(got["sigma0"]) - rr.sigma_zero) < 1e-12
    assert int(got["dof"]) == rr.dof and int(got["outliers"]) == rr.outliers
    V = ref["vcv"]
    vs = np.abs(np.diag(V)).max()
    qd = np.stack([V[3 * s:3 * s + 3, 3 * s:3 * s + 3] for s in range(n)])
    assert np.abs(got["q"] - qd).max() < 2e-8 * vs
    rec = msr.reshape(-1, 3)
    for i, b in enumerate(range(0, len(rec), 7)):
        s1, s2 = int(rec["station1"][b, 0]), int(rec["station2"][b, 0])
        assert np.abs(got["blocks"][i] - V[3 * s1:3 * s1 + 3, 3 * s2:3 * s2 + 3]).max() < 2e-8 * vs
    assert np.abs(got["measCorr"] - msr["measCorr"]).max() < 1e-9
    assert np.abs(got["nstat"] - msr["NStat"]).max() < 1e-6
