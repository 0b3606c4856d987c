"""Worker of the multi-rank CPU tests: one rank of a sharded adjustment (hostsim kernels), one process per rank; the
peer handles travel over gloo, the data over shared memory — the CPU image of NCCL-launched ranks exchanging cudaIpc
handles and then talking over NVLink."""
import ctypes
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dynadjust_b200 import multigpu, synth  # noqa: E402


def run_rank(lib, rank, world, exchange, n, m, seed, leaf, terrestrial=False, **net):
    if terrestrial:
        from dynadjust_b200 import synth_terrestrial
        stn, msr, _, _ = synth_terrestrial.terrestrial_network(n, m, seed, scalars={"S": n // 2, "L": n // 3}, n_dir_sets=n // 8)
    else:
        stn, msr, _, _ = synth.gnss_network(n, m, seed, **net)
    adj = multigpu.ShardedAdjustment(stn, msr, rank, world, exchange, lib_path=lib, leaf_stations=leaf)
    info = adj.prepare()
    # wipe the device copy of the records, then restore it by the sharded upload (each rank sends 1/world of the list and
    # pulls the rest from its peers): the adjustment below is only right if every rank got the whole list back
    ptr, nbytes = adj.buffer(multigpu.BUF_MSR)
    ctypes.memset(ptr, 0, msr.nbytes)
    adj.upload_measurements()
    last = adj.adjust()
    st = adj.statistics(write_back=True)
    est = adj.estimates()
    q = adj.station_vcvs()
    rec = msr[:3 * m].reshape(-1, 3)
    blocks = np.stack([adj.vcv_block(int(rec["station1"][b, 0]), int(rec["station2"][b, 0])) for b in range(0, len(rec), 7)])
    out = dict(est=est, q=q, blocks=blocks, sigma0=st.sigma_zero, chi2=st.chi_squared, dof=st.dof, iters=last.iteration,
               outliers=st.outliers, top=info.top_fronts, cut=info.cut_level, fronts=info.nfronts,
               share=info.rank_factor_flops / info.factor_flops, measCorr=msr["measCorr"].copy(), nstat=msr["NStat"].copy())
    adj.close()
    return out


def main():
    lib, n, m, seed, leaf, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    res = run_rank(lib, rank, world, multigpu.TorchExchange(), n, m, seed, leaf)
    np.savez(f"{out}.rank{rank}.npz", **res)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
