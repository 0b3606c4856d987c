"""Worker of the world_size-2 CPU test: one rank of a sharded adjustment over gloo (hostsim kernels)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dynadjust_b200 import multigpu, synth  # noqa: E402


def main():
    lib, n, m, seed, leaf, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    stn, msr, _, _ = synth.gnss_network(n, m, seed)
    adj = multigpu.ShardedAdjustment(stn, msr, rank, world, lib_path=lib, leaf_stations=leaf)
    info = adj.prepare()
    # wipe the device copy of the records, then restore it by the sharded upload (each rank copies 1/world of the list,
    # the device copies are all-gathered): the adjustment below is only right if every rank got the whole list back
    import torch
    adj._buffer(multigpu.BUF_MSR, torch.uint8).zero_()
    adj.upload_measurements()
    last = adj.adjust()
    st = adj.statistics(write_back=True)
    est = adj.estimates()
    q = adj.station_vcvs()
    rec = msr.reshape(-1, 3)
    blocks = np.stack([adj.vcv_block(int(rec["station1"][b, 0]), int(rec["station2"][b, 0])) for b in range(0, len(rec), 7)])
    if rank == 0:
        np.savez(out, est=est, q=q, blocks=blocks, sigma0=st.sigma_zero, chi2=st.chi_squared, dof=st.dof, iters=last.iteration,
                 outliers=st.outliers, top=info.top_fronts, cut=info.cut_level, fronts=info.nfronts,
                 share=info.rank_factor_flops / info.factor_flops, measCorr=msr["measCorr"], nstat=msr["NStat"])
    adj.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
