"""Worker of the multi-GPU process test: one rank per GPU under torchrun (NCCL carries the handles once)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dynadjust_b200 import multigpu, synth  # noqa: E402


def main():
    cfg, leaf, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    stn, msr, _, _ = synth.config_network(cfg)
    adj = multigpu.ShardedAdjustment(stn, msr, rank, world, multigpu.TorchExchange(torch.device("cuda", local)), device=local,
                                     leaf_stations=leaf)
    adj.prepare()
    adj.upload_measurements()
    last = adj.adjust()
    st = adj.statistics(write_back=True)
    np.savez(f"{out}.rank{rank}.npz", est=adj.estimates(), q=adj.station_vcvs(), sigma0=st.sigma_zero, iters=last.iteration)
    adj.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
