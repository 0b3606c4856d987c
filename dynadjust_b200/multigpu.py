"""Multi-GPU host side: one rank per GPU, the dissection tree sharded by subtree.

All of the exchange between the ranks happens on the devices (csrc/plan.cpp, kernels_gemm.cu, kernels_front.cu): the
fronts above the subtree cut are replicated; each is factorised by an owner rank, its Schur update and inverse tiles
shared out among the ranks, finished blocks copied into every replica over NVLink peer mappings by a push kernel,
partial Schur sums joined by an all-reduce kernel, and the ranks meet at device-side barriers.  The host only has to pass the ranks' buffer handles around once after
``prepare`` — that is all this module does, with two transports:

* ``TorchExchange``: one process per GPU under torchrun (``torch.distributed`` all-gather of the handle records;
  NCCL on the GPUs, gloo in the CPU tests) — what ``bench.py --gpus N`` uses;
* ``ThreadExchange``: the ranks are threads of one process (what the C++ command line ``dnaadjust --gpus N`` does
  natively) — used by the tests.

This is the sum form of the reference's junction-station carry between phased blocks
(CarryStnEstimatesandVariances{Forward,Reverse,Combine}, dnaadjust.cpp:998-1281, 3196-3333) and of its thread pool
over blocks (dnaadjust-multi.cpp:92-310), with the junction stations = the stations of the replicated fronts.
"""
import ctypes as C
import threading

from . import engine

BUF_X, BUF_PANELS, BUF_STATION_VCV, BUF_EDGE_VCV, BUF_INFO, BUF_MSR, BUF_WBUF, BUF_POOL = range(8)


class _CudaView:
    """Zero-copy torch view of library-owned device memory (diagnostics)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


class TorchExchange:
    """All-gather of the ranks' handle records over an initialised ``torch.distributed`` process group."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device

    def __call__(self, blob):
        import torch
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
        if self.device is not None:
            mine = mine.to(self.device)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine)
        return [bytes(p.cpu().numpy().tobytes()) for p in parts]


class ThreadExchange:
    """The ranks are threads of this process: a shared table and a barrier."""

    def __init__(self, world):
        self.world = world
        self.table = [None] * world
        self.barrier = threading.Barrier(world)

    def for_rank(self, rank):
        def exchange(blob):
            self.table[rank] = blob
            self.barrier.wait()
            out = list(self.table)
            self.barrier.wait()
            return out
        return exchange


class ShardedAdjustment(engine.Adjustment):
    """One rank of a multi-GPU adjustment.  After ``prepare`` it behaves like ``Adjustment``: every rank makes the same
    calls (``iterate`` / ``adjust`` / ``statistics`` / getters) and gets the same answers."""

    def __init__(self, stn, msr, rank, world, exchange, lib_path=None, **opts):
        super().__init__(stn, msr, lib_path=lib_path, **opts)
        self.rank, self.world = rank, world
        self._exchange = exchange
        self._check(self.L.gadj_mg_init(self.h, rank, world))

    def prepare(self):
        info = super().prepare()
        mine = engine.GadjPeerInfo()
        self._check(self.L.gadj_mg_export(self.h, C.byref(mine)))
        blobs = self._exchange(bytes(mine))
        table = (engine.GadjPeerInfo * self.world)()
        for q, blob in enumerate(blobs):
            C.memmove(C.addressof(table[q]), blob, C.sizeof(engine.GadjPeerInfo))
        self._check(self.L.gadj_mg_connect(self.h, table))
        return info

    def sync(self):
        self._check(self.L.gadj_sync(self.h))

    def buffer(self, which, typestr="<f8"):
        """Raw device buffer as (pointer, element count) — diagnostics."""
        ptr, cnt = C.c_void_p(), C.c_uint64()
        self._check(self.L.gadj_mg_buffer(self.h, which, C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value
