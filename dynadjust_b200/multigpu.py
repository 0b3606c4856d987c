"""Multi-GPU driver: one process per GPU, the dissection tree sharded by subtree.

Every rank holds the whole measurement list and assembles N, w redundantly (a few ms); the factorisation,
solves and the selected inverse of a front run on the rank that owns it.  Fronts above the subtree cut
("top" fronts, the separators shared by several ranks' subtrees) are the only data exchanged, with
``torch.distributed`` collectives (NCCL over NVLink on GPUs; gloo in the CPU tests) issued between the stages
of ``gadj_stage_run``:

    factorise : partial Schur sums of a top front's panel   -- reduce  --> its owner
    forward   : its slice of the right-hand side             -- all-reduce
    backward  : its slice of the solution                    -- broadcast from the owner
    inverse   : its inverse panel                            -- broadcast from the owner

This is the sum form of the reference's junction-station carry between phased blocks
(CarryStnEstimatesandVariances{Forward,Reverse,Combine}, dnaadjust.cpp:998-1281, 3196-3333), with the
junction stations = the stations of the top fronts.
"""
import ctypes as C
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import engine

PH_FACTOR, PH_FORWARD, PH_BACKWARD, PH_INVERSE = 0, 1, 2, 3
BUF_X, BUF_PANELS, BUF_STATION_VCV, BUF_EDGE_VCV, BUF_INFO, BUF_MSR = 0, 1, 2, 3, 4, 5


class _CudaView:
    """Zero-copy torch view of library-owned device memory."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


class ShardedAdjustment(engine.Adjustment):
    def __init__(self, stn, msr, rank, world, lib_path=None, **opts):
        super().__init__(stn, msr, lib_path=lib_path, **opts)
        self.rank, self.world = rank, world
        self._check(self.L.gadj_mg_init(self.h, rank, world))
        self.cuda = lib_path is None or "hostsim" not in str(lib_path)
        self._bufs = {}
        self._tops = {}
        # GADJ_MG_TRACE=1: wall-clock (phase, level, kind, ms) records of every stage and exchange on this rank
        self._trace = [] if os.environ.get("GADJ_MG_TRACE") else None

    # ---- buffers of the library as torch tensors --------------------------------------------
    def _buffer(self, which, dtype=torch.float64):
        if which in self._bufs:
            return self._bufs[which]
        ptr, cnt = C.c_void_p(), C.c_uint64()
        self._check(self.L.gadj_mg_buffer(self.h, which, C.byref(ptr), C.byref(cnt)))
        n = cnt.value
        if self.cuda:
            ts = {torch.float64: "<f8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
            t = torch.as_tensor(_CudaView(ptr.value, n, ts), device=torch.device("cuda", torch.cuda.current_device()))
        else:
            ct = {torch.float64: C.c_double, torch.int32: C.c_int32, torch.uint8: C.c_uint8}[dtype]
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))
            t = torch.from_numpy(arr)
        self._bufs[which] = t
        return t

    def _top_fronts(self, level):
        if level in self._tops:
            return self._tops[level]
        cap = 64
        while True:
            n = C.c_uint32()
            po, pl, xo, xl = (np.zeros(cap, np.uint64) for _ in range(4))
            ow = np.zeros(cap, np.int32)
            self._check(self.L.gadj_mg_top_fronts(self.h, level, cap, C.byref(n), self._p(po), self._p(pl), self._p(xo),
                                                  self._p(xl), self._p(ow)))
            if n.value <= cap:
                break
            cap = n.value
        k = n.value
        out = [(int(po[i]), int(pl[i]), int(xo[i]), int(xl[i]), int(ow[i])) for i in range(k)]
        self._tops[level] = out
        return out

    def _lib_sync(self):
        self._check(self.L.gadj_sync(self.h))

    def _torch_sync(self):
        if self.cuda:
            torch.cuda.synchronize()

    # ---- Adjustment interface ---------------------------------------------------------------
    def prepare(self):
        info = super().prepare()
        self._bufs.clear()
        self._tops.clear()
        return info

    def upload_measurements(self):
        """Host -> device copy of the measurement records, sharded: every rank copies 1/world of the list over its own
        PCIe link and the device copies are all-gathered over NVLink (each rank assembles from the whole list)."""
        n = len(self.msr)
        chunk = -(-n // self.world)
        first = min(n, self.rank * chunk)
        self._check(self.L.gadj_upload_measurements_range(self.h, first, min(chunk, n - first)))
        self._lib_sync()
        buf = self._buffer(BUF_MSR, torch.uint8)
        rec = self.msr.dtype.itemsize
        whole = buf[:self.world * chunk * rec]
        mine = whole[self.rank * chunk * rec:(self.rank + 1) * chunk * rec]
        if self.cuda:
            dist.all_gather_into_tensor(whole, mine)
        else:
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine.clone())
            for r, p in enumerate(parts):
                whole[r * chunk * rec:(r + 1) * chunk * rec] = p
        self._torch_sync()

    def _run_phase(self, phase, exchange, before):
        """Run one phase; `exchange(level)` is called at every sync marker.  `before`: the marker precedes the
        level's launches (factor / forward) — purely informational, the library places the markers."""
        cur, lvl = C.c_int64(0), C.c_int32(-1)
        trace = self._trace
        while True:
            t0 = time.perf_counter() if trace is not None else 0.0
            self._check(self.L.gadj_stage_run(self.h, phase, C.byref(cur), C.byref(lvl)))
            if lvl.value < 0:
                if trace is not None:
                    self._lib_sync()
                    trace.append((phase, -1, "compute", (time.perf_counter() - t0) * 1e3))
                break
            self._lib_sync()
            t1 = time.perf_counter() if trace is not None else 0.0
            exchange(lvl.value)
            self._torch_sync()
            if trace is not None:
                t2 = time.perf_counter()
                trace.append((phase, lvl.value, "compute", (t1 - t0) * 1e3))
                trace.append((phase, lvl.value, "exchange", (t2 - t1) * 1e3))

    def _reduce_panels(self, level):
        panels = self._buffer(BUF_PANELS)
        for po, pl, _, _, owner in self._top_fronts(level):
            dist.reduce(panels[po:po + pl], dst=owner, op=dist.ReduceOp.SUM)

    def _bcast_panels(self, level):
        panels = self._buffer(BUF_PANELS)
        for po, pl, _, _, owner in self._top_fronts(level):
            dist.broadcast(panels[po:po + pl], src=owner)

    def _sum_x(self, level):
        x = self._buffer(BUF_X)
        for _, _, xo, xl, _ in self._top_fronts(level):
            dist.all_reduce(x[xo:xo + xl], op=dist.ReduceOp.SUM)

    def _bcast_x(self, level):
        x = self._buffer(BUF_X)
        for _, _, xo, xl, owner in self._top_fronts(level):
            dist.broadcast(x[xo:xo + xl], src=owner)

    def iterate(self, normals=True, inverse=False):
        flags = (engine.ITER_NORMALS if normals else 0) | (engine.ITER_INVERSE if inverse else 0)
        self._check(self.L.gadj_stage_begin(self.h, flags))
        if self.L.gadj_stage_normals_pending(self.h):
            self._run_phase(PH_FACTOR, self._reduce_panels, True)
        self._check(self.L.gadj_stage_solve_begin(self.h))
        self._run_phase(PH_FORWARD, self._sum_x, True)
        self._run_phase(PH_BACKWARD, self._bcast_x, False)
        self._check(self.L.gadj_stage_solve_end(self.h))
        self._lib_sync()
        dist.all_reduce(self._buffer(BUF_X), op=dist.ReduceOp.SUM)
        info = self._buffer(BUF_INFO, torch.int32)
        ginfo = info[:1].clone()
        dist.all_reduce(ginfo, op=dist.ReduceOp.MAX)
        self._torch_sync()
        self._check(self.L.gadj_stage_apply(self.h))
        if inverse:
            self._run_phase(PH_INVERSE, self._bcast_panels, False)
        r = engine.GadjIterResult()
        self._check(self.L.gadj_stage_end(self.h, flags, int(ginfo.item()), C.byref(r)))
        self._vcv_ready = 0
        return r

    def adjust(self):
        """AdjustSimultaneous loop (dnaadjust.cpp:2413-2511) over the sharded iteration."""
        r = None
        for i in range(self.opts.max_iterations):
            last = i + 1 >= self.opts.max_iterations
            r = self.iterate(normals=(i == 0), inverse=False)
            if abs(r.max_corr) <= self.opts.iteration_threshold or last:
                break
        # rigorous variances: refactorise at the converged estimates is not needed for GNSS-only networks (the
        # normals never change) — the factor of the first iteration is still in the panels
        self._run_phase(PH_INVERSE, self._bcast_panels, False)
        self._check(self.L.gadj_stage_mark_inverse(self.h))
        self._vcv_ready = 0
        return r

    def _gather_vcv(self, with_edges):
        need = 2 if with_edges else 1
        if getattr(self, "_vcv_ready", 0) >= need:
            return
        self._vcv_ready = need
        self._check(self.L.gadj_mg_extract_vcv(self.h))
        self._lib_sync()
        dist.all_reduce(self._buffer(BUF_STATION_VCV), op=dist.ReduceOp.SUM)
        if with_edges:
            dist.all_reduce(self._buffer(BUF_EDGE_VCV), op=dist.ReduceOp.SUM)
        self._torch_sync()

    def station_vcvs(self):
        self._gather_vcv(False)
        return super().station_vcvs()

    def vcv_block(self, si, sj):
        self._gather_vcv(True)
        return super().vcv_block(si, sj)

    def statistics(self, write_back=True):
        self._gather_vcv(True)
        return super().statistics(write_back)
