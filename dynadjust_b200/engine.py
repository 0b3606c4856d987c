"""Python host mirror of the adjustment engine's C-ABI (include/gadj.h).

``Adjustment`` follows the reference's ``dna_adjust`` call sequence
(PrepareAdjustment -> AdjustNetwork -> GenerateStatistics; dnaadjustprogress.cpp:49-67,
dnaadjustwrapper.cpp:1142-1432).  The compute path is the CUDA library
``dynadjust_b200/libgadj.so``; there is no CPU fallback — if the library is missing or
no B200 is visible the constructor raises.
"""
import ctypes as C
import os

import numpy as np

from .records import MSR_DTYPE, STN_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgadj.so")

ORDER_AUTO, ORDER_DENSE = 0, 1
ITER_NORMALS, ITER_INVERSE = 1, 2


class GadjOpts(C.Structure):
    _fields_ = [("fixed_std_dev", C.c_double), ("free_std_dev", C.c_double), ("iteration_threshold", C.c_double),
                ("semi_major", C.c_double), ("inv_flattening", C.c_double), ("confidence_interval", C.c_double),
                ("workspace_gb", C.c_double), ("max_iterations", C.c_uint32), ("scale_normals_to_unity", C.c_int32),
                ("ordering", C.c_int32), ("leaf_stations", C.c_uint32), ("device", C.c_int32), ("gemm_tile", C.c_int32)]


class GadjIterResult(C.Structure):
    _fields_ = [("max_corr", C.c_double), ("max_corr_station", C.c_uint32), ("max_corr_axis", C.c_uint32),
                ("iteration", C.c_uint32), ("converged", C.c_int32), ("ms_assemble", C.c_float),
                ("ms_factor", C.c_float), ("ms_solve", C.c_float), ("ms_inverse", C.c_float),
                ("max_corr_xyz", C.c_double * 3)]


class GadjStats(C.Structure):
    _fields_ = [("chi_squared", C.c_double), ("sigma_zero", C.c_double), ("dof", C.c_int64),
                ("measurement_params", C.c_uint32), ("unknown_params", C.c_uint32), ("outliers", C.c_uint32),
                ("reserved", C.c_uint32), ("global_pelzer", C.c_double), ("critical_value", C.c_double)]


class GadjInfo(C.Structure):
    _fields_ = [("nstations", C.c_uint64), ("nbaselines", C.c_uint64), ("nedges", C.c_uint64),
                ("nfronts", C.c_uint64), ("nlevels", C.c_uint64), ("panel_bytes", C.c_uint64),
                ("pool_bytes", C.c_uint64), ("device_bytes", C.c_uint64), ("factor_flops", C.c_double),
                ("inverse_flops", C.c_double), ("launches_factor", C.c_uint64), ("launches_solve", C.c_uint64),
                ("launches_inverse", C.c_uint64), ("max_front_rows", C.c_uint32), ("max_front_cols", C.c_uint32),
                ("rank_factor_flops", C.c_double), ("rank_inverse_flops", C.c_double), ("cut_level", C.c_int32),
                ("top_fronts", C.c_uint32), ("nvlink_read_bytes", C.c_double), ("nvlink_write_bytes", C.c_double),
                ("barriers_per_step", C.c_uint64)]


class GadjProfile(C.Structure):
    _fields_ = [("ms_gemm", C.c_double), ("ms_diag", C.c_double), ("ms_tri", C.c_double), ("ms_gemv", C.c_double),
                ("ms_transpose", C.c_double), ("ms_gather", C.c_double), ("ms_zero", C.c_double),
                ("ms_assemble", C.c_double), ("ms_other", C.c_double), ("flops_gemm", C.c_double),
                ("launches", C.c_uint64), ("gemm_launches", C.c_uint64), ("gemm_tiles", C.c_uint64)]


class GadjPeerInfo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("device", C.c_int32), ("pid", C.c_int64), ("ptr", C.c_uint64 * 9),
                ("bytes", C.c_uint64 * 9), ("handle", (C.c_uint8 * 64) * 9), ("top_panel_doubles", C.c_uint64),
                ("nstations", C.c_uint64)]


EXPORTS = ["gadj_default_opts", "gadj_create", "gadj_destroy", "gadj_last_error", "gadj_set_stations",
           "gadj_set_measurements", "gadj_set_measurements_reduced", "gadj_set_blocks", "gadj_prepare", "gadj_get_info", "gadj_upload_measurements",
           "gadj_upload_measurements_range",
           "gadj_reset_estimates", "gadj_iterate", "gadj_form_inverse", "gadj_adjust", "gadj_statistics",
           "gadj_update_ignored_measurements", "gadj_compute_measurements", "gadj_get_estimates",
           "gadj_get_corrections", "gadj_get_station_vcvs", "gadj_get_station_vcv", "gadj_get_vcv_block",
           "gadj_get_normals_block", "gadj_get_rhs", "gadj_get_block_vcv", "gadj_get_pair_vcvs", "gadj_profile_enable", "gadj_profile_read", "gadj_test_gemm", "gadj_test_gemm_ex",
           "gadj_mg_init", "gadj_mg_export", "gadj_mg_connect", "gadj_sync", "gadj_mg_buffer"]

_libs = {}


def load_library(path=None):
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the adjustment engine has no CPU fallback)")
    L = C.CDLL(path)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.gadj_last_error.restype = C.c_char_p
    L.gadj_last_error.argtypes = [vp]
    L.gadj_default_opts.argtypes = [C.POINTER(GadjOpts)]
    L.gadj_create.argtypes = [C.POINTER(GadjOpts), C.POINTER(vp)]
    L.gadj_destroy.argtypes = [vp]
    L.gadj_set_stations.argtypes = [vp, vp, u32]
    L.gadj_set_measurements.argtypes = [vp, vp, u64]
    L.gadj_set_measurements_reduced.argtypes = [vp, i32]
    L.gadj_set_blocks.argtypes = [vp, u32, vp, vp]
    L.gadj_prepare.argtypes = [vp]
    L.gadj_get_info.argtypes = [vp, C.POINTER(GadjInfo)]
    L.gadj_upload_measurements.argtypes = [vp]
    L.gadj_upload_measurements_range.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.gadj_reset_estimates.argtypes = [vp]
    L.gadj_iterate.argtypes = [vp, i32, C.POINTER(GadjIterResult)]
    L.gadj_adjust.argtypes = [vp, C.POINTER(GadjIterResult)]
    L.gadj_form_inverse.argtypes = [vp]
    L.gadj_statistics.argtypes = [vp, C.POINTER(GadjStats), i32]
    L.gadj_update_ignored_measurements.argtypes = [vp]
    L.gadj_compute_measurements.argtypes = [vp]
    L.gadj_get_estimates.argtypes = [vp, vp]
    L.gadj_get_corrections.argtypes = [vp, vp]
    L.gadj_get_station_vcvs.argtypes = [vp, vp]
    L.gadj_get_station_vcv.argtypes = [vp, u32, vp]
    L.gadj_get_vcv_block.argtypes = [vp, u32, u32, vp]
    L.gadj_get_block_vcv.argtypes = [vp, u32, vp, vp, u32, vp]
    L.gadj_get_pair_vcvs.argtypes = [vp, C.c_uint64, vp, vp, vp]
    L.gadj_get_normals_block.argtypes = [vp, u32, u32, vp]
    L.gadj_get_rhs.argtypes = [vp, vp]
    L.gadj_profile_enable.argtypes = [vp, i32]
    L.gadj_profile_read.argtypes = [vp, C.POINTER(GadjProfile), i32]
    L.gadj_mg_init.argtypes = [vp, C.c_int32, C.c_int32]
    L.gadj_mg_export.argtypes = [vp, C.POINTER(GadjPeerInfo)]
    L.gadj_mg_connect.argtypes = [vp, C.POINTER(GadjPeerInfo)]
    L.gadj_sync.argtypes = [vp]
    L.gadj_mg_buffer.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(u64)]
    L.gadj_test_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, C.POINTER(C.c_float)]
    L.gadj_test_gemm_ex.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_float)]
    _libs[path] = L
    return L


class AdjustmentError(RuntimeError):
    pass


class Adjustment:
    """One adjustment context.  ``stn`` / ``msr`` are NumPy record arrays (STN_DTYPE / MSR_DTYPE);
    they are borrowed by the library and mutated like the reference mutates its in-memory records."""

    def __init__(self, stn=None, msr=None, lib_path=None, **opts):
        self.L = load_library(lib_path)
        o = GadjOpts()
        self.L.gadj_default_opts(C.byref(o))
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown option {k}")
            setattr(o, k, v)
        self.opts = o
        h = C.c_void_p()
        if self.L.gadj_create(C.byref(o), C.byref(h)) != 0:
            raise AdjustmentError(self.L.gadj_last_error(None).decode())
        self.h = h
        self.stn = self.msr = None
        self._blocks = None
        if stn is not None:
            self.set_stations(stn)
        if msr is not None:
            self.set_measurements(msr)

    def close(self):
        if getattr(self, "h", None):
            self.L.gadj_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise AdjustmentError(self.L.gadj_last_error(self.h).decode())

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def set_stations(self, stn):
        assert stn.dtype == STN_DTYPE and stn.flags.c_contiguous
        self.stn = stn
        self._check(self.L.gadj_set_stations(self.h, self._p(stn), len(stn)))

    def set_measurements(self, msr, reduced=False):
        """reduced: the records come from a measurement file an earlier adjustment has reduced (.bms metadata flag)."""
        assert msr.dtype == MSR_DTYPE and msr.flags.c_contiguous
        self.msr = msr
        self._check(self.L.gadj_set_measurements(self.h, self._p(msr), len(msr)))
        self._check(self.L.gadj_set_measurements_reduced(self.h, 1 if reduced else 0))

    def set_blocks(self, inner_station_lists):
        """Chain segmentation as in a .seg file: one list of inner-station indices per block."""
        off = np.zeros(len(inner_station_lists) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(b) for b in inner_station_lists])
        isl = np.concatenate([np.asarray(b, dtype=np.uint32) for b in inner_station_lists]) if len(
            inner_station_lists) else np.zeros(0, np.uint32)
        self._blocks = (off, isl)
        self._check(self.L.gadj_set_blocks(self.h, len(inner_station_lists), self._p(off), self._p(isl)))

    # --- dna_adjust::PrepareAdjustment (ADJ:258)
    def prepare(self):
        self._check(self.L.gadj_prepare(self.h))
        return self.info()

    def info(self):
        i = GadjInfo()
        self._check(self.L.gadj_get_info(self.h, C.byref(i)))
        return i

    def upload_measurements(self):
        self._check(self.L.gadj_upload_measurements(self.h))

    def reset_estimates(self):
        self._check(self.L.gadj_reset_estimates(self.h))

    # --- one pass of the AdjustSimultaneous loop body (ADJ:2457-2466)
    def iterate(self, normals=True, inverse=False):
        r = GadjIterResult()
        flags = (ITER_NORMALS if normals else 0) | (ITER_INVERSE if inverse else 0)
        self._check(self.L.gadj_iterate(self.h, flags, C.byref(r)))
        return r

    def form_inverse(self):
        self._check(self.L.gadj_form_inverse(self.h))

    # --- dna_adjust::AdjustNetwork (ADJ:2140)
    def adjust(self):
        r = GadjIterResult()
        self._check(self.L.gadj_adjust(self.h, C.byref(r)))
        return r

    # --- dna_adjust::GenerateStatistics (ADJ:6802)
    def statistics(self, write_back=True):
        s = GadjStats()
        self._check(self.L.gadj_statistics(self.h, C.byref(s), 1 if write_back else 0))
        return s

    def update_ignored_measurements(self):
        """A-posteriori computed value / difference of the measurements flagged as ignored (ADJ:8750-9980)."""
        self._check(self.L.gadj_update_ignored_measurements(self.h))

    def estimates(self):
        out = np.zeros((len(self.stn), 3))
        self._check(self.L.gadj_get_estimates(self.h, self._p(out)))
        return out

    def corrections(self):
        out = np.zeros((len(self.stn), 3))
        self._check(self.L.gadj_get_corrections(self.h, self._p(out)))
        return out

    def station_vcvs(self):
        out = np.zeros((len(self.stn), 3, 3))
        self._check(self.L.gadj_get_station_vcvs(self.h, self._p(out)))
        return out

    def vcv_block(self, si, sj):
        out = np.zeros((3, 3))
        self._check(self.L.gadj_get_vcv_block(self.h, si, sj, self._p(out)))
        return out

    def pair_vcvs(self, si, sj):
        """Bulk vcv_block: (npairs, 3, 3) blocks of N^-1 at the station pairs (si[p], sj[p]) of the stored pattern."""
        si, sj = np.ascontiguousarray(si, np.uint32), np.ascontiguousarray(sj, np.uint32)
        out = np.zeros((len(si), 3, 3))
        self._check(self.L.gadj_get_pair_vcvs(self.h, len(si), self._p(si), self._p(sj), self._p(out)))
        return out

    def normals_block(self, si, sj):
        out = np.zeros((3, 3))
        self._check(self.L.gadj_get_normals_block(self.h, si, sj, self._p(out)))
        return out

    def rhs(self):
        out = np.zeros((len(self.stn), 3))
        self._check(self.L.gadj_get_rhs(self.h, self._p(out)))
        return out

    def profile_enable(self, on=True):
        self._check(self.L.gadj_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, reset=True):
        p = GadjProfile()
        self._check(self.L.gadj_profile_read(self.h, C.byref(p), 1 if reset else 0))
        return p

    def block_vcv(self, block):
        """(station indices, dense 3n x 3n variance matrix) of one block / front: inner stations, then junction stations."""
        n = C.c_uint32()
        self._check(self.L.gadj_get_block_vcv(self.h, block, C.byref(n), None, 0, None))
        stations = np.zeros(n.value, np.uint32)
        dim = 3 * n.value
        packed = np.zeros(dim * (dim + 1) // 2)
        self._check(self.L.gadj_get_block_vcv(self.h, block, C.byref(n), self._p(stations), n.value, self._p(packed)))
        # packed lower, column-major = the upper triangle walked row by row with the roles of row and column swapped
        cols, rows = np.triu_indices(dim)
        V = np.zeros((dim, dim))
        V[rows, cols] = packed
        V[cols, rows] = packed
        return stations, V

    def test_gemm(self, A, B, reps=1):
        A = np.ascontiguousarray(A, dtype=np.float64)
        B = np.ascontiguousarray(B, dtype=np.float64)
        M, K = A.shape
        N = B.shape[0]
        Cm = np.zeros((M, N))
        ms = C.c_float()
        self._check(self.L.gadj_test_gemm(self.h, self._p(A), self._p(B), self._p(Cm), M, N, K, reps, C.byref(ms)))
        return Cm, ms.value

    def test_gemm_ex(self, A, B, C0=None, flags=0, tile=128, reps=0):
        """The tile kernel with its epilogue / K-range flags and either tile shape: returns (C, Ct or None, ms)."""
        A = np.ascontiguousarray(A, dtype=np.float64)
        B = np.ascontiguousarray(B, dtype=np.float64)
        M, K = A.shape
        N = B.shape[0]
        Cm = np.zeros((M, N)) if C0 is None else np.array(C0, dtype=np.float64, order="C")
        Ct = np.zeros((N, M)) if flags & 128 else None
        ms = C.c_float()
        self._check(self.L.gadj_test_gemm_ex(self.h, self._p(A), self._p(B), self._p(Cm), self._p(Ct) if Ct is not None else None,
                                             M, N, K, flags, tile, reps, C.byref(ms)))
        return Cm, Ct, ms.value
