/*
 * gadj.h — C-ABI of the B200-native least-squares adjustment engine (libgadj.so).
 *
 * This is the drop-in seam for the solve path of DynAdjust's `dnaadjust`: each
 * entry point replaces a member of the reference's `dna_adjust` class
 * (dynadjust/dynadjust/dnaadjust/dnaadjust.hpp:212-1362) that the reference
 * wrapper drives (dnaadjustwrapper.cpp:1142, dnaadjustprogress.cpp:49-67).
 * Plain pointers and sizes only; the caller owns every host array, the context
 * owns all device memory.  All functions return 0 on success, non-zero on error
 * (text from gadj_last_error); singular normals report the reference's message
 * "Matrix inversion failed, the matrix is singular." (dnamatrix_contiguous.cpp:983).
 * One host thread drives a context; streams and kernels are internal.
 *
 *   reference member (file:line)                               entry point here
 *   ---------------------------------------------------------  ---------------------------
 *   dna_adjust ctor + InitialiseAdjustment   ADJ:198-246        gadj_create
 *   LoadNetworkFiles -> bstBinaryRecords_    ADJ:10107          gadj_set_stations
 *   LoadNetworkFiles -> bmsBinaryRecords_    ADJ:10107          gadj_set_measurements
 *   LoadSegmentationFile (v_ISL_/v_JSL_)     ADJ:10644-10664    gadj_set_blocks
 *   PrepareAdjustment                        ADJ:258            gadj_prepare
 *   Solve + estimates update (one iteration) ADJ:6586, 2457-63  gadj_iterate
 *   AdjustNetwork / AdjustSimultaneous       ADJ:2140, 2413     gadj_adjust
 *   GenerateStatistics / ComputeStatistics   ADJ:6802, 7116     gadj_statistics
 *   v_estimatedStations_                     ADJH:1246-1270     gadj_get_estimates
 *   v_corrections_                                              gadj_get_corrections
 *   v_rigorousVariances_ (3x3 diagonal)      PRN:3917-4070      gadj_get_station_vcv(s)
 *   v_rigorousVariances_ (pattern blocks)    ADJ:7784-8060      gadj_get_vcv_block
 *   v_rigorousVariances_.at(block) (dense)   ADJ:3805, PRN:2942   gadj_get_block_vcv
 *   v_normals_ / At V^-1 l (before Solve)    ADJ:893-896        gadj_get_normals_block / gadj_get_rhs
 *   bms_meta_.reduced / InitialiseMeasurement ADJ:296, 3913-3935 gadj_set_measurements_reduced
 *   v_precAdjMsrsFull_ (bulk pattern blocks) ADJ:7784-8060, 6770 gadj_get_pair_vcvs
 *   UpdateIgnoredMeasurements                ADJ:8750-9980      gadj_update_ignored_measurements
 *   PrintCompMeasurements ("a-priori")       PRN:1938-2023      gadj_compute_measurements
 */
#ifndef GADJ_H_
#define GADJ_H_

#include <stdint.h>

#include "dna_records.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gadj_ctx gadj_ctx;

enum {
    GADJ_ORDER_AUTO = 0,     /* blocks from gadj_set_blocks if given, else nested dissection */
    GADJ_ORDER_DENSE = 1     /* one dense front (reference simultaneous mode on dense normals) */
};

enum {
    GADJ_ITER_NORMALS = 1,   /* re-assemble N and refactorise (always done on the first call) */
    GADJ_ITER_INVERSE = 2    /* also form the rigorous (selected) inverse in this call */
};

typedef struct gadj_opts {
    double fixed_std_dev;          /* sigma of 'C' components, default 1e-6 m   (dnaoptions.hpp:432) */
    double free_std_dev;           /* sigma of 'F' components, default 10 m */
    double iteration_threshold;    /* default (double)(float)0.0005 m           (dnaoptions.hpp:432, ADJ:2477) */
    double semi_major;             /* ellipsoid, default GRS80 */
    double inv_flattening;
    double confidence_interval;    /* default 95.0 */
    double workspace_gb;           /* selected-inverse workspace budget; 0 = automatic */
    uint32_t max_iterations;       /* default 10 */
    int32_t scale_normals_to_unity;/* default 1: symmetric diagonal equilibration (ADJ:6614-6645), always safe */
    int32_t ordering;              /* GADJ_ORDER_* */
    uint32_t leaf_stations;        /* nested-dissection leaf size, default 96 */
    int32_t device;                /* CUDA device ordinal */
    int32_t gemm_tile;             /* tile shape of the FP64 tensor GEMM launches: 0 = chosen per launch from the fronts'
                                      sizes (default), 64 or 128 = that shape wherever it is allowed (tuning / tests) */
} gadj_opts;

typedef struct gadj_iter_result {
    double max_corr;               /* signed largest-magnitude Cartesian correction (ADJ:2466) */
    uint32_t max_corr_station;     /* station index it belongs to */
    uint32_t max_corr_axis;        /* 0=X 1=Y 2=Z */
    uint32_t iteration;            /* 1-based count of iterations run on this context */
    int32_t converged;             /* |max_corr| <= iteration_threshold */
    float ms_assemble, ms_factor, ms_solve, ms_inverse;   /* device time of each phase (CUDA events) */
    double max_corr_xyz[3];        /* the three Cartesian corrections of that station (OutputLargestCorrection, ADJ:7357) */
} gadj_iter_result;

typedef struct gadj_stats {
    double chi_squared;
    double sigma_zero;             /* chi^2 / dof  (the reference's sigmaZero_) */
    int64_t dof;
    uint32_t measurement_params;
    uint32_t unknown_params;
    uint32_t outliers;
    uint32_t reserved;
    double global_pelzer;
    double critical_value;
} gadj_stats;

typedef struct gadj_info {
    uint64_t nstations, nbaselines, nedges;
    uint64_t nfronts, nlevels;
    uint64_t panel_bytes, pool_bytes, device_bytes;
    double factor_flops, inverse_flops;    /* algorithmic, from the symbolic factorisation actually used */
    uint64_t launches_factor, launches_solve, launches_inverse;
    uint32_t max_front_rows, max_front_cols;
    double rank_factor_flops, rank_inverse_flops;   /* the share of this rank (multi-GPU sharding) */
    int32_t cut_level;                               /* first tree level holding a shared ("top") front; -1 single GPU */
    uint32_t top_fronts;
    /* multi-GPU, per full iteration (factor + solve + inverse) of this rank, from the launch plan: bytes read from and
     * stored into the other ranks' replicas over NVLink, and the device-side barriers */
    double nvlink_read_bytes, nvlink_write_bytes;
    uint64_t barriers_per_step;
} gadj_info;

/* per-kernel-family device time (CUDA events around every launch) accumulated since the last reset */
typedef struct gadj_profile {
    double ms_gemm, ms_diag, ms_tri, ms_gemv, ms_transpose, ms_gather, ms_zero;
    double ms_assemble;            /* init + assembly kernels */
    double ms_other;               /* scale/scatter/permute/update kernels */
    double flops_gemm;             /* algorithmic flops of the tile-GEMM launches */
    uint64_t launches;             /* kernel launches of this library (memsets excluded) */
    uint64_t gemm_launches, gemm_tiles;
} gadj_profile;

void gadj_default_opts(gadj_opts* o);
int gadj_create(const gadj_opts* o, gadj_ctx** out);
void gadj_destroy(gadj_ctx* c);
const char* gadj_last_error(const gadj_ctx* c);   /* c may be NULL: error of the last failed gadj_create */

/* borrowed: the arrays must outlive the context (or the next set_* call) */
int gadj_set_stations(gadj_ctx* c, dna_stn_t* stn, uint32_t count);
int gadj_set_measurements(gadj_ctx* c, dna_msr_t* msr, uint64_t count);
/* Tell the library that the measurement file was reduced by an earlier adjustment (the `reduced` flag of the .bms
 * metadata; isFirstTimeAdjustment_ ADJ:296, InitialiseMeasurement ADJ:3913-3935): measured values are restored from
 * preAdjMeas before the reductions, and the one-off steps (variance scaling of GNSS measurements, conversion of
 * latitude / longitude / height point clusters) are not repeated.  Call between gadj_set_measurements and gadj_prepare. */
int gadj_set_measurements_reduced(gadj_ctx* c, int reduced);
/* optional chain segmentation (.seg): block b's inner stations are isl[isl_off[b] .. isl_off[b+1]) */
int gadj_set_blocks(gadj_ctx* c, uint32_t nblocks, const uint32_t* isl_off, const uint32_t* isl);

/* symbolic analysis, device allocation, upload, first-run variance scaling (written back to msr, ADJ:4281) */
int gadj_prepare(gadj_ctx* c);
int gadj_get_info(const gadj_ctx* c, gadj_info* info);

/* re-send the host measurement records to the device (multi-GPU: every rank sends its 1/world share over its own PCIe
 * link and pulls the rest from the peers' device copies over NVLink) */
int gadj_upload_measurements(gadj_ctx* c);
/* records [first, first + count) only */
int gadj_upload_measurements_range(gadj_ctx* c, uint64_t first, uint64_t count);
int gadj_reset_estimates(gadj_ctx* c);

int gadj_iterate(gadj_ctx* c, int flags, gadj_iter_result* res);
/* rigorous (selected) inverse from the factorisation of the last iteration; the panels then hold N^-1 */
int gadj_form_inverse(gadj_ctx* c);
/* iterate to convergence with the reference's loop logic, then form the rigorous inverse */
int gadj_adjust(gadj_ctx* c, gadj_iter_result* last);
/* needs the rigorous inverse; write_back != 0 copies the statistics fields into the host msr records
 * and the adjusted geographic coordinates into the host stn records */
int gadj_statistics(gadj_ctx* c, gadj_stats* st, int write_back);
/* A-posteriori values of the measurements flagged as ignored (UpdateIgnoredMeasurements, ADJ:8750-9980;
 * PrintIgnoredAdjMeasurements PRN:1784-1923): fills preAdjMeas, measAdj (computed from the adjusted coordinates),
 * measCorr (computed - measured) and preAdjCorr of the ignored records.  Host-side reporting; call after
 * gadj_statistics(..., write_back = 1). */
int gadj_update_ignored_measurements(gadj_ctx* c);
/* The same evaluation for the measurements that take part, at the current estimates: measAdj = computed value,
 * measCorr = computed - measured ("Computed Measurements" of --output-iter-cmp-msr, PRN:1938-2023).  Host-side
 * reporting; meant for the a-priori table before the first iteration. */
int gadj_compute_measurements(gadj_ctx* c);

int gadj_get_estimates(gadj_ctx* c, double* xyz /* 3*nstn */);
int gadj_get_corrections(gadj_ctx* c, double* dxyz /* 3*nstn */);
int gadj_get_station_vcvs(gadj_ctx* c, double* q /* 9*nstn, row-major 3x3 per station */);
int gadj_get_station_vcv(gadj_ctx* c, uint32_t stn, double q[9]);
/* N^-1 block (rows of station si, columns of station sj); only pairs joined by a measurement */
int gadj_get_vcv_block(gadj_ctx* c, uint32_t si, uint32_t sj, double q[9]);
/* Dense rigorous variance matrix of one block = front of the elimination tree (block b of gadj_set_blocks, in order):
 * its inner stations followed by its junction stations, as the reference keeps per block (v_rigorousVariances_.at(b),
 * read by the SINEX / covariance printers).  *nstations receives the station count n; when `stations` (capacity cap >= n)
 * and `packed_lower` (3n(3n+1)/2 doubles) are given they receive the station indices in matrix order and the matrix in
 * the reference's packed-lower column-major layout, idx(i,j) = j*3n - j(j-1)/2 + (i-j), i >= j (MATH:363-369). */
/* Bulk form of gadj_get_vcv_block: q[9*p .. 9*p+8] = the 3x3 block of N^-1 at (si[p], sj[p]) for npairs station pairs of the
 * stored pattern (si == sj: the station block), with one device -> host copy.  Used for the precisions of adjusted
 * baselines in alternate units and the -pam.mtx file (v_precAdjMsrsFull_, ADJ:7784-8060, 6770-6799). */
int gadj_get_pair_vcvs(gadj_ctx* c, uint64_t npairs, const uint32_t* si, const uint32_t* sj, double* q);
int gadj_get_block_vcv(gadj_ctx* c, uint32_t block, uint32_t* nstations, uint32_t* stations, uint32_t cap, double* packed_lower);
/* assembled normals (constraints included) of the last iterate call that built them, and its right-hand side */
int gadj_get_normals_block(gadj_ctx* c, uint32_t si, uint32_t sj, double n[9]);
int gadj_get_rhs(gadj_ctx* c, double* w /* 3*nstn */);

/*
 * Multi-GPU: one rank per GPU (threads of one process or one process per GPU), the dissection tree cut into subtrees,
 * one set per rank.  The fronts above the cut ("top" fronts = the separators shared by several ranks' subtrees, the
 * junction stations of the reference's phased mode) are stored by every rank at the same offsets; their work is shared
 * out tile by tile, and the exchange happens inside the kernels over NVLink peer mappings:
 *   - the ranks' partial Schur sums of a top front meet in an all-reduce kernel (each rank reduces a slice, reading the
 *     peers' replicas, and stores the sum into every replica);
 *   - each top front is factorised by one owner rank (owners spread by load, level by level) and its finished panel
 *     and pivot-tile inverses are copied into every replica by a push kernel (scatter to the peers, then forward);
 *     its Schur update and its inverse panels are computed tile-share by tile-share and pushed the same way;
 *   - the ranks meet at device-side barriers (counters in peer memory), never on the host.
 * This is the sum form of the reference's junction-station carry between blocks (ADJ:998-1281, 3196-3333) and of
 * its thread pool over blocks (dnaadjust-multi.cpp:92-310).  Call sequence per rank:
 *   gadj_create, gadj_mg_init(rank, world), gadj_set_*, gadj_prepare, gadj_mg_export  -> exchange the gadj_peer_info
 *   records of all ranks (any host transport: shared memory between threads, a torch.distributed all_gather, a pipe) ->
 *   gadj_mg_connect(all)  ->  gadj_iterate / gadj_adjust / gadj_statistics / getters as on one GPU.  Every rank must make
 *   the same calls in the same order (each one contains the same barriers).
 */
enum { GADJ_IPC_HANDLE_BYTES = 64, GADJ_PEER_BUFFERS = 9 };
typedef struct gadj_peer_info {
    int32_t rank, device;
    int64_t pid;                                        /* ranks of one process use the raw pointers (peer access enabled) */
    uint64_t ptr[GADJ_PEER_BUFFERS];                    /* device address in the exporting process */
    uint64_t bytes[GADJ_PEER_BUFFERS];
    uint8_t handle[GADJ_PEER_BUFFERS][GADJ_IPC_HANDLE_BYTES];   /* cudaIpcMemHandle of each buffer (other processes) */
    uint64_t top_panel_doubles;                         /* consistency check: the replicated layout */
    uint64_t nstations;
} gadj_peer_info;
int gadj_mg_init(gadj_ctx* c, int32_t rank, int32_t world);          /* before gadj_prepare */
int gadj_mg_export(gadj_ctx* c, gadj_peer_info* out);                /* after gadj_prepare */
int gadj_mg_connect(gadj_ctx* c, const gadj_peer_info* all /* world entries, in rank order */);
int gadj_sync(gadj_ctx* c);                                           /* wait for the context's stream */
/* raw device buffers (diagnostics and tests): x, panels, station / pair variance blocks, info words, records,
 * pivot-block inverses, inverse workspace */
enum { GADJ_BUF_X = 0, GADJ_BUF_PANELS = 1, GADJ_BUF_STATION_VCV = 2, GADJ_BUF_EDGE_VCV = 3, GADJ_BUF_INFO = 4, GADJ_BUF_MSR = 5,
       GADJ_BUF_WBUF = 6, GADJ_BUF_POOL = 7 };
int gadj_mg_buffer(gadj_ctx* c, int which, void** ptr, uint64_t* count);

/* optional per-launch timing; small overhead (two event records per launch) */
int gadj_profile_enable(gadj_ctx* c, int on);
int gadj_profile_read(gadj_ctx* c, gadj_profile* out, int reset);

/* FP64 GEMM self-test / micro-benchmark of the tensor-core tile kernel: C = A * B^T (row-major host arrays).
 * returns device milliseconds per call in *ms (averaged over reps) */
int gadj_test_gemm(gadj_ctx* c, const double* A, const double* B, double* C, int M, int N, int K, int reps, float* ms);
/* The same with the kernel's epilogue / K-range flags (GemmFlags of csrc/kernels.h: 1 accumulate into C, 2 negate, 4 lower
 * triangle only, 16 / 32 / 64 triangular operands, 128 also store the transpose into Ct [N x M]) and the tile shape
 * (64 or 128).  C holds the initial values on entry.  reps = 0: no timing. */
int gadj_test_gemm_ex(gadj_ctx* c, const double* A, const double* B, double* C, double* Ct, int M, int N, int K, int flags,
                      int tile, int reps, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* GADJ_H_ */
