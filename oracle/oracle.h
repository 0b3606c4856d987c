/*
 * oracle/oracle.h — TEST INFRASTRUCTURE (the parity checker), not product code.
 *
 * CPU restatement of the DynAdjust `dnaadjust` solve path (dense, single block,
 * simultaneous mode), used only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  The product library
 * (dynadjust_b200/libgadj.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: the dominant numerical step (packed Cholesky inverse + dspmv)
 * runs through the reference's own unmodified matrix_2d class when
 * oracle/_ref/libref_matrix.so is present (built by oracle/Makefile from
 * /root/reference).  The assembly / statistics restatement is pinned by the
 * reference's unit-test vectors for matrix_2d (tests/test_matrix.cpp) and by
 * NumPy/SciPy cross-checks; no reference test pins assembled normals or
 * station VCVs numerically -> "parity unpinned by reference
 * tests beyond 3-4 decimal .adj text" for those quantities.
 */
#ifndef GADJ_ORACLE_H_
#define GADJ_ORACLE_H_

#include <stdint.h>
#include "../include/dna_records.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_opts {
    double fixed_std_dev;         /* 1e-6  (config/dnaoptions.hpp:432) */
    double free_std_dev;          /* 10.0 */
    double iteration_threshold;   /* (double)(float)0.0005 */
    double semi_major;            /* GRS80 6378137.0 */
    double inv_flattening;        /* GRS80 298.257222101 */
    double confidence_interval;   /* 95.0 */
    uint32_t max_iterations;      /* 10 */
    int32_t scale_normals_to_unity;
    int32_t use_ref;              /* 1: route Cholesky inverse / dspmv through oracle/_ref when loaded */
    int32_t threads;              /* BLAS threads for the _ref path */
} oracle_opts;

typedef struct oracle_result {
    uint32_t iterations;          /* iterations performed */
    int32_t converged;            /* |maxCorr| <= threshold at exit */
    double max_corr;              /* signed largest-magnitude correction of last iteration */
    uint32_t max_corr_row;        /* its row in the parameter vector */
    double chi_squared;
    double sigma_zero;            /* chi^2 / dof (the reference's sigmaZero_) */
    int64_t dof;
    uint32_t measurement_params;  /* measurement rows */
    uint32_t unknown_params;      /* 3S - #constrained components */
    uint32_t outliers;
    double global_pelzer;
    double critical_value;
    int32_t used_ref;             /* 1 when the _ref library did the inversions */
    double seconds_prepare;
    double seconds_solve;         /* sum over iterations of Solve() */
    double seconds_inverse;       /* part of seconds_solve spent in the Cholesky inverse */
} oracle_result;

void oracle_default_opts(oracle_opts* o);

/* try to dlopen the reference-compiled helper; returns 1 when loaded */
int oracle_load_ref(const char* path);
int oracle_ref_loaded(void);

/*
 * Dense simultaneous adjustment of every non-ignored measurement in `msr`
 * over all `nstn` stations (parameter order = station index, LDR:146-160).
 * Mutates `stn` (current lat/lon/h <- adjusted) and `msr` (variance scaling
 * write-back ADJ:4281; statistics fields ADJ:8187-8298) like the reference.
 *
 * Optional outputs (may be NULL):
 *   est_xyz      3*nstn   adjusted Cartesian coordinates
 *   normals_full n*n      column-major full symmetric N of the first iteration (constraints included)
 *   rhs          n        At V^-1 l of the first iteration
 *   first_corr   n        corrections of the first iteration
 *   vcv_full     n*n      column-major full symmetric N^-1 (rigorous variances)
 */
int oracle_adjust_simultaneous(const oracle_opts* opts,
                               dna_stn_t* stn, uint32_t nstn,
                               dna_msr_t* msr, uint64_t nmsr,
                               double* est_xyz, double* normals_full, double* rhs,
                               double* first_corr, double* vcv_full,
                               oracle_result* res);

/*
 * Phased adjustment (AdjustPhased, ADJ:2579-2670) of a network segmented into a chain of blocks: per iteration a
 * forward pass (ADJ:2756-2852) that carries the junction stations' estimates and inverted variances into the next
 * block as pseudo measurements (CarryStnEstimatesandVariancesForward, ADJ:998-1128), then a reverse pass with the
 * combination adjustment of every intermediate block (ADJ:3461-3590, 1133-1281, 3196-3333); parameter-station
 * constraints enter once per pass through the first-appearance rule (ADJ:1884-2002, seg_file.cpp:432-486).
 * Every block is solved dense with the reference's explicit inverse, like oracle_adjust_simultaneous does the network.
 *
 * Blocks: the inner stations of block b are isl[isl_off[b] .. isl_off[b+1]); measurements belong to the block that
 * holds their first-eliminated station, junction lists follow (the .seg chain rule: JSL(b) is part of block b+1).
 * Outputs (may be NULL): est_xyz 3*nstn; stn_vcv 9*nstn (each station's rigorous 3x3 block from the block it is
 * inner to); for block `want_block` (>= 0): its station count, station list (inner then junction) and dense
 * column-major variance matrix (capacity (3 n)^2, n <= nstn).
 */
int oracle_adjust_phased(const oracle_opts* opts, dna_stn_t* stn, uint32_t nstn, dna_msr_t* msr, uint64_t nmsr,
                         uint32_t nblocks, const uint32_t* isl_off, const uint32_t* isl, double* est_xyz, double* stn_vcv,
                         int32_t want_block, uint32_t* block_nstn, uint32_t* block_stations, double* block_vcv,
                         oracle_result* res);

/* geodesy restatements exposed for unit tests */
void oracle_geo_to_cart(double lat, double lon, double h, double a, double invf, double* xyz);
void oracle_cart_to_geo(double x, double y, double z, double a, double invf, double* llh);

/* SPD inverse of a dense column-major n x n matrix (lower triangle valid), in place.
 * use_ref=1 -> reference matrix_2d::cholesky_inverse (packed path).  returns 0 ok. */
int oracle_spd_inverse(double* a, uint32_t n, int use_ref);

const char* oracle_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
