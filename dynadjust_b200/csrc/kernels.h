// kernels.h — device-side work descriptors and the launch interface.
//
// Every numeric phase of an iteration is a list of launches; each launch
// consumes an array of small descriptors ("ops") that lives in device memory
// and is built once by the planner (plan.cpp) after symbolic analysis.
// The CUDA implementations are in kernels_*.cu.  tests/hostsim/ provides
// plain-loop stand-ins with the same signatures so that the planner, the
// index maps and the algorithm can be exercised on a machine without a GPU;
// that build is test infrastructure and is never linked into libgadj.so.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/dna_records.h"

namespace gadj {

// ---- tile geometry of the FP64 tensor-core GEMM -------------------------------
constexpr int TILE_M = 128;
constexpr int TILE_N = 128;
constexpr int TILE_K = 16;   // doubles per TMA box row (= 128 bytes, SWIZZLE_128B)
// Tile shape of a GEMM launch (all tiles of a launch have the same shape; GemmTile::tm / tn and the ops' tensor-map
// boxes are in its units): 64 x 64 wherever it costs no more padded work, 128 x 128 for launches of long products
// (K >= ~2000), shared / pushed tiles of the top fronts and wide in-place products (plan.cpp, choose_shape).
enum TileShape : int32_t { TILE_SHAPE_128 = 0, TILE_SHAPE_64 = 1 };
constexpr int tile_dim(int shape) { return shape == TILE_SHAPE_64 ? 64 : TILE_M; }
constexpr int NB = 128;      // pivot block width of the blocked factorisation

struct alignas(64) TmaDesc {
    uint64_t opaque[16];     // CUtensorMap (128 bytes)
};

// ---- multi-GPU: replicated buffers and the peer table ---------------------------------------------------------
// The fronts above the subtree cut ("top" fronts) are stored by every rank at identical offsets of its panel,
// pivot-inverse and workspace buffers.  A rank that finishes a piece of such a front (a factorised panel, tiles of the
// inverse) pushes it into every replica through NVLink peer mappings (launch_push): byte address in rank p's copy =
// local address + delta[buf][p].
constexpr int MAX_RANKS = 8;
enum McBuf : int32_t { MC_NONE = 0, MC_PANELS = 1, MC_WBUF = 2, MC_POOL = 3, MC_X = 4, MC_VCVD = 5, MC_VCVO = 6, MC_MSR = 7, MC_BUFS = 8 };
struct PeerTable {
    int32_t nranks, rank;
    int64_t delta[MC_BUFS][MAX_RANKS];             // delta[b][rank] == 0
    unsigned long long* counter[MAX_RANKS];        // every rank's barrier counter, as mapped into this rank
    int32_t* info[MAX_RANKS];                      // every rank's info words: [0] first front with a non-positive pivot + 1, [1] barrier time-out
};

// a rows x cols block (row pitch ld, doubles) at offset `off` of replicated buffer `buf`, copied into every peer's replica
struct PushOp {
    uint64_t off;
    int64_t ld;
    int32_t rows, cols;
    int32_t buf;
    int16_t target;          // >= 0: to that rank only; -1: to every peer
    int16_t skip;            // >= 0: not to that rank (it holds the block already); -1: nobody skipped
};

// sum of `count` doubles at offset `off` of a replicated buffer over all ranks, result stored into every replica
// (each rank reduces its 1/nranks slice; fixed rank order, so every replica ends up with the same bits)
struct ReduceOp {
    uint64_t off, count;
};

enum GemmFlags : int32_t {
    GEMM_ACCUM = 1,     // C += alpha*A*B^T   (else C = alpha*A*B^T)
    GEMM_NEG = 2,       // alpha = -1         (else +1)
    GEMM_LOWER = 4,     // keep only elements with (i + tri_off) >= j; tiles wholly above are skipped
    GEMM_KLO_ROW = 16,  // A is upper triangular (zero for k < row): start the K loop at the tile's first row
    GEMM_KLO_MAX = 32,  // A and B upper triangular: start the K loop at max(first row, first column) of the tile
    GEMM_KHI_ROW = 64,  // A is lower triangular (zero for k > row): end the K loop after the tile's last row
    GEMM_DUAL = 128,    // also store the transpose: Ct[j][i] = C[i][j] (row-major, pitch ldct); not with ACCUM / LOWER / SCATTER
    GEMM_SCATTER = 8,   // Schur update of a front, scattered into its ancestors' panels: column j belongs to boundary
                        // station j/3, whose owning ancestor is target t = coltgt[j/3]; element (i, j) is added
                        // atomically to tgt[t].C at row 3*tgt[t].rowmap[i/3 - tgt[t].jb] + i%3 and column
                        // 3*tgt[t].rowmap[j/3 - tgt[t].jb] + j%3 (station-level maps)
};

// One update target of a front: the ancestor that owns boundary stations [jb, je) of the front.
struct ScatterTarget {
    double* C;               // the ancestor's panel
    const int32_t* rowmap;   // entry i - jb: local station row in the ancestor of the front's boundary station i >= jb
    int64_t ldc;
    int32_t jb;
    int32_t pad;
};

// C[M x N] (row-major, ldc)  (+)= alpha * A[M x K] (row-major, lda) * B[N x K]^T (row-major, ldb)
struct alignas(64) GemmOp {
    TmaDesc tmA, tmB;
    const double* A;
    const double* B;
    double* C;
    const int32_t* coltgt;   // GEMM_SCATTER: per boundary station, index into tgt
    const ScatterTarget* tgt;
    double* Ct;              // GEMM_DUAL: transposed copy of the result
    int64_t ldct;
    int64_t lda, ldb, ldc;
    int32_t M, N, K;
    int32_t flags;
    int32_t tri_off;
    int32_t pad0;
    int32_t tiles_m, tiles_n;
};

// One output tile (128 x 128 or 64 x 64: the launch's TileShape) of one op.  The planner lists only the tiles that
// hold work (tiles wholly above the diagonal of a lower-only output are left out); a launch is a contiguous run of
// this list and the persistent GEMM CTAs stride through it.
struct GemmTile {
    int32_t op;              // index into the launch's op array
    uint16_t tm, tn;         // tile row / column inside the op
};

// One pivot tile (w <= NB).  factor: D (row-major, ldd) <- chol(D) lower, in place.
// W  (optional): inverse of the lower factor, row-major, pitch ldw, zeros above the diagonal, rows/cols >= w untouched
// Wt (optional): its transpose, row-major, pitch ldwt, zeros below the diagonal
struct DiagOp {
    double* D;
    double* W;
    double* Wt;
    int64_t ldd, ldw, ldwt;
    int32_t w;
    int32_t factor;
    int32_t front;           // for error reporting
    int32_t pad;
};

// y[row0 + i] = sum_c A[i][c] * x[c]  for one chunk of rows of a k x k triangular matrix (W = L11^-1 in the forward
// substitution, Wt = W^T in the backward one).  Only the triangle is read: c <= row (lower) or c >= row (upper) —
// the other half of the buffer holds scratch.
struct TrimvOp {
    const double* A;         // first row of the chunk (row-major, pitch ld)
    const double* x;         // the front's slice of the input vector (k entries)
    double* y;               // the front's slice of the output vector (k entries)
    int64_t ld;
    int32_t row0, nrows;     // rows [row0, row0 + nrows) of the front
    int32_t k;
    int32_t upper;           // 0: lower triangular, 1: upper triangular
    int32_t wide;            // 1: nrows <= TRIMV_WIDE_ROWS and the whole CTA strides across the columns (wide fronts)
};
constexpr int TRIMV_WIDE_ROWS = 8;     // rows per CTA of a wide front
constexpr int TRIMV_WIDE_K = 1024;     // fronts at least this wide take the wide form

// forward : x[rowidx[i]] -= sum_c P[i][c] * xj[c]         for i in [0, nrows)
// backward: xj[c]        -= sum_i P[i][c] * x[rowidx[i]]
struct GemvOp {
    const double* P;         // first row of the chunk, first column of the pivot tile
    const int32_t* rowidx;   // global unknown index of each row in the chunk
    double* xj;              // the pivot tile's slice of the solution vector
    int64_t ld;
    int32_t nrows;
    int32_t w;
};

// dst[c][r] = src[r][c]
struct TransposeOp {
    const double* src;
    double* dst;
    int64_t lds, ldd;
    int32_t rows, cols;
};

// G[(i)][(j)] = Z entry of boundary pair (i, j) read from the owning ancestor panel
// (station-level maps, 3x3 blocks): for boundary station range [jb, je) owned by
// ancestor panel `Z` (pitch ld, first own column col0) and every i >= jb:
//   block(i, j) = Z[3*rowmap[i-jb] .. +3][3*rowmap[j-jb] .. +3]; mirrored into (j, i).
struct GatherOp {
    const double* Z;
    const int32_t* rowmap;
    double* G;               // r x r row-major workspace, pitch ldg
    int64_t ld, ldg;
    int32_t jb, je, nb;      // nb = boundary station count of the front
    int32_t col0;
};

// One 16 x 16-station tile of one gather op, on or below the op's diagonal (the planner lists only those; the kernel
// mirrors them into the upper part of G).
constexpr int GATHER_TILE_STATIONS = 16;
struct GatherTile {
    int32_t op;              // index into the launch's op array
    uint16_t ti, tj;         // tile row / column, in units of 16 stations relative to the op's jb
};

// keys of dev::first_use: per-device one-time kernel attribute set-up
enum FirstUseKey : int { KEY_GEMM = 1, KEY_DIAG = 2, KEY_ASSEMBLE = 3, KEY_GEMM64 = 4, KEY_GEMM_LDG = 5, KEY_GEMM64_LDG = 6 };

// ---- launches -----------------------------------------------------------------
// All pointers are device pointers; `stream` is the backend's stream handle.
void launch_gemm(const GemmOp* ops, int nops, const GemmTile* tiles, int ntiles, int shape /* TileShape */, void* stream);
void launch_diag(const DiagOp* ops, int nops, int* info, void* stream);
// multi-GPU: blocks of the replicated buffers copied into every peer's replica (bases[buf] = this rank's buffer);
// grid_x CTAs per op
void launch_push(const PushOp* ops, int nops, int grid_x, const PeerTable* pt, double* const* bases, void* stream);
// all ranks meet: every store issued before the barrier by any rank is visible to every rank after it.  `target` =
// (number of barriers so far on this context) * nranks.  A rank that waits longer than ~20 s sets info[1] and gives up.
void launch_barrier(const PeerTable* pt, unsigned long long target, int* info, void* stream);
void launch_allreduce(const ReduceOp* ops, int nops, const PeerTable* pt, double* base, int buf, void* stream);
// info[0] <- max over ranks (non-positive pivot met by any rank); between two barriers
void launch_share_info(const PeerTable* pt, int* info, void* stream);
void launch_trimv(const TrimvOp* ops, int nops, void* stream);
void launch_gemv(const GemvOp* ops, int nops, const double* x_ro, double* x, int backward, void* stream);
// grid_x: CTAs per op (each op loops over its tiles); the planner passes min(cap, largest tile count)
void launch_transpose(const TransposeOp* ops, int nops, int grid_x, void* stream);
void launch_gather(const GatherOp* ops, int nops, const GatherTile* tiles, int ntiles, void* stream);

// ---- assembly -----------------------------------------------------------------
// Edge word of a GNSS baseline: bits 0..29 edge slot, bit 31 = station1 is eliminated after station2, bit 30 = this
// baseline is the only contribution to its off-diagonal block: the block is not stored at all, the scatter into the
// panels reads it from the baseline's slot (ScatterParams::edge_bsl).
constexpr uint32_t EDGE_SLOT_MASK = 0x3FFFFFFFu;
constexpr uint32_t EDGE_EXCLUSIVE = 0x40000000u;
// element k (row-major) of a symmetric 3x3 block -> index into its stored upper triangle {00 01 02 11 12 22}
#define GADJ_SYM3(k) ((k) == 0 ? 0 : (k) == 1 || (k) == 3 ? 1 : (k) == 2 || (k) == 6 ? 2 : (k) == 4 ? 3 : (k) == 8 ? 5 : 4)

struct AssembleParams {
    const dna_msr_t* msr;        // device copy of the raw .bms records
    const uint32_t* first;       // per GNSS baseline: index of its X record
    const uint32_t* edge;        // per baseline: edge word (above)
    const double* est;           // 3 x nstn estimated Cartesian coordinates (station order)
    double* bq;                  // 9 x nbaselines: V^-1 (upper triangle, 6) and V^-1 l (3) of every baseline
    const uint32_t* inc_ptr;     // nstn + 1: incidence lists of the stations ...
    const uint32_t* inc;         // ... entries = baseline index | bit 31 when the station is the baseline's station2
    double* ndiag;               // nstn x 9 diagonal blocks (station order, row-major 3x3), initialised by init_normals
    double* noff;                // nedge x 9 off-diagonal blocks: N[later station, earlier station]
    double* w;                   // 3 x nstn  A^T V^-1 l (station order)
    uint64_t nbaselines;
    uint32_t nstn;
    int32_t contiguous;          // first[b] == first[0] + 3b for all b
    int32_t normals;             // 0: rhs only
};
// two launches: per-baseline pass (records -> bq, off-diagonal blocks), per-station gather (bq -> ndiag, w)
void launch_assemble_g(const AssembleParams& p, void* stream);
void launch_station_sum(const AssembleParams& p, void* stream);

// design rows of every other measurement type and the D / X / Y clusters: parameter blocks in rows.h
struct RowsParams;
struct ClusterParams;
struct ClusterDesc;
void launch_rows(const RowsParams& p, void* stream);
void launch_rows_stats(const RowsParams& p, void* stream);
void launch_clusters(const ClusterParams& p, void* stream);
void launch_cluster_chi(const ClusterParams& p, void* stream);
// V -> V^-1 for every cluster matrix of the pool (work is destroyed); info receives (first failing cluster + 1)
void launch_cluster_inverse(const ClusterDesc* clusters, uint32_t nclusters, double* work, double* out, int* info, void* stream);

// diagonal blocks <- per-station constraint block (FormConstraintStationVarianceMatrix, ADJ:2041-2137)
void launch_init_normals(const double* cblock, double* ndiag, double* noff, double* w, uint32_t nstn, uint64_t nedge,
                         void* stream);

struct ScatterParams {
    const double* ndiag;
    const double* noff;
    const double* bq;            // per-baseline slots of the assembly (V^-1 upper triangle + V^-1 l)
    const uint32_t* edge_bsl;    // per edge: the one GNSS baseline that forms its block (N[hi,lo] = -V^-1, read from bq), or ~0u: block is in noff
    const uint64_t* diag_dest;   // per station (station order); ~0 = assembled by another rank
    const uint32_t* diag_ld;
    const uint64_t* off_dest;    // per edge
    const uint32_t* off_ld;
    const uint32_t* edge_hi;     // per edge: station of the row block (eliminated later)
    const uint32_t* edge_lo;
    double* panels;
    double* dscale;              // 3 x nstn, station order: 1/sqrt(N_ii) (or 1 when scaling is off)
    uint32_t nstn;
    uint64_t nedge;
    int32_t scale;
};
void launch_compute_scale(const ScatterParams& p, void* stream);
void launch_scatter_normals(const ScatterParams& p, void* stream);

// b[3*pos[s]+c] = dscale[3s+c] * w[3s+c]   (0 where pos_owned[pos[s]] == 0, when a mask is given)
void launch_permute_rhs(const double* w, const double* dscale, const uint32_t* pos_of_stn, const uint8_t* pos_owned,
                        double* b, uint32_t nstn, void* stream);
// x[3p+c] = 0 for positions this rank does not own (before the cross-rank sum of the solution vector)
void launch_mask_positions(double* x, const uint8_t* pos_owned, uint32_t nstn, void* stream);
// corr[3s+c] = dscale * x[3*pos[s]+c]; est += corr; tracks the largest |corr| (first in station order on ties)
// scratch: APPLY_SCRATCH_DOUBLES doubles of the context (per-CTA partial maxima)
constexpr int APPLY_MAX_PARTS = 148 * 8;
constexpr int APPLY_SCRATCH_DOUBLES = 2 * APPLY_MAX_PARTS;
void launch_apply_corrections(const double* x, const double* dscale, const uint32_t* pos_of_stn, double* corr,
                              double* est, uint32_t nstn, double* scratch, void* stream);
// vcv[s] (9 doubles, row-major) = dscale_i * Z_ss * dscale_j read from the panels (lower triangle mirrored)
void launch_extract_station_vcv(const double* panels, const uint64_t* diag_dest, const uint32_t* diag_ld,
                                const double* dscale, double* vcv, uint32_t nstn, void* stream);
// q (9 doubles) per edge: dscale-scaled off-diagonal block N^-1[hi station, lo station]
void launch_extract_edge_vcv(const double* panels, const uint64_t* off_dest, const uint32_t* off_ld,
                             const uint32_t* edge_hi, const uint32_t* edge_lo, const double* dscale, double* q,
                             uint64_t nedge, void* stream);

struct StatsParams {
    dna_msr_t* msr;
    const uint32_t* first;
    const uint32_t* edge;
    const double* est;
    const double* vcv_diag;      // nstn x 9
    const double* vcv_off;       // nedge x 9 : N^-1[hi, lo]
    double* sums;                // [0] chi2, [1] pelzer sum, [2] pelzer count, [3] outliers
    uint64_t nbaselines;
    double critical;
};
void launch_stats_g(const StatsParams& p, void* stream);

// geographic <- Cartesian for every station (CartToGeo, GEO:154-225)
void launch_cart_to_geo(const double* est, double* llh, uint32_t nstn, double a, double invf, void* stream);

}  // namespace gadj
