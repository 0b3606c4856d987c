// kernels_front.cu — per-front helper kernels around the tile GEMM:
//   pivot-tile Cholesky + triangular inverse, triangular matrix-vector products, panel GEMV updates for the
//   forward/backward substitution, transposes and the selected-inverse gather.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "dev.h"
#include "kernels.h"

namespace gadj {
namespace {

// ---- pivot tile: L = chol(D), W = L^-1 ----------------------------------------------
// One CTA (256 threads) per tile.  Only the lower triangle is ever needed, so the tile lives in shared memory in
// packed form (row i at offset i(i+1)/2: 66 KB instead of 132 KB) and three CTAs share an SM — the kernel is bound
// by barrier and shared-memory latency, not by arithmetic, so co-resident tiles are what fills the SM.  The tile is
// padded to 128 x 128 with an identity so that every step runs on full blocks.
//   factor : blocked right-looking Cholesky, panel width 16.  Panel: one thread per row keeps its 16 panel
//            entries in registers; the 16 x 16 diagonal block is factorised inside one warp with shuffles, the rows
//            below it by substitution against the finished block (three CTA barriers per panel instead of two per
//            column: the kernel is latency-bound — one tile per launch at the top of the tree); trailing update:
//            4 x 4 register tiles over all 256 threads, rank-16 per step.
//   inverse: in place in the lower triangle (L has been stored to global memory by then): 16 x 16 diagonal
//            blocks by forward substitution in registers, then block doubling 16 -> 32 -> 64 -> 128 with
//            W21 = -W22 (L21 W11) as two register-tiled products per level.
constexpr int DIAG_THREADS = 256;
constexpr int PBW = 16;
constexpr int TRI_DOUBLES = NB * (NB + 1) / 2;

__device__ __forceinline__ int tri(int i, int j) { return ((i * (i + 1)) >> 1) + j; }            // j <= i
__device__ __forceinline__ double tri_ld(const double* __restrict__ S, int i, int j)              // zero above the diagonal
{
    return j <= i ? S[tri(i, j)] : 0.0;
}

// acc[x][y] += sum_k P(prow + ti + tstride x, pcol + k) * Q(qrow + k, qcol + tj + tstride y),  k in [0, kn);
// PTRI / QTRI: the operand block sits on the diagonal (lower triangular: entries above it read as zero)
template <bool PTRI, bool QTRI>
__device__ __forceinline__ void tile_mm(const double* __restrict__ S, int prow, int pcol, int qrow, int qcol, int ti, int tj,
                                        int tstride, int kn, double acc[4][4])
{
#pragma unroll 4
    for (int k = 0; k < kn; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
            a[x] = PTRI ? tri_ld(S, prow + ti + tstride * x, pcol + k) : S[tri(prow + ti + tstride * x, pcol + k)];
#pragma unroll
        for (int y = 0; y < 4; ++y)
            b[y] = QTRI ? tri_ld(S, qrow + k, qcol + tj + tstride * y) : S[tri(qrow + k, qcol + tj + tstride * y)];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y)
                acc[x][y] += a[x] * b[y];
    }
}

template <int CTAS_PER_SM>
__global__ void __launch_bounds__(DIAG_THREADS, CTAS_PER_SM) diag_kernel(const DiagOp* __restrict__ ops, int* __restrict__ info)
{
    extern __shared__ double S[];
    __shared__ double Lb[PBW][PBW + 1];   // the finished 16 x 16 diagonal block of the current panel
    __shared__ double pinv[PBW];          // reciprocals of its diagonal
    const DiagOp op = ops[blockIdx.x];
    const int w = op.w, tid = threadIdx.x;
    const int64_t ld = op.ldd;
    const int wpad = (w + PBW - 1) & ~(PBW - 1);
    // coalesced load of the lower triangle (row-major source), identity padding; a warp takes rows warp + 8 t and issues
    // the loads of four rows together (one tile per launch at the top of the tree: latency is all that counts here)
    const int warp = tid >> 5, lane = tid & 31;
    for (int i0 = warp; i0 < NB; i0 += 32) {
        double v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 8 * u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = lane + 32 * q;
                v[u][q] = (j <= i && i < w) ? op.D[i * ld + j] : (i == j ? 1.0 : 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 8 * u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = lane + 32 * q;
                if (j <= i)
                    S[tri(i, j)] = v[u][q];
            }
        }
    }
    __syncthreads();
    if (op.factor) {
        for (int p0 = 0; p0 < wpad; p0 += PBW) {
            // ---- panel: rows p0 .. wpad-1, columns p0 .. p0+15; one thread per row keeps its 16 entries in registers
            const int i = tid;
            const bool act = tid < wpad && i >= p0;
            double r[PBW];
            if (act) {
#pragma unroll
                for (int kk = 0; kk < PBW; ++kk)
                    r[kk] = tri_ld(S, i, p0 + kk);
            }
            // (a) the 16 x 16 diagonal block, by the warp that holds its rows (16 consecutive lanes), through shuffles:
            //     no CTA barrier inside the 16 column steps
            if ((tid >> 5) == (p0 >> 5)) {
                const int lane = tid & 31, l0 = p0 & 31;
                const bool inblk = lane >= l0 && lane < l0 + PBW;
#pragma unroll
                for (int jj = 0; jj < PBW; ++jj) {
                    double d = __shfl_sync(0xffffffffu, r[jj], l0 + jj);
                    if (!(d > 0.0)) {
                        if (lane == l0 + jj)
                            atomicCAS(info, 0, op.front + 1);
                        d = 1.0;
                    }
                    d = sqrt(d);
                    const double pv = 1.0 / d;      // one division per column; the rows below multiply
                    const bool below = inblk && lane > l0 + jj;
                    if (lane == l0 + jj) {
                        r[jj] = d;
                        pinv[jj] = pv;
                    } else if (below)
                        r[jj] = r[jj] * pv;
#pragma unroll
                    for (int kk = 0; kk < PBW; ++kk) {
                        if (kk <= jj)
                            continue;   // (fixed bounds: the loops unroll completely and r[] stays in registers)
                        const double c = __shfl_sync(0xffffffffu, r[jj], l0 + kk);   // L[p0 + kk][p0 + jj]
                        if (below)
                            r[kk] -= r[jj] * c;
                    }
                }
                if (inblk) {
#pragma unroll
                    for (int kk = 0; kk < PBW; ++kk)
                        Lb[lane - l0][kk] = r[kk];
                }
            }
            __syncthreads();
            // (b) the rows below the block: X L11^T = R by substitution against the finished block (broadcast reads),
            //     the same operations in the same order as a right-looking sweep, without its two barriers per column
            if (act && i >= p0 + PBW) {
#pragma unroll
                for (int jj = 0; jj < PBW; ++jj) {
                    double v = r[jj];
#pragma unroll
                    for (int kk = 0; kk < PBW; ++kk)
                        if (kk < jj)
                            v -= r[kk] * Lb[jj][kk];
                    r[jj] = v * pinv[jj];
                }
            }
            if (act) {
#pragma unroll
                for (int kk = 0; kk < PBW; ++kk)
                    if (p0 + kk <= i)
                        S[tri(i, p0 + kk)] = r[kk];
            }
            __syncthreads();
            // ---- trailing update: S[i][k] -= sum_jj S[i][p0+jj] S[k][p0+jj] for rows / columns >= p0 + 16
            const int t0 = p0 + PBW, tn = wpad - t0;
            if (tn > 0) {
                const int nt = tn >> 2;               // threads per dimension; each owns rows ti + nt x, columns tj + nt y
                for (int t = tid; t < nt * nt; t += DIAG_THREADS) {
                    const int ti = t / nt, tj = t - ti * nt;
                    double acc[4][4] = {};
#pragma unroll 4
                    for (int k = 0; k < PBW; ++k) {
                        double a[4], b[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x)
                            a[x] = S[tri(t0 + ti + nt * x, p0 + k)];
#pragma unroll
                        for (int y = 0; y < 4; ++y)
                            b[y] = S[tri(t0 + tj + nt * y, p0 + k)];
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = 0; y < 4; ++y)
                                acc[x][y] += a[x] * b[y];
                    }
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            const int ii = t0 + ti + nt * x, kk = t0 + tj + nt * y;
                            if (kk <= ii)
                                S[tri(ii, kk)] -= acc[x][y];
                        }
                }
            }
            __syncthreads();
        }
        for (int i = warp; i < w; i += DIAG_THREADS / 32)
            for (int j = lane; j <= i; j += 32)
                op.D[i * ld + j] = S[tri(i, j)];
    }
    if (op.W == nullptr && op.Wt == nullptr)
        return;
    __syncthreads();
    // ---- W = L^-1, in place in the lower triangle -------------------------------------------------
    {   // 16 x 16 diagonal blocks: thread (block d, column c) solves L x = e_c in registers
        const int d = tid >> 4, c = tid & 15, base = d * PBW;
        double x[PBW];
        const bool act = base < wpad;
        if (act) {
#pragma unroll
            for (int i = 0; i < PBW; ++i) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k)
                    s += S[tri(base + i, base + k)] * x[k];
                const double dii = S[tri(base + i, base + i)];
                x[i] = (i < c) ? 0.0 : (i == c ? 1.0 / dii : -s / dii);
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int i = 0; i < PBW; ++i)
                if (i >= c)
                    S[tri(base + i, base + c)] = x[i];
        }
        __syncthreads();
    }
    for (int b = PBW; b < wpad; b <<= 1) {
        // pairs of finished b x b diagonal blocks: W21 = -W22 (L21 W11)
        const int q4 = b >> 2, tpp = q4 * q4;
        const int pair = tid / tpp, rem = tid - pair * tpp;
        const int ti = rem / q4, tj = rem - ti * q4;
        const int r0 = pair * 2 * b;
        const bool act = r0 + b < wpad;
        double acc[4][4] = {};
        if (act)
            tile_mm<false, true>(S, r0 + b, r0, r0, r0, ti, tj, q4, b, acc);            // T = L21 W11
        __syncthreads();
        if (act) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    S[tri(r0 + b + ti + q4 * x, r0 + tj + q4 * y)] = acc[x][y];
                    acc[x][y] = 0.0;
                }
        }
        __syncthreads();
        if (act)
            tile_mm<true, false>(S, r0 + b, r0 + b, r0 + b, r0, ti, tj, q4, b, acc);    // W22 T
        __syncthreads();
        if (act) {
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y)
                    S[tri(r0 + b + ti + q4 * x, r0 + tj + q4 * y)] = -acc[x][y];
        }
        __syncthreads();
    }
    for (int i = warp; i < w; i += DIAG_THREADS / 32)
        for (int j = lane; j < w; j += 32) {
            if (op.W)
                op.W[i * op.ldw + j] = (j <= i) ? S[tri(i, j)] : 0.0;     // W[i][j] (row-major, lower)
            if (op.Wt)
                op.Wt[i * op.ldwt + j] = (j >= i) ? S[tri(j, i)] : 0.0;   // Wt[i][j] = W[j][i] (upper)
        }
}

// ---- multi-GPU: finished regions of a replicated front pushed into every peer's replica ----------------------
// One op = a rows x cols block of a replicated buffer (a tile a rank has just computed, or a whole panel its owner has
// just factorised); a warp moves one row at a time with 16-byte accesses, so the NVLink traffic is whole lines.
// (Storing the tiles into the peers straight from the GEMM epilogue was tried first: the scattered 16-byte remote stores
// of 288 threads stalled the tensor pipe — 237 ms per C4 iteration on 8 GPUs against 204 ms for the host-driven NCCL
// exchange of round 1, profiles/r2_bench_8gpu_c4_first.json.)
__global__ void __launch_bounds__(256) push_kernel(const PushOp* __restrict__ ops, const PeerTable* __restrict__ pt, double* const* __restrict__ bases)
{
    const PushOp op = ops[blockIdx.y];
    const int n = pt->nranks, me = pt->rank;
    const int64_t* d = pt->delta[op.buf];
    // destinations of this block: every peer, or one rank only, possibly leaving out a rank that has the block already
    unsigned dst = 0;
    for (int q = 0; q < n; ++q)
        if (q != me && q != op.skip && (op.target < 0 || q == op.target))
            dst |= 1u << q;
    double* base = bases[op.buf] + op.off;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool vec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((op.ld & 1) == 0);
    for (int r = blockIdx.x * 8 + warp; r < op.rows; r += gridDim.x * 8) {
        double* row = base + (int64_t)r * op.ld;
        if (vec) {
            const int pairs = op.cols >> 1;
            for (int i = lane; i < pairs; i += 32) {
                const double2 v = reinterpret_cast<const double2*>(row)[i];
                for (int q = 0; q < n; ++q)
                    if (dst >> q & 1)
                        reinterpret_cast<double2*>(reinterpret_cast<char*>(row) + d[q])[i] = v;
            }
            if ((op.cols & 1) && lane == 0) {
                const double v = row[op.cols - 1];
                for (int q = 0; q < n; ++q)
                    if (dst >> q & 1)
                        *reinterpret_cast<double*>(reinterpret_cast<char*>(row + op.cols - 1) + d[q]) = v;
            }
        } else {
            for (int i = lane; i < op.cols; i += 32) {
                const double v = row[i];
                for (int q = 0; q < n; ++q)
                    if (dst >> q & 1)
                        *reinterpret_cast<double*>(reinterpret_cast<char*>(row + i) + d[q]) = v;
            }
        }
    }
}

// ---- multi-GPU: barrier, all-reduce, shared status -------------------------------------------------
// Every rank adds one to every rank's counter (its own included) and waits until its own counter has seen all the ranks
// of this round: counter == rounds * nranks.  The launches before the barrier on every rank — stores into peers'
// replicas included — have completed when the stream reaches this kernel; the system-scope fences order them against
// the counter updates and the reads that follow.  A rank that waits for more than ~20 s (a peer has failed) flags
// info[1] and leaves; the host reports it.
__global__ void barrier_kernel(const PeerTable* __restrict__ pt, unsigned long long target, int* __restrict__ info)
{
    const int p = threadIdx.x;
    __threadfence_system();
    if (p < pt->nranks)
        atomicAdd_system(pt->counter[p], 1ull);
    if (p == 0) {
        volatile unsigned long long* c = pt->counter[pt->rank];
        const long long t0 = clock64();
        while (*c < target) {
            __nanosleep(64);
            if (clock64() - t0 > 40000000000ll) {
                atomicExch(info + 1, 1);
                break;
            }
        }
        __threadfence_system();
    }
}

// Sum over the ranks' replicas of the ranges in `ops` (doubles at base + off), the result stored into every replica.
// Rank r reduces the r-th slice of every range, reading the other replicas over NVLink and adding them in rank order,
// so all replicas receive identical bits (and the same bits run after run).
__global__ void __launch_bounds__(256) allreduce_kernel(const ReduceOp* __restrict__ ops, const PeerTable* __restrict__ pt,
                                                       double* __restrict__ base, int buf)
{
    const ReduceOp op = ops[blockIdx.y];
    const int n = pt->nranks, me = pt->rank;
    const int64_t* d = pt->delta[buf];
    // slices in units of two doubles when the range starts on a 16-byte boundary (vector accesses), else single doubles
    double* p0 = base + op.off;
    const bool vec = ((reinterpret_cast<uintptr_t>(p0) & 15) == 0);
    if (vec) {
        const uint64_t pairs = op.count / 2;
        const uint64_t lo = pairs * me / n, hi = pairs * (me + 1) / n;
        for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
            double2 s = make_double2(0.0, 0.0);
            for (int q = 0; q < n; ++q) {
                const double2 v = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(p0 + 2 * i) + d[q]);
                s.x += v.x;
                s.y += v.y;
            }
            for (int q = 0; q < n; ++q)
                *reinterpret_cast<double2*>(reinterpret_cast<char*>(p0 + 2 * i) + d[q]) = s;
        }
        if ((op.count & 1) && me == n - 1 && blockIdx.x == 0 && threadIdx.x == 0) {
            const uint64_t i = op.count - 1;
            double s = 0.0;
            for (int q = 0; q < n; ++q)
                s += *reinterpret_cast<const double*>(reinterpret_cast<const char*>(p0 + i) + d[q]);
            for (int q = 0; q < n; ++q)
                *reinterpret_cast<double*>(reinterpret_cast<char*>(p0 + i) + d[q]) = s;
        }
    } else {
        const uint64_t lo = op.count * me / n, hi = op.count * (me + 1) / n;
        for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x) {
            double s = 0.0;
            for (int q = 0; q < n; ++q)
                s += *reinterpret_cast<const double*>(reinterpret_cast<const char*>(p0 + i) + d[q]);
            for (int q = 0; q < n; ++q)
                *reinterpret_cast<double*>(reinterpret_cast<char*>(p0 + i) + d[q]) = s;
        }
    }
}

__global__ void share_info_kernel(const PeerTable* __restrict__ pt, int* __restrict__ info)
{
    const int p = threadIdx.x;
    if (p < pt->nranks && p != pt->rank) {
        if (info[0] != 0)
            atomicMax_system(pt->info[p], info[0]);
        if (info[1] != 0)
            atomicMax_system(pt->info[p] + 1, info[1]);
    }
}

// ---- triangular matrix-vector product with the inverse pivot block ------------------------------
// y[row0 + i] = sum_c A[i][c] x[c] over the triangle only (lower: c <= row, upper: c >= row).  One CTA per chunk of
// up to 64 rows; each warp owns groups of 4 rows and walks the columns with coalesced 32-wide strides, so one load
// of x feeds four rows.  HBM-bound: the triangle of W / Wt is read once per substitution.
constexpr int TRIMV_THREADS = 256;
__global__ void __launch_bounds__(TRIMV_THREADS) trimv_kernel(const TrimvOp* __restrict__ ops)
{
    const TrimvOp op = ops[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* __restrict__ x = op.x;
    if (op.wide) {
        // wide fronts (k >= TRIMV_WIDE_K: few fronts per level, long rows): up to 8 rows per CTA, the 256 threads stride
        // across the columns — 8 loads in flight per thread and k / 8 CTAs per front, where the row-group form below left
        // most of the GPU idle (a 3 400-wide top front: 53 CTAs, 60 GB/s)
        __shared__ double red[TRIMV_THREADS / 32][TRIMV_WIDE_ROWS];
        const int gr = op.row0, nr = op.nrows;
        const int c_lo = op.upper ? gr : 0;
        const int c_hi = op.upper ? op.k : gr + nr;        // exclusive
        double acc[TRIMV_WIDE_ROWS];
#pragma unroll
        for (int r = 0; r < TRIMV_WIDE_ROWS; ++r)
            acc[r] = 0.0;
        for (int c = (c_lo & ~31) + (int)threadIdx.x; c < c_hi; c += TRIMV_THREADS) {
            if (c < c_lo)
                continue;
            const double xv = x[c];
#pragma unroll
            for (int r = 0; r < TRIMV_WIDE_ROWS; ++r) {
                const bool in = r < nr && (op.upper ? c >= gr + r : c <= gr + r);
                if (in)
                    acc[r] += op.A[(int64_t)r * op.ld + c] * xv;
            }
        }
#pragma unroll
        for (int r = 0; r < TRIMV_WIDE_ROWS; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
            if (lane == 0)
                red[warp][r] = acc[r];
        }
        __syncthreads();
        if ((int)threadIdx.x < nr) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < TRIMV_THREADS / 32; ++w)
                s += red[w][threadIdx.x];
            op.y[gr + threadIdx.x] = s;
        }
        return;
    }
    for (int g0 = 4 * warp; g0 < op.nrows; g0 += 4 * (TRIMV_THREADS / 32)) {
        const int gr = op.row0 + g0;                       // first row (front numbering) of the group
        const int nr = op.nrows - g0 < 4 ? op.nrows - g0 : 4;
        const int c_lo = op.upper ? gr : 0;
        const int c_hi = op.upper ? op.k : gr + nr;        // exclusive
        const double* __restrict__ a0 = op.A + (int64_t)g0 * op.ld;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int c = (c_lo & ~31) + lane; c < c_hi; c += 32) {
            if (c < c_lo)
                continue;
            const double xv = x[c];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool in = r < nr && (op.upper ? c >= gr + r : c <= gr + r);
                if (in)
                    acc[r] += a0[(int64_t)r * op.ld + c] * xv;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
        if (lane < nr)
            op.y[gr + lane] = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
    }
}

// ---- panel GEMV updates of the substitution ---------------------------------------------
// forward : x[rowidx[i]] -= P[i][:] . xj          one warp per row, lanes along the tile columns
__global__ void __launch_bounds__(256) gemv_fwd_kernel(const GemvOp* __restrict__ ops, double* __restrict__ x)
{
    __shared__ double xj[NB];
    const GemvOp op = ops[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < NB)
        xj[threadIdx.x] = threadIdx.x < op.w ? op.xj[threadIdx.x] : 0.0;
    __syncthreads();
    for (int i = warp; i < op.nrows; i += 8) {
        const double* __restrict__ p = op.P + (int64_t)i * op.ld;
        double s = 0.0;
        for (int c = lane; c < op.w; c += 32)
            s += p[c] * xj[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && s != 0.0)
            atomicAdd(x + op.rowidx[i], -s);
    }
}

// backward: xj[c] -= sum_i P[i][c] * x[rowidx[i]]   threads along the tile columns, two row halves
__global__ void __launch_bounds__(256) gemv_bwd_kernel(const GemvOp* __restrict__ ops, const double* __restrict__ x)
{
    __shared__ double xr[256];
    __shared__ double part[NB];
    const GemvOp op = ops[blockIdx.x];
    const int tid = threadIdx.x;
    xr[tid] = tid < op.nrows ? x[op.rowidx[tid]] : 0.0;
    __syncthreads();
    const int c = tid & (NB - 1), half = tid >> 7;
    double s = 0.0;
    if (c < op.w)
        for (int i = half; i < op.nrows; i += 2)
            s += op.P[(int64_t)i * op.ld + c] * xr[i];
    if (half == 1)
        part[c] = s;
    __syncthreads();
    if (half == 0 && c < op.w) {
        s += part[c];
        if (s != 0.0)
            atomicAdd(op.xj + c, -s);
    }
}

// ---- transpose --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const TransposeOp* __restrict__ ops)
{
    __shared__ double tile[32][33];
    const TransposeOp op = ops[blockIdx.y];
    const int tiles_c = (op.cols + 31) >> 5, tiles_r = (op.rows + 31) >> 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int tl = blockIdx.x; tl < tiles_c * tiles_r; tl += gridDim.x) {
        const int tr = tl / tiles_c, tc = tl - tr * tiles_c;
        const int r0 = tr << 5, c0 = tc << 5;
        for (int k = ty; k < 32; k += 8) {
            int r = r0 + k, c = c0 + tx;
            tile[k][tx] = (r < op.rows && c < op.cols) ? op.src[(int64_t)r * op.lds + c] : 0.0;
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {
            int c = c0 + k, r = r0 + tx;  // dst row = source column
            if (c < op.cols && r < op.rows)
                op.dst[(int64_t)c * op.ldd + r] = tile[tx][k];
        }
        __syncthreads();
    }
}

// ---- selected-inverse gather ---------------------------------------------------------------
// 16 x 16-station tiles (48 x 48 doubles) staged through shared memory so that both the lower block
// and its mirror image are written with coalesced rows.  Read: one 3 x 3 station block per thread (256 blocks per
// tile), so the index arithmetic — two row-map look-ups and one 64-bit address — is done once per block; write: one
// tile row per warp and pass.  (The first version indexed element by element and was bound by its integer
// arithmetic: ncu showed 69 % issue-slot use at 10 % of the DRAM bandwidth, profiles/r2_gather_ncu_summary.txt.)
constexpr int GT = 16;            // stations per tile edge
constexpr int GE = 3 * GT;        // doubles per tile edge
__global__ void __launch_bounds__(256) gather_kernel(const GatherOp* __restrict__ ops, const GatherTile* __restrict__ tiles, int ntiles)
{
    __shared__ double T[GE][GE + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int bi = tid >> 4, bj = tid & 15;            // this thread's station block of the tile
    const int c0 = lane, c1 = lane + 32;               // the two tile columns this lane writes (c1 < 48 for lanes 0..15)
    const int cs0 = c0 / 3, cs1 = c1 / 3;              // their stations
    int cur = -1;
    GatherOp op;
    // the planner lists exactly the tiles on or below the diagonal of every op; the CTAs stride through the list
    for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const GatherTile gt = tiles[tl];
        if (gt.op != cur) {                            // (uniform) consecutive tiles mostly belong to one op
            op = ops[gt.op];
            cur = gt.op;
        }
        const int ni = op.nb - op.jb, nj = op.je - op.jb;
        const int i0 = gt.ti * GT, j0 = gt.tj * GT;      // station offsets relative to jb
        {
            const int ii = i0 + bi, jj = j0 + bj;
            double v[9];
#pragma unroll
            for (int e = 0; e < 9; ++e)
                v[e] = 0.0;
            if (ii < ni && jj < nj && jj <= ii) {
                const double* __restrict__ z = op.Z + 3ll * op.rowmap[ii] * op.ld + 3ll * op.rowmap[jj];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b)
                        v[3 * a + b] = z[a * op.ld + b];
                if (ii == jj) {                        // a diagonal block: the panel holds its lower triangle only
                    v[1] = v[3];
                    v[2] = v[6];
                    v[5] = v[7];
                }
            }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    T[3 * bi + a][3 * bj + b] = v[3 * a + b];
        }
        __syncthreads();
        // lower part: rows of G along i, contiguous along j
        for (int rr = warp; rr < GE; rr += 8) {
            const int ii = i0 + rr / 3;
            if (ii >= ni)
                break;
            double* __restrict__ g = op.G + (3ll * (op.jb + i0) + rr) * op.ldg + 3 * (op.jb + j0);
            if (j0 + cs0 < nj && j0 + cs0 <= ii)
                g[c0] = T[rr][c0];
            if (c1 < GE && j0 + cs1 < nj && j0 + cs1 <= ii)
                g[c1] = T[rr][c1];
        }
        // mirror: rows of G along j, contiguous along i
        for (int cc = warp; cc < GE; cc += 8) {
            const int jj = j0 + cc / 3;
            if (jj >= nj)
                break;
            double* __restrict__ g = op.G + (3ll * (op.jb + j0) + cc) * op.ldg + 3 * (op.jb + i0);
            if (i0 + cs0 < ni && jj <= i0 + cs0)
                g[c0] = T[c0][cc];
            if (c1 < GE && i0 + cs1 < ni && jj <= i0 + cs1)
                g[c1] = T[c1][cc];
        }
        __syncthreads();
    }
}

constexpr int TILE_SMEM = TRI_DOUBLES * 8;

}  // namespace

void launch_diag(const DiagOp* ops, int nops, int* info, void* stream)
{
    if (nops <= 0)
        return;
    static const int ctas = [] {
        const char* e = getenv("GADJ_DIAG_CTAS");   // register budget: 2 CTAs/SM without spills, 3 with a few spilled values
        return (e && e[0] == '3') ? 3 : 2;
    }();
    if (dev::first_use(KEY_DIAG)) {
        cudaFuncSetAttribute(diag_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM);
        cudaFuncSetAttribute(diag_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM);
    }
    if (ctas == 3)
        diag_kernel<3><<<nops, DIAG_THREADS, TILE_SMEM, (cudaStream_t)stream>>>(ops, info);
    else
        diag_kernel<2><<<nops, DIAG_THREADS, TILE_SMEM, (cudaStream_t)stream>>>(ops, info);
}

void launch_barrier(const PeerTable* pt, unsigned long long target, int* info, void* stream)
{
    barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pt, target, info);
}

void launch_allreduce(const ReduceOp* ops, int nops, const PeerTable* pt, double* base, int buf, void* stream)
{
    if (nops <= 0)
        return;
    allreduce_kernel<<<dim3(148 * 2, nops), 256, 0, (cudaStream_t)stream>>>(ops, pt, base, buf);
}

void launch_push(const PushOp* ops, int nops, int grid_x, const PeerTable* pt, double* const* bases, void* stream)
{
    for (int o = 0; o < nops; o += 65535) {
        const int n = nops - o < 65535 ? nops - o : 65535;
        push_kernel<<<dim3(grid_x < 1 ? 1 : (grid_x > 64 ? 64 : grid_x), n), 256, 0, (cudaStream_t)stream>>>(ops + o, pt, bases);
    }
}

void launch_share_info(const PeerTable* pt, int* info, void* stream)
{
    share_info_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pt, info);
}

void launch_trimv(const TrimvOp* ops, int nops, void* stream)
{
    if (nops <= 0)
        return;
    trimv_kernel<<<nops, TRIMV_THREADS, 0, (cudaStream_t)stream>>>(ops);
}

void launch_gemv(const GemvOp* ops, int nops, const double* x_ro, double* x, int backward, void* stream)
{
    if (nops <= 0)
        return;
    if (backward)
        gemv_bwd_kernel<<<nops, 256, 0, (cudaStream_t)stream>>>(ops, x_ro);
    else
        gemv_fwd_kernel<<<nops, 256, 0, (cudaStream_t)stream>>>(ops, x);
}

void launch_transpose(const TransposeOp* ops, int nops, int grid_x, void* stream)
{
    if (nops <= 0)
        return;
    for (int o = 0; o < nops; o += 65535) {
        int n = nops - o < 65535 ? nops - o : 65535;
        transpose_kernel<<<dim3(grid_x < 1 ? 1 : (grid_x > 592 ? 592 : grid_x), n), 256, 0, (cudaStream_t)stream>>>(ops + o);
    }
}

void launch_gather(const GatherOp* ops, int nops, const GatherTile* tiles, int ntiles, void* stream)
{
    if (nops <= 0 || ntiles <= 0)
        return;
    static_assert(GATHER_TILE_STATIONS == GT, "the planner's tile edge");
    const int cap = dev::sm_count() * 8;   // 8 resident CTAs per SM (19 KB of shared memory each) keep enough loads in flight
    gather_kernel<<<ntiles < cap ? ntiles : cap, 256, 0, (cudaStream_t)stream>>>(ops, tiles, ntiles);
}

}  // namespace gadj
