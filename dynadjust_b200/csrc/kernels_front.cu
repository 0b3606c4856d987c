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
//            entries in registers (two barriers per column); trailing update: 4 x 4 register tiles over all
//            256 threads, rank-16 per step.
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
    __shared__ double colbuf[PBW];
    __shared__ double piv;
    const DiagOp op = ops[blockIdx.x];
    const int w = op.w, tid = threadIdx.x;
    const int64_t ld = op.ldd;
    const int wpad = (w + PBW - 1) & ~(PBW - 1);
    // coalesced load of the lower triangle (row-major source); identity padding
    for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
        const int i = idx >> 7, j = idx & (NB - 1);
        if (j <= i)
            S[tri(i, j)] = i < w ? op.D[i * ld + j] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    if (op.factor) {
        for (int p0 = 0; p0 < wpad; p0 += PBW) {
            // ---- panel: rows p0 .. wpad-1, columns p0 .. p0+15
            const int i = tid;
            const bool act = tid < wpad && i >= p0;
            double r[PBW];
            if (act) {
#pragma unroll
                for (int kk = 0; kk < PBW; ++kk)
                    r[kk] = tri_ld(S, i, p0 + kk);
            }
#pragma unroll
            for (int jj = 0; jj < PBW; ++jj) {
                if (tid == p0 + jj) {
                    double d = r[jj];
                    if (!(d > 0.0)) {
                        atomicCAS(info, 0, op.front + 1);
                        d = 1.0;
                    }
                    d = sqrt(d);
                    r[jj] = d;
                    piv = 1.0 / d;      // one division per column; the rows below multiply
                }
                __syncthreads();
                const bool below = act && i > p0 + jj;
                if (below) {
                    r[jj] = r[jj] * piv;
                    if (i < p0 + PBW)
                        colbuf[i - p0] = r[jj];
                }
                __syncthreads();
                if (below) {
                    const double l = r[jj];
#pragma unroll
                    for (int kk = jj + 1; kk < PBW; ++kk)
                        r[kk] -= l * colbuf[kk];
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
        for (int idx = tid; idx < w * w; idx += DIAG_THREADS) {
            const int i = idx / w, j = idx - i * w;
            if (j <= i)
                op.D[i * ld + j] = S[tri(i, j)];
        }
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
    for (int idx = tid; idx < w * w; idx += DIAG_THREADS) {
        const int i = idx / w, j = idx - i * w;
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
// and its mirror image are written with coalesced rows.
constexpr int GT = 16;            // stations per tile edge
constexpr int GE = 3 * GT;        // doubles per tile edge
__global__ void __launch_bounds__(256) gather_kernel(const GatherOp* __restrict__ ops, const GatherTile* __restrict__ tiles, int ntiles)
{
    __shared__ double T[GE][GE + 1];
    // the planner lists exactly the tiles on or below the diagonal of every op; the CTAs stride through the list
    for (int tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
        const GatherTile gt = tiles[tl];
        const GatherOp op = ops[gt.op];
        const int ni = op.nb - op.jb, nj = op.je - op.jb;
        const int i0 = gt.ti * GT, j0 = gt.tj * GT;      // station offsets relative to jb
        for (int e = threadIdx.x; e < GE * GE; e += 256) {
            const int ar = e / GE, bc = e - ar * GE;
            const int ii = i0 + ar / 3, jj = j0 + bc / 3;
            int a = ar % 3, b = bc % 3;
            double v = 0.0;
            if (ii < ni && jj < nj && jj <= ii) {
                if (ii == jj && a < b) {
                    const int t = a;
                    a = b;
                    b = t;
                }
                v = op.Z[(3ll * op.rowmap[ii] + a) * op.ld + 3ll * op.rowmap[jj] + b];
            }
            T[ar][bc] = v;
        }
        __syncthreads();
        // lower part: rows of G along i, contiguous along j
        for (int e = threadIdx.x; e < GE * GE; e += 256) {
            const int ar = e / GE, bc = e - ar * GE;
            const int ii = i0 + ar / 3, jj = j0 + bc / 3;
            if (ii < ni && jj < nj && jj <= ii)
                op.G[(3ll * (op.jb + ii) + ar % 3) * op.ldg + 3 * (op.jb + jj) + bc % 3] = T[ar][bc];
        }
        // mirror: rows of G along j, contiguous along i
        for (int e = threadIdx.x; e < GE * GE; e += 256) {
            const int bc = e / GE, ar = e - bc * GE;
            const int ii = i0 + ar / 3, jj = j0 + bc / 3;
            if (ii < ni && jj < nj && jj <= ii)
                op.G[(3ll * (op.jb + jj) + bc % 3) * op.ldg + 3 * (op.jb + ii) + ar % 3] = T[ar][bc];
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
