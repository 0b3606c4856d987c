// kernels_gemm.cu — batched FP64 tile GEMM on the DMMA tensor path, fed by TMA.
//
//   C[M x N] (+)= alpha * A[M x K] * B[N x K]^T        (all row-major, K contiguous)
//
// The unit of work is one output tile of one op; a launch covers the tiles of many ops (all the
// panel updates / Schur complements / inverse products of one tree level), listed by the planner (only tiles
// that hold work), and persistent CTAs stride through the list.  Two tile shapes, chosen per launch by the planner:
//   128 x 128 — 8 consumer warps (2 x 4, 64 x 32 each), a 132 KB ring, one CTA per SM: the large fronts;
//    64 x 64  — 4 consumer warps (2 x 2, 32 x 32 each), a 66 KB ring, three CTAs per SM: the small fronts of the low
//               tree levels, whose 128-wide tiles are mostly padding and whose short K loops leave the epilogue
//               (stores, RED.ADD scatter) exposed — with three resident CTAs one tile's epilogue runs under the
//               others' products.
//
// Data path: a producer warp issues cp.async.bulk.tensor (TMA) loads of TM x 16 / TN x 16 FP64
// boxes of A and B into a 4-stage shared-memory ring (SWIZZLE_128B, mbarrier full/empty
// pairs); the consumer warps (2 along M, 4 or 2 along N) read fragments with
// bank-conflict-free LDS.64 and issue mma.sync.m8n8k4.f64 (DMMA).  The swizzle makes a
// fragment's 8 rows conflict-free only when they are 2 apart, so fragment i of a warp
// covers rows {16*(i/2) + 2g + (i&1)}: the same map is applied to A rows, B rows (= C
// columns) and the accumulator addresses, so the product is unchanged.
//
// Roofline: FP64 tensor pipe.  Per tile and 16-deep K step: 2*128*128*16 = 524k flop against
// 32 KB of TMA traffic (16 flop/B from L2; panels are read ~once from HBM per launch).
// Measured DMMA issue ceiling from registers: 36.4 TFLOP/s (tools/probes/dmma_pred_probe.cu); cuBLAS DGEMM 35.4;
// this kernel 35.5 at 4096^3 and 28.6 (executed useful flops: 0.81 of the sustained DGEMM rate) over a whole C4 iteration
// with the planner's per-launch choice of tile shape (25.8 with 128 x 128 tiles only).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "dev.h"
#include "kernels.h"

namespace gadj {
namespace {

constexpr int STAGES = 4;
// geometry of one tile shape: TM x TN output tile, CW consumer warps as 2 (along M) x CW/2 (along N)
template <int TM, int TN, int CW>
struct TileCfg {
    static constexpr int A_TILE_BYTES = TM * TILE_K * 8;
    static constexpr int B_TILE_BYTES = TN * TILE_K * 8;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static constexpr int THREADS = (CW + 1) * 32;
    static constexpr int WTM = TM / 2, WTN = TN / (CW / 2);   // warp tile
    static constexpr int FI = WTM / 8, FJ = WTN / 8;          // 8-row fragments of A / B per warp
    static_assert(TM % 32 == 0 && TN % 32 == 0 && FI % 2 == 0 && FJ % 2 == 0, "fragment pairs");
    static_assert(TM * TN * 8 <= STAGES * STAGE_BYTES, "the scatter epilogue parks the tile in the ring");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// FP64 add into global memory, result not needed.  (atomicAdd on a pointer read from a descriptor is a generic-space
// atomic: the compiler emits a shared / global space check and both code paths around every one of the up to 4 096 adds
// of a scatter tile; the targets are always panels in global memory.)
__device__ __forceinline__ void red_add_f64(double* p, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// K range (in 16-deep steps) of the tile at (row0, col0): triangular operands carry explicit zeros outside it
__device__ __forceinline__ void tile_k_range(int flags, int row0, int col0, int K, int tm_rows, int& kc0, int& nk)
{
    int k_lo = 0, k_hi = K;
    if (flags & GEMM_KLO_ROW)
        k_lo = row0;
    if (flags & GEMM_KLO_MAX)
        k_lo = row0 > col0 ? row0 : col0;
    if (flags & GEMM_KHI_ROW)
        k_hi = (row0 + tm_rows < K) ? row0 + tm_rows : K;
    kc0 = k_lo / TILE_K;
    nk = k_hi > k_lo ? (k_hi + TILE_K - 1) / TILE_K - kc0 : 0;
}

// Persistent CTAs: the grid is MINB CTAs per SM (or fewer), each CTA strides through the launch's tile list.  The
// mbarrier ring runs on across tiles, so while the consumer warps store one tile the producer warp is already
// fetching the first stages of the next one; the per-tile cost is the epilogue, not a CTA launch + pipeline fill.
// LOADER 0: TMA producer warp + mbarrier ring (the product path).
// LOADER 1: debug aid (GADJ_GEMM_LOADER=ldg) — the consumers fill one stage themselves with plain loads
//           into the same swizzled layout; isolates tensor-map problems from fragment-layout problems.
template <int LOADER, int TM, int TN, int CW, int MINB>
__global__ void __launch_bounds__((CW + 1) * 32, MINB)
    gemm_tile_kernel(const GemmOp* __restrict__ ops, const GemmTile* __restrict__ tiles, int ntiles)
{
    using Cfg = TileCfg<TM, TN, CW>;
    constexpr int A_TILE_BYTES = Cfg::A_TILE_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int FI = Cfg::FI, FJ = Cfg::FJ;
    constexpr int CTHREADS = CW * 32;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base + STAGES * STAGE_BYTES;
    const uint32_t bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_tile = bar_empty + 8 * STAGES;   // scatter tiles: the stages double as the epilogue's staging buffer
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (LOADER == 0 && threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, CW);
        }
        mbar_init(bar_tile, CW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        // ---- TMA producer -------------------------------------------------------
        if (LOADER == 0 && lane == 0) {
            uint32_t gk = 0, nscatter = 0;
            for (int it = blockIdx.x; it < ntiles; it += gridDim.x) {
                const GemmTile tl = tiles[it];
                const GemmOp* op = ops + tl.op;
                const int row0 = tl.tm * TM, col0 = tl.tn * TN;
                const int flags = op->flags;
                int kc0, nk;
                tile_k_range(flags, row0, col0, op->K, TM, kc0, nk);
                const void* tmA = &op->tmA;
                const void* tmB = &op->tmB;
                asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmA) : "memory");
                asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmB) : "memory");
                for (int kc = 0; kc < nk; ++kc, ++gk) {
                    const uint32_t s = gk % STAGES;
                    if (gk >= STAGES)
                        mbar_wait(bar_empty + 8 * s, ((gk / STAGES) - 1) & 1);
                    const uint32_t sa = base + s * STAGE_BYTES;
                    mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
                    tma_load_2d(sa, tmA, (kc0 + kc) * TILE_K, row0, bar_full + 8 * s);
                    tma_load_2d(sa + A_TILE_BYTES, tmB, (kc0 + kc) * TILE_K, col0, bar_full + 8 * s);
                }
                if (flags & GEMM_SCATTER) {
                    // no prefetch past a scatter tile: its epilogue parks the accumulators in the stages
                    mbar_wait(bar_tile, nscatter & 1);
                    ++nscatter;
                }
            }
        }
        return;
    }

    // ---- DMMA consumers ----------------------------------------------------------
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * Cfg::WTM, wn = (warp >> 1) * Cfg::WTN;
    // per-thread constant parts of the swizzled fragment addresses
    // element (r, k) of a rows x 16 tile: r*128 + (((k>>1) ^ (r&7)) << 4) + ((k&1) << 3)
    // fragment i of A sits at row wm + 16*(i>>1) + 2g + (i&1): a per-thread base plus a compile-time offset; the
    // swizzle term depends on (row & 7) = (2g + (i&1)) & 7 only.  Likewise B.
    const uint32_t a_base = (uint32_t)(wm + 2 * g) * 128u, b_base = (uint32_t)(wn + 2 * g) * 128u;
    const int x0 = (2 * g) & 7, x1 = (2 * g + 1) & 7;
    const uint32_t khalf = (uint32_t)(t & 1) << 3;
    const int kq = t >> 1;

    uint32_t gk = 0;
    GemmTile next = tiles[blockIdx.x];
    for (int it = blockIdx.x; it < ntiles; it += gridDim.x) {
    const GemmTile tl = next;
    if (it + (int)gridDim.x < ntiles)
        next = tiles[it + gridDim.x];   // in flight during this tile
    const GemmOp* op = ops + tl.op;
    const int row0 = tl.tm * TM, col0 = tl.tn * TN;
    const int M = op->M, N = op->N, K = op->K;
    const int flags = op->flags, tri_off = op->tri_off;
    int kc0, nk;
    tile_k_range(flags, row0, col0, K, TM, kc0, nk);

    double acc[FI][FJ][2];
#pragma unroll
    for (int i = 0; i < FI; ++i)
#pragma unroll
        for (int j = 0; j < FJ; ++j)
            acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kc = 0; kc < nk; ++kc, ++gk) {
        const uint32_t s = LOADER == 0 ? gk % STAGES : 0;
        const uint32_t sa = base + s * STAGE_BYTES;
        const uint32_t sb = sa + A_TILE_BYTES;
        if (LOADER == 0) {
            mbar_wait(bar_full + 8 * s, (gk / STAGES) & 1);
        } else {
            asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory");
            const double* __restrict__ gA = op->A;
            const double* __restrict__ gB = op->B;
            for (int idx = threadIdx.x; idx < (TM + TN) * TILE_K; idx += CTHREADS) {
                const bool isb = idx >= TM * TILE_K;
                const int id = isb ? idx - TM * TILE_K : idx;
                const int r = id / TILE_K, k = id - r * TILE_K;
                const int gkk = (kc0 + kc) * TILE_K + k;
                const uint32_t off = (uint32_t)r * 128u + ((uint32_t)((k >> 1) ^ (r & 7)) << 4) + ((uint32_t)(k & 1) << 3);
                double v;
                if (!isb)
                    v = (row0 + r < M && gkk < K) ? gA[(int64_t)(row0 + r) * op->lda + gkk] : 0.0;
                else
                    v = (col0 + r < N && gkk < K) ? gB[(int64_t)(col0 + r) * op->ldb + gkk] : 0.0;
                asm volatile("st.shared.f64 [%0], %1;" ::"r"((isb ? sb : sa) + off), "d"(v) : "memory");
            }
            asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory");
        }
#pragma unroll
        for (int k4 = 0; k4 < TILE_K / 4; ++k4) {
            const int chunk = 2 * k4 + kq;
            const uint32_t o0 = ((uint32_t)(chunk ^ x0) << 4) + khalf;
            const uint32_t o1 = ((uint32_t)(chunk ^ x1) << 4) + khalf + 128u;
            const uint32_t pa0 = sa + a_base + o0, pa1 = sa + a_base + o1;
            const uint32_t pb0 = sb + b_base + o0, pb1 = sb + b_base + o1;
            double a[FI], b[FJ];
#pragma unroll
            for (int i = 0; i < FI; ++i)
                a[i] = lds_f64(((i & 1) ? pa1 : pa0) + (uint32_t)(i >> 1) * 2048u);
#pragma unroll
            for (int j = 0; j < FJ; ++j)
                b[j] = lds_f64(((j & 1) ? pb1 : pb0) + (uint32_t)(j >> 1) * 2048u);
#pragma unroll
            for (int i = 0; i < FI; ++i)
#pragma unroll
                for (int j = 0; j < FJ; ++j)
                    dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (LOADER == 0) {
            // Hand the stage back to the TMA producer.  mbarrier.arrive does not wait for this warp's outstanding
            // ld.shared (ptxas schedules the arrive right behind the last LDS, ahead of the DMMAs that consume it, and the
            // SYNCS unit can overtake a load still queued in a busy LSU — seen once in a few hundred adjustments as a
            // 16 x 32 patch of a product computed from the *next* tile's operands).  The proxy fence orders the generic
            // reads before the async-proxy refill of the stage.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                mbar_arrive(bar_empty + 8 * s);
        }
    }

    // ---- epilogue ------------------------------------------------------------------
    const double alpha = (flags & GEMM_NEG) ? -1.0 : 1.0;
    double* __restrict__ C = op->C;
    const int64_t ldc = op->ldc;
    const bool lower = (flags & GEMM_LOWER) != 0;
    if (flags & GEMM_SCATTER) {
        // Scatter through the station-level row map with FP64 RED.ADD.  The accumulators are first parked in the
        // (now idle) pipeline stages as a TM x TN row-major tile, XOR-swizzled in groups of four columns so that
        // the row-order loads are bank-conflict free; the atomics are then issued row by row with the 32 lanes on
        // 32 consecutive source columns (consecutive boundary stations are mostly consecutive in the ancestor, so
        // a warp instruction touches far fewer 32-byte sectors than in fragment order) and the column part of
        // the map is looked up once per lane instead of once per element.
        constexpr int CC = TN / 32;
        asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory");   // every consumer has finished reading the last stages
#pragma unroll
        for (int i = 0; i < FI; ++i) {
            const int rr = wm + 16 * (i >> 1) + 2 * g + (i & 1);
            const uint32_t rbase = base + (uint32_t)rr * (TN * 8);
            const int sw = (rr >> 1) & 7;
#pragma unroll
            for (int p = 0; p < FJ / 2; ++p) {
                const int cg = ((wn + 16 * p) >> 2) + t;            // group of four columns owned by this thread
                const uint32_t addr = rbase + (uint32_t)((cg ^ sw) << 5);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(acc[i][2 * p][0]), "d"(acc[i][2 * p + 1][0])
                             : "memory");
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr + 16), "d"(acc[i][2 * p][1]),
                             "d"(acc[i][2 * p + 1][1])
                             : "memory");
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory");
        // per lane: its columns' targets (ancestor panel, pitch, station row map) and column positions
        double* cbase[CC];
        const int32_t* rmap[CC];
        int64_t cld[CC];
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
            const int c = col0 + 32 * cc + lane;
            cbase[cc] = nullptr;
            if (c < N) {
                const ScatterTarget tg = op->tgt[op->coltgt[c / 3]];
                rmap[cc] = tg.rowmap - tg.jb;                                    // indexed by boundary station
                cld[cc] = tg.ldc;
                cbase[cc] = tg.C + 3ll * rmap[cc][c / 3] + c % 3;
            }
        }
        for (int rr = warp; rr < TM; rr += CW) {
            const int r = row0 + rr;
            if (r >= M)
                break;
            const int si = r / 3, rc = r - 3 * si;
            const uint32_t rbase = base + (uint32_t)rr * (TN * 8);
            const int sw = (rr >> 1) & 7;
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                const int cl = 32 * cc + lane;
                if (cbase[cc] == nullptr || (lower && r + tri_off < col0 + cl))
                    continue;
                const double v = lds_f64(rbase + (uint32_t)((((cl >> 2) ^ sw) << 5) + ((cl & 3) << 3)));
                red_add_f64(cbase[cc] + (3ll * rmap[cc][si] + rc) * cld[cc], alpha * v);
            }
        }
        if (LOADER == 0) {
            // the staging area (written and read through the generic proxy) goes back to the async proxy: same fence
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                mbar_arrive(bar_tile);   // this warp no longer reads the staging area: the producer may refill the stages
        }
    } else {
        const bool accum = (flags & GEMM_ACCUM) != 0;
        // fragments j = 2p and 2p+1 interleave: together a thread owns 4 consecutive columns
        // c0 .. c0+3 = {acc[i][2p][0], acc[i][2p+1][0], acc[i][2p][1], acc[i][2p+1][1]} -> two 16-byte accesses
#pragma unroll
        for (int i = 0; i < FI; ++i) {
            const int r = row0 + wm + 16 * (i >> 1) + 2 * g + (i & 1);
            if (r >= M)
                continue;
            double* __restrict__ crow = C + (int64_t)r * ldc;
#pragma unroll
            for (int p = 0; p < FJ / 2; ++p) {
                const int c0 = col0 + wn + 16 * p + 4 * t;
                double v[4] = {alpha * acc[i][2 * p][0], alpha * acc[i][2 * p + 1][0], alpha * acc[i][2 * p][1],
                               alpha * acc[i][2 * p + 1][1]};
                if (flags & GEMM_DUAL) {
                    // transposed copy (e.g. Z21^T next to Z21): lanes of a quad write four neighbouring rows of Ct, the
                    // eight quads of the warp eight elements 2 apart along a row; the (i&1) partner instruction fills the gaps
                    double* __restrict__ ct = op->Ct + (int64_t)c0 * op->ldct + r;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (c0 + e < N)
                            ct[(int64_t)e * op->ldct] = v[e];
                }
                if (c0 + 3 < N && (!lower || r + tri_off >= c0 + 3)) {
                    double2* q = reinterpret_cast<double2*>(crow + c0);
                    if (accum) {
                        const double2 o0 = q[0], o1 = q[1];
                        v[0] += o0.x;
                        v[1] += o0.y;
                        v[2] += o1.x;
                        v[3] += o1.y;
                    }
                    q[0] = make_double2(v[0], v[1]);
                    q[1] = make_double2(v[2], v[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = c0 + e;
                        if (c >= N || (lower && r + tri_off < c))
                            continue;
                        crow[c] = accum ? crow[c] + v[e] : v[e];
                    }
                }
            }
        }
    }
    if (LOADER != 0)
        asm volatile("bar.sync 1, %0;" ::"n"(CTHREADS) : "memory");   // debug loader: the next tile's self-loads reuse stage 0
    }   // tile loop
}

template <int LOADER, int TM, int TN, int CW, int MINB>
void launch_shape(const GemmOp* ops, const GemmTile* tiles, int ntiles, int first_use_key, cudaStream_t st)
{
    using Cfg = TileCfg<TM, TN, CW>;
    auto* k = gemm_tile_kernel<LOADER, TM, TN, CW, MINB>;
    if (dev::first_use(first_use_key))   // kernel attributes are per device
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    const int cap = dev::sm_count() * MINB;   // persistent CTAs: as many as are resident at once
    const int grid = ntiles < cap ? ntiles : cap;
    k<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(ops, tiles, ntiles);
}

}  // namespace

void launch_gemm(const GemmOp* ops, int nops, const GemmTile* tiles, int ntiles, int shape, void* stream)
{
    if (nops <= 0 || ntiles <= 0)
        return;
    static const int loader = [] {
        const char* e = getenv("GADJ_GEMM_LOADER");
        return (e && e[0] == 'l') ? 1 : 0;
    }();
    cudaStream_t st = (cudaStream_t)stream;
    if (shape == TILE_SHAPE_64) {
        if (loader == 0)
            launch_shape<0, 64, 64, 4, 3>(ops, tiles, ntiles, KEY_GEMM64, st);
        else
            launch_shape<1, 64, 64, 4, 3>(ops, tiles, ntiles, KEY_GEMM64_LDG, st);
    } else {
        if (loader == 0)
            launch_shape<0, TILE_M, TILE_N, 8, 1>(ops, tiles, ntiles, KEY_GEMM, st);
        else
            launch_shape<1, TILE_M, TILE_N, 8, 1>(ops, tiles, ntiles, KEY_GEMM_LDG, st);
    }
}

}  // namespace gadj
