// kernels_assemble.cu — the HBM-bound passes over the measurement list and the station arrays:
//   normal-equation assembly (N = A^T V^-1 A, w = A^T V^-1 l), equilibration + scatter of N into the
//   front panels, right-hand-side permutation, estimate update, VCV extraction and statistics.
//
// Assembly follows UpdateDesignNormalMeasMatrices_G / LoadVarianceMatrix_G / UpdateNormals_G
// (ADJ:5353-5397, ADJ:4214-4309, ADJ:1664-1684) in two passes without floating-point atomics on the diagonal:
//   pass 1, one GNSS baseline per thread:  l = term1 - (X2 - X1);  V^-1 from the six variance terms of the three
//           records;  V^-1 and V^-1 l are written to a 72-byte slot per baseline and  N[s2,s1] = -V^-1  goes straight
//           to the baseline's off-diagonal block (plain store when no other measurement shares the station pair).
//           The raw 208-byte records are streamed into shared memory by 1-D bulk async copies (cp.async.bulk +
//           mbarrier, i.e. the TMA engine, L2 evict-first) one 64-baseline tile (39,936 B) at a time; several CTAs
//           per SM overlap one CTA's copy with the others' arithmetic.
//   pass 2, 16 lanes per station:  N[s,s] += sum of V^-1 over the station's incidence list,  w[s] += sum of -/+ V^-1 l
//           — a gather in a fixed order, so the assembled normals are bit-reproducible from run to run.
// Algorithmic traffic per baseline: 3*208 B records + 8 B plan words + 48 B station XYZ
// + 27*8 B block updates + 48 B rhs updates = 944 B  (936 B in SURVEY.md §8d + the two plan words).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "geodesy.h"
#include "dev.h"
#include "kernels.h"

namespace gadj {
namespace {

constexpr int ASM_TILE = 64;                 // baselines per tile = threads per CTA
constexpr int ASM_BYTES_PER_BSL = 3 * 208;   // three records
constexpr int ASM_SMEM = ASM_TILE * ASM_BYTES_PER_BSL;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// the records are read exactly once: evict-first keeps the station arrays and the 72-byte slots in L2 instead
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

// arithmetic of one baseline given its three records
__device__ __forceinline__ void baseline_contribution(const dna_msr_t* __restrict__ m, const AssembleParams& p, uint64_t b,
                                                      uint32_t edge_word)
{
    const uint32_t s1 = m[0].station1, s2 = m[0].station2;
    const double* __restrict__ e1 = p.est + 3 * (size_t)s1;
    const double* __restrict__ e2 = p.est + 3 * (size_t)s2;
    double l[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        l[r] = m[r].term1 - (e2[r] - e1[r]);
    const double up[6] = {m[0].term2, m[1].term2, m[2].term2, m[1].term3, m[2].term3, m[2].term4};
    double q[6];
    if (!spd3_inverse(up, q)) {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            q[k] = __longlong_as_double(0x7ff8000000000000ll);
    }
    const double V[9] = {q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5]};
    double* __restrict__ slot = p.bq + 9 * b;
#pragma unroll
    for (int r = 0; r < 3; ++r)
        slot[6 + r] = V[3 * r] * l[0] + V[3 * r + 1] * l[1] + V[3 * r + 2] * l[2];
    if (p.normals) {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            slot[k] = q[k];
        if (!(edge_word & EDGE_EXCLUSIVE)) {     // a station pair shared with other measurements: accumulate its block
            double* __restrict__ o = p.noff + 9 * (size_t)(edge_word & EDGE_SLOT_MASK);
#pragma unroll
            for (int k = 0; k < 9; ++k)
                atomicAdd(o + k, -V[k]);
        }
    }
}

template <bool STAGED>
__global__ void __launch_bounds__(ASM_TILE) assemble_g_kernel(const AssembleParams p, uint64_t ntiles)
{
    extern __shared__ __align__(128) uint8_t stage[];
    __shared__ __align__(8) uint64_t bar_storage;
    const int tid = threadIdx.x;
    const uint32_t bar = smem_u32(&bar_storage);
    if (STAGED) {
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    uint32_t phase = 0;
    const uint64_t pol = policy_evict_first();
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t b0 = tile * ASM_TILE;
        const int nb = (int)((p.nbaselines - b0 < (uint64_t)ASM_TILE) ? (p.nbaselines - b0) : ASM_TILE);
        const uint64_t b = b0 + tid;
        const bool active = tid < nb;
        const uint32_t first = active ? p.first[b] : 0u;
        const uint32_t ew = (active && p.normals) ? p.edge[b] : 0u;
        const dna_msr_t* m;
        if (STAGED) {
            if (p.contiguous) {
                if (tid == 0) {
                    const uint32_t bytes = (uint32_t)nb * ASM_BYTES_PER_BSL;
                    mbar_expect_tx(bar, bytes);
                    bulk_g2s(smem_u32(stage), p.msr + first, bytes, bar, pol);
                }
            } else {
                if (tid == 0)
                    mbar_expect_tx(bar, (uint32_t)nb * ASM_BYTES_PER_BSL);
                __syncthreads();  // the expectation is posted before any copy can complete
                if (active)
                    bulk_g2s(smem_u32(stage) + tid * ASM_BYTES_PER_BSL, p.msr + first, ASM_BYTES_PER_BSL, bar, pol);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            m = reinterpret_cast<const dna_msr_t*>(stage + tid * ASM_BYTES_PER_BSL);
        } else {
            m = p.msr + first;
        }
        if (active)
            baseline_contribution(m, p, b, ew);
        if (STAGED) {
            // everyone is done with the stage before the next bulk copy (async proxy) overwrites it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
    }
}

// pass 2: 16 lanes per station (12 used: the 9 elements of the diagonal block and the 3 of the right-hand side);
// every lane walks the station's incidence list and adds its element of each baseline's slot
__global__ void __launch_bounds__(256) station_sum_kernel(const AssembleParams p)
{
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t s = gid >> 4;
    const int k = (int)(gid & 15);
    if (s >= p.nstn || k >= 12 || (k < 9 && !p.normals))
        return;
    // element k of the 3x3 block -> index into the stored upper triangle {00 01 02 11 12 22}
    const int sym = k < 9 ? GADJ_SYM3(k) : k - 3;
    const uint32_t i0 = p.inc_ptr[s], i1 = p.inc_ptr[s + 1];
    double acc = 0.0;
    // four incidences per round: the index loads, then the four slot loads, are in flight together; the sum order
    // (ascending list position) is the same for every run
    const bool rhs = k >= 9;
    uint32_t i = i0;
    for (; i + 4 <= i1; i += 4) {
        uint32_t e[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            e[u] = p.inc[i + u];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            v[u] = p.bq[9ull * (e[u] & 0x7FFFFFFFu) + sym];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            acc += (rhs && !(e[u] & 0x80000000u)) ? -v[u] : v[u];      // w[s1] -= V^-1 l,  w[s2] += V^-1 l
    }
    for (; i < i1; ++i) {
        const uint32_t e = p.inc[i];
        const double v = p.bq[9ull * (e & 0x7FFFFFFFu) + sym];
        acc += (rhs && !(e & 0x80000000u)) ? -v : v;
    }
    if (rhs)
        p.w[3 * s + (k - 9)] += acc;
    else
        p.ndiag[9 * s + k] += acc;
}

__global__ void init_normals_kernel(const double* __restrict__ cblock, double* __restrict__ ndiag, double* __restrict__ noff,
                                    double* __restrict__ w, uint64_t n_diag, uint64_t n_off, uint64_t n_w)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cblock) {
        for (uint64_t i = i0; i < n_diag; i += stride)
            ndiag[i] = cblock[i];
        for (uint64_t i = i0; i < n_off; i += stride)
            noff[i] = 0.0;
    }
    for (uint64_t i = i0; i < n_w; i += stride)
        w[i] = 0.0;
}

__global__ void compute_scale_kernel(const ScatterParams p)
{
    const uint64_t n = 3ull * p.nstn;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / 3;
        const int c = (int)(i - 3 * s);
        p.dscale[i] = p.scale ? 1.0 / sqrt(p.ndiag[9 * s + 4 * c]) : 1.0;
    }
}

// one thread per 3x3 block element; diagonal blocks first, then edge blocks
__global__ void scatter_normals_kernel(const ScatterParams p)
{
    const uint64_t nd = 9ull * p.nstn, total = nd + 9ull * p.nedge;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        if (i < nd) {
            const uint64_t s = i / 9;
            const int k = (int)(i - 9 * s), r = k / 3, c = k - 3 * r;
            const uint64_t d = p.diag_dest[s];
            if (d != ~0ull)
                p.panels[d + (uint64_t)r * p.diag_ld[s] + c] = p.ndiag[i] * p.dscale[3 * s + r] * p.dscale[3 * s + c];
        } else {
            const uint64_t j = i - nd, e = j / 9;
            const int k = (int)(j - 9 * e), r = k / 3, c = k - 3 * r;
            const uint64_t d = p.off_dest[e];
            if (d != ~0ull) {
                const uint32_t b = p.edge_bsl[e];
                const double v = b != ~0u ? -p.bq[9ull * b + GADJ_SYM3(k)] : p.noff[j];
                p.panels[d + (uint64_t)r * p.off_ld[e] + c] = v * p.dscale[3ull * p.edge_hi[e] + r] * p.dscale[3ull * p.edge_lo[e] + c];
            }
        }
    }
}

__global__ void permute_rhs_kernel(const double* __restrict__ w, const double* __restrict__ dscale,
                                   const uint32_t* __restrict__ pos, const uint8_t* __restrict__ pos_owned,
                                   double* __restrict__ b, uint32_t nstn)
{
    const uint64_t n = 3ull * nstn;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / 3;
        const uint32_t p = pos[s];
        b[3ull * p + (i - 3 * s)] = (pos_owned == nullptr || pos_owned[p]) ? dscale[i] * w[i] : 0.0;
    }
}

__global__ void mask_positions_kernel(double* __restrict__ x, const uint8_t* __restrict__ pos_owned, uint32_t nstn)
{
    const uint64_t n = 3ull * nstn;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (!pos_owned[i / 3])
            x[i] = 0.0;
}

// corrections + estimates + per-block (|max|, first index) partials
__global__ void __launch_bounds__(256) apply_corrections_kernel(const double* __restrict__ x, const double* __restrict__ dscale,
                                                                const uint32_t* __restrict__ pos, double* __restrict__ corr,
                                                                double* __restrict__ est, uint32_t nstn,
                                                                double* __restrict__ part_val, unsigned long long* __restrict__ part_idx)
{
    const uint64_t n = 3ull * nstn;
    double best = 0.0;
    unsigned long long besti = ~0ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / 3;
        const double c = dscale[i] * x[3ull * pos[s] + (i - 3 * s)];
        corr[i] = c;
        est[i] += c;
        if (besti == ~0ull || fabs(c) > fabs(best)) {
            best = c;
            besti = i;
        }
    }
    __shared__ double sv[256];
    __shared__ unsigned long long si[256];
    sv[threadIdx.x] = best;
    si[threadIdx.x] = besti;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v2 = sv[threadIdx.x + o];
            const unsigned long long i2 = si[threadIdx.x + o];
            const double v1 = sv[threadIdx.x];
            const unsigned long long i1 = si[threadIdx.x];
            const bool take = (i2 != ~0ull) && (i1 == ~0ull || fabs(v2) > fabs(v1) || (fabs(v2) == fabs(v1) && i2 < i1));
            if (take) {
                sv[threadIdx.x] = v2;
                si[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part_val[blockIdx.x] = sv[0];
        part_idx[blockIdx.x] = si[0];
    }
}

__global__ void __launch_bounds__(256) finish_max_kernel(const double* __restrict__ part_val,
                                                         const unsigned long long* __restrict__ part_idx, int nparts,
                                                         const double* __restrict__ corr, double* __restrict__ tail)
{
    __shared__ double sv[256];
    __shared__ unsigned long long si[256];
    double best = 0.0;
    unsigned long long besti = ~0ull;
    for (int i = threadIdx.x; i < nparts; i += 256) {
        const double v = part_val[i];
        const unsigned long long id = part_idx[i];
        if (id != ~0ull && (besti == ~0ull || fabs(v) > fabs(best) || (fabs(v) == fabs(best) && id < besti))) {
            best = v;
            besti = id;
        }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = besti;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v2 = sv[threadIdx.x + o];
            const unsigned long long i2 = si[threadIdx.x + o];
            const double v1 = sv[threadIdx.x];
            const unsigned long long i1 = si[threadIdx.x];
            const bool take = (i2 != ~0ull) && (i1 == ~0ull || fabs(v2) > fabs(v1) || (fabs(v2) == fabs(v1) && i2 < i1));
            if (take) {
                sv[threadIdx.x] = v2;
                si[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tail[0] = sv[0];
        tail[1] = (double)si[0];
        const unsigned long long s3 = (si[0] / 3) * 3;   // the station's three Cartesian corrections
        tail[2] = corr[s3];
        tail[3] = corr[s3 + 1];
        tail[4] = corr[s3 + 2];
    }
}

__global__ void extract_station_vcv_kernel(const double* __restrict__ panels, const uint64_t* __restrict__ diag_dest,
                                           const uint32_t* __restrict__ diag_ld, const double* __restrict__ dscale,
                                           double* __restrict__ vcv, uint32_t nstn)
{
    const uint64_t n = 9ull * nstn;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t s = i / 9;
        const int k = (int)(i - 9 * s), r = k / 3, c = k - 3 * r;
        const int a = r > c ? r : c, b = r > c ? c : r;
        const uint64_t d = diag_dest[s];
        vcv[i] = d != ~0ull ? panels[d + (uint64_t)a * diag_ld[s] + b] * dscale[3 * s + r] * dscale[3 * s + c] : 0.0;
    }
}

__global__ void extract_edge_vcv_kernel(const double* __restrict__ panels, const uint64_t* __restrict__ off_dest,
                                        const uint32_t* __restrict__ off_ld, const uint32_t* __restrict__ edge_hi,
                                        const uint32_t* __restrict__ edge_lo, const double* __restrict__ dscale,
                                        double* __restrict__ q, uint64_t nedge)
{
    const uint64_t n = 9ull * nedge;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = i / 9;
        const int k = (int)(i - 9 * e), r = k / 3, c = k - 3 * r;
        const uint64_t d = off_dest[e];
        q[i] = d != ~0ull ? panels[d + (uint64_t)r * off_ld[e] + c] * dscale[3ull * edge_hi[e] + r] * dscale[3ull * edge_lo[e] + c]
                          : 0.0;
    }
}

// ComputePrecisionAdjMsrs_GX + UpdateMsrRecords_GXY + ComputeChiSquare_G + ComputeGlobalPelzer_GXY
// (ADJ:8006-8032, ADJ:8152-8298, ADJ:8530-8549, ADJ:8396-8427) for one baseline per thread
__global__ void __launch_bounds__(128) stats_g_kernel(const StatsParams p)
{
    double s_chi = 0.0, s_pel = 0.0, s_cnt = 0.0, s_out = 0.0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < p.nbaselines; b += (uint64_t)gridDim.x * blockDim.x) {
        dna_msr_t* m = p.msr + p.first[b];
        const uint32_t s1 = m[0].station1, s2 = m[0].station2;
        double l[3], t1[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            t1[r] = m[r].term1;
            l[r] = t1[r] - (p.est[3 * (size_t)s2 + r] - p.est[3 * (size_t)s1 + r]);
        }
        const uint32_t ew = p.edge[b];
        const double* __restrict__ Qo = p.vcv_off + 9 * (size_t)(ew & EDGE_SLOT_MASK);
        const bool s1_is_hi = (ew & 0x80000000u) != 0;
        const double* __restrict__ Q11 = p.vcv_diag + 9 * (size_t)s1;
        const double* __restrict__ Q22 = p.vcv_diag + 9 * (size_t)s2;
        double prec[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double q21 = Qo[4 * i];  // Q21(i,i): same element either orientation
            const double t0 = -Q11[4 * i] + q21;
            const double tt = -q21 + Q22[4 * i];
            prec[i] = tt - t0;
        }
        (void)s1_is_hi;
        const double up[6] = {m[0].term2, m[1].term2, m[2].term2, m[1].term3, m[2].term3, m[2].term4};
        const double mprec[3] = {up[0], up[3], up[5]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double corr = -l[r];
            double rp = mprec[r] - prec[r];
            if (rp < 0.0)
                rp = fabs(rp);
            double pel = sqrt(mprec[r]) / sqrt(rp);
            if (pel < 0. || pel > 700.)
                pel = 999.99;
            const double nstat = corr / sqrt(rp);
            if (fabs(nstat) > p.critical)
                s_out += 1.0;
            if (pel > 0. && pel < 999.99) {
                s_pel += pel * pel - 1.;
                s_cnt += 1.0;
            } else
                pel = 999.99;
            m[r].measCorr = corr;
            m[r].measAdj = t1[r] + corr;
            m[r].measAdjPrec = prec[r];
            m[r].residualPrec = rp;
            m[r].NStat = nstat;
            m[r].PelzerRel = pel;
        }
        double q[6];
        if (!spd3_inverse(up, q)) {
#pragma unroll
            for (int k = 0; k < 6; ++k)
                q[k] = __longlong_as_double(0x7ff8000000000000ll);
        }
        const double V[9] = {q[0], q[1], q[2], q[1], q[3], q[4], q[2], q[4], q[5]};
        double cs = 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                cs += V[3 * r + c] * l[r] * l[c];
        s_chi += cs;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_chi += __shfl_xor_sync(0xffffffffu, s_chi, o);
        s_pel += __shfl_xor_sync(0xffffffffu, s_pel, o);
        s_cnt += __shfl_xor_sync(0xffffffffu, s_cnt, o);
        s_out += __shfl_xor_sync(0xffffffffu, s_out, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(p.sums + 0, s_chi);
        atomicAdd(p.sums + 1, s_pel);
        atomicAdd(p.sums + 2, s_cnt);
        atomicAdd(p.sums + 3, s_out);
    }
}

__global__ void cart_to_geo_kernel(const double* __restrict__ est, double* __restrict__ llh, uint32_t nstn, double a, double invf)
{
    const Ellipsoid e = make_ellipsoid(a, invf);
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nstn; s += (uint64_t)gridDim.x * blockDim.x)
        cart_to_geo(e, est[3 * s], est[3 * s + 1], est[3 * s + 2], llh + 3 * s);
}

inline int grid_for(uint64_t n, int block, int max_blocks = 148 * 16)
{
    uint64_t g = (n + block - 1) / block;
    if (g < 1)
        g = 1;
    return (int)(g > (uint64_t)max_blocks ? max_blocks : g);
}

constexpr int MAX_PARTS = APPLY_MAX_PARTS;

}  // namespace

void launch_assemble_g(const AssembleParams& p, void* stream)
{
    static const int mode = [] {   // 0 staged (bulk async copies), 1 direct loads (debug aid: GADJ_ASSEMBLE_DIRECT=1)
        const char* e = getenv("GADJ_ASSEMBLE_DIRECT");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    if (dev::first_use(KEY_ASSEMBLE))
        cudaFuncSetAttribute(assemble_g_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ASM_SMEM);
    if (p.nbaselines == 0)
        return;
    const uint64_t ntiles = (p.nbaselines + ASM_TILE - 1) / ASM_TILE;
    const int grid = (int)(ntiles < (uint64_t)(148 * 5) ? ntiles : (uint64_t)(148 * 5));
    if (mode == 0)
        assemble_g_kernel<true><<<grid, ASM_TILE, ASM_SMEM, (cudaStream_t)stream>>>(p, ntiles);
    else
        assemble_g_kernel<false><<<grid * 4, ASM_TILE, 0, (cudaStream_t)stream>>>(p, ntiles);
}

void launch_station_sum(const AssembleParams& p, void* stream)
{
    if (p.nbaselines == 0)
        return;
    const uint64_t threads = 16ull * p.nstn;
    station_sum_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
}

void launch_init_normals(const double* cblock, double* ndiag, double* noff, double* w, uint32_t nstn, uint64_t nedge,
                         void* stream)
{
    const uint64_t nmax = 9ull * (nedge > nstn ? nedge : nstn);
    init_normals_kernel<<<grid_for(nmax, 256), 256, 0, (cudaStream_t)stream>>>(cblock, ndiag, noff, w, 9ull * nstn, 9ull * nedge,
                                                                             3ull * nstn);
}

void launch_compute_scale(const ScatterParams& p, void* stream)
{
    compute_scale_kernel<<<grid_for(3ull * p.nstn, 256), 256, 0, (cudaStream_t)stream>>>(p);
}

void launch_scatter_normals(const ScatterParams& p, void* stream)
{
    scatter_normals_kernel<<<grid_for(9ull * (p.nstn + p.nedge), 256), 256, 0, (cudaStream_t)stream>>>(p);
}

void launch_permute_rhs(const double* w, const double* dscale, const uint32_t* pos_of_stn, const uint8_t* pos_owned, double* b,
                        uint32_t nstn, void* stream)
{
    permute_rhs_kernel<<<grid_for(3ull * nstn, 256), 256, 0, (cudaStream_t)stream>>>(w, dscale, pos_of_stn, pos_owned, b, nstn);
}

void launch_mask_positions(double* x, const uint8_t* pos_owned, uint32_t nstn, void* stream)
{
    mask_positions_kernel<<<grid_for(3ull * nstn, 256), 256, 0, (cudaStream_t)stream>>>(x, pos_owned, nstn);
}

void launch_apply_corrections(const double* x, const double* dscale, const uint32_t* pos_of_stn, double* corr, double* est,
                              uint32_t nstn, double* scratch, void* stream)
{
    double* g_part_val = scratch;
    unsigned long long* g_part_idx = reinterpret_cast<unsigned long long*>(scratch + MAX_PARTS);
    const int grid = grid_for(3ull * nstn, 256, MAX_PARTS);
    apply_corrections_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, dscale, pos_of_stn, corr, est, nstn, g_part_val,
                                                                     g_part_idx);
    finish_max_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(g_part_val, g_part_idx, grid, corr, corr + 3ull * nstn);
}

void launch_extract_station_vcv(const double* panels, const uint64_t* diag_dest, const uint32_t* diag_ld, const double* dscale,
                                double* vcv, uint32_t nstn, void* stream)
{
    extract_station_vcv_kernel<<<grid_for(9ull * nstn, 256), 256, 0, (cudaStream_t)stream>>>(panels, diag_dest, diag_ld, dscale,
                                                                                           vcv, nstn);
}

void launch_extract_edge_vcv(const double* panels, const uint64_t* off_dest, const uint32_t* off_ld, const uint32_t* edge_hi,
                             const uint32_t* edge_lo, const double* dscale, double* q, uint64_t nedge, void* stream)
{
    if (nedge == 0)
        return;
    extract_edge_vcv_kernel<<<grid_for(9ull * nedge, 256), 256, 0, (cudaStream_t)stream>>>(panels, off_dest, off_ld, edge_hi,
                                                                                         edge_lo, dscale, q, nedge);
}

void launch_stats_g(const StatsParams& p, void* stream)
{
    if (p.nbaselines == 0)
        return;
    stats_g_kernel<<<grid_for(p.nbaselines, 128), 128, 0, (cudaStream_t)stream>>>(p);
}

void launch_cart_to_geo(const double* est, double* llh, uint32_t nstn, double a, double invf, void* stream)
{
    cart_to_geo_kernel<<<grid_for(nstn, 128), 128, 0, (cudaStream_t)stream>>>(est, llh, nstn, a, invf);
}

}  // namespace gadj
