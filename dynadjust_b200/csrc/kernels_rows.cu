// kernels_rows.cu — device passes over the design rows of every measurement type other than the single GNSS
// baseline (which has its own streaming kernel in kernels_assemble.cu): terrestrial rows, derived angles of
// direction sets, and the rows of GNSS baseline / point clusters.  The arithmetic lives in rows.h (shared with the
// CPU stand-in build); the kernels here only distribute rows / clusters over threads.
//
//   rows_kernel            one thread per row: l, partials -> row arrays; independent rows scatter p a^T a, p a^T l
//   cluster_kernel         one CTA per cluster: t = V^-1 l; w += A^T t; N += A^T V^-1 A over the cluster's station pairs
//   cluster_inverse_kernel one CTA per cluster: V -> V^-1 (Cholesky, triangular inverse, W^T W) — FormInverseVarianceMatrix
//   rows_stats_kernel      one thread per row: precision of the adjusted measurement, record statistics, chi-square
//   cluster_chi_kernel     one CTA per X / Y cluster: l^T V^-1 l
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"
#include "rows.h"

namespace gadj {
namespace {

inline int grid_for(uint64_t n, int block, int max_blocks = 148 * 16)
{
    uint64_t g = (n + block - 1) / block;
    if (g < 1)
        g = 1;
    return (int)(g > (uint64_t)max_blocks ? max_blocks : g);
}

__global__ void __launch_bounds__(128) rows_kernel(const RowsParams p)
{
    const Ellipsoid el = make_ellipsoid(p.semi_major, p.inv_flattening);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nrows; i += (uint64_t)gridDim.x * blockDim.x)
        row_body(p, i, el);
}

__global__ void __launch_bounds__(128) rows_stats_kernel(const RowsParams p)
{
    const Ellipsoid el = make_ellipsoid(p.semi_major, p.inv_flattening);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nrows; i += (uint64_t)gridDim.x * blockDim.x)
        row_stats_body(p, i, el, acc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if ((threadIdx.x & 31) == 0 && acc[k] != 0.0)
            atomicAdd(p.sums + k, acc[k]);
    }
}

__global__ void __launch_bounds__(128) cluster_kernel(const ClusterParams p)
{
    for (uint32_t ci = blockIdx.x; ci < p.nclusters; ci += gridDim.x) {
        const ClusterDesc c = p.clusters[ci];
        for (uint32_t r = threadIdx.x; r < c.n; r += blockDim.x)
            cluster_t_body(p, c, r);
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < c.ns; j += blockDim.x)
            cluster_rhs_body(p, c, j);
        if (p.normals) {
            const uint64_t npairs = (uint64_t)c.ns * (c.ns + 1) / 2;
            for (uint64_t q = threadIdx.x; q < npairs; q += blockDim.x)
                cluster_pair_body(p, c, q);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128) cluster_chi_kernel(const ClusterParams p)
{
    double acc = 0.0;
    for (uint32_t ci = blockIdx.x; ci < p.nclusters; ci += gridDim.x) {
        const ClusterDesc c = p.clusters[ci];
        if (c.type == 'D')
            continue;   // direction sets: diagonal form, summed per row (ComputeChiSquare_D, ADJ:8440-8469)
        for (uint32_t r = threadIdx.x; r < c.n; r += blockDim.x) {
            cluster_t_body(p, c, r);
            acc += p.row_l[c.row0 + r] * p.row_t[c.row0 + r];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0.0)
        atomicAdd(p.sums, acc);
}

// V (n x n, row-major, full symmetric, in `a`) -> V^-1 in `out`.  One CTA per cluster; the matrix stays in global
// memory (L2-resident for all but the largest clusters).  info: first failing cluster index + 1.
__global__ void __launch_bounds__(256) cluster_inverse_kernel(const ClusterDesc* __restrict__ clusters, uint32_t nclusters,
                                                              double* __restrict__ work, double* __restrict__ out,
                                                              int* __restrict__ info)
{
    __shared__ double pivot;
    __shared__ int bad;
    for (uint32_t ci = blockIdx.x; ci < nclusters; ci += gridDim.x) {
        const ClusterDesc c = clusters[ci];
        const uint32_t n = c.n;
        double* A = work + c.vinv_off;
        double* O = out + c.vinv_off;
        if (threadIdx.x == 0)
            bad = 0;
        __syncthreads();
        // 1. right-looking Cholesky of the lower triangle, in place
        for (uint32_t j = 0; j < n; ++j) {
            if (threadIdx.x == 0) {
                double d = A[(size_t)j * n + j];
                if (!(d > 0.0)) {
                    bad = 1;
                    d = 1.0;
                }
                pivot = sqrt(d);
                A[(size_t)j * n + j] = pivot;
            }
            __syncthreads();
            const double inv = 1.0 / pivot;
            for (uint32_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x)
                A[(size_t)i * n + j] *= inv;
            __syncthreads();
            // trailing update: row i, columns j+1..i
            for (uint32_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x) {
                const double lij = A[(size_t)i * n + j];
                double* row = A + (size_t)i * n;
                for (uint32_t k = j + 1; k <= i; ++k)
                    row[k] -= lij * A[(size_t)k * n + j];
            }
            __syncthreads();
        }
        // 2. W = L^-1 column by column; W[i][c] (i > c) is stored at A[c][i] (the unused upper triangle)
        for (uint32_t col = threadIdx.x; col < n; col += blockDim.x) {
            const double xc = 1.0 / A[(size_t)col * n + col];
            for (uint32_t i = col + 1; i < n; ++i) {
                double s = A[(size_t)i * n + col] * xc;
                for (uint32_t k = col + 1; k < i; ++k)
                    s += A[(size_t)i * n + k] * A[(size_t)col * n + k];
                A[(size_t)col * n + i] = -s / A[(size_t)i * n + i];
            }
        }
        __syncthreads();
        // 3. V^-1 = W^T W:  out[i][j] = sum_{k >= i} W[k][i] W[k][j],  i >= j
        const uint64_t total = (uint64_t)n * (n + 1) / 2;
        for (uint64_t e = threadIdx.x; e < total; e += blockDim.x) {
            uint32_t i = (uint32_t)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
            while ((uint64_t)i * (i + 1) / 2 > e)
                --i;
            while ((uint64_t)(i + 1) * (i + 2) / 2 <= e)
                ++i;
            const uint32_t j = (uint32_t)(e - (uint64_t)i * (i + 1) / 2);
            const double wii = 1.0 / A[(size_t)i * n + i];
            double s = wii * (i == j ? wii : A[(size_t)j * n + i]);
            for (uint32_t k = i + 1; k < n; ++k)
                s += A[(size_t)i * n + k] * A[(size_t)j * n + k];
            O[(size_t)i * n + j] = s;
            O[(size_t)j * n + i] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0 && bad)
            atomicCAS(info, 0, (int)ci + 1);
        __syncthreads();
    }
}

}  // namespace

void launch_rows(const RowsParams& p, void* stream)
{
    if (p.nrows == 0)
        return;
    rows_kernel<<<grid_for(p.nrows, 128), 128, 0, (cudaStream_t)stream>>>(p);
}

void launch_rows_stats(const RowsParams& p, void* stream)
{
    if (p.nrows == 0)
        return;
    rows_stats_kernel<<<grid_for(p.nrows, 128), 128, 0, (cudaStream_t)stream>>>(p);
}

void launch_clusters(const ClusterParams& p, void* stream)
{
    if (p.nclusters == 0)
        return;
    cluster_kernel<<<(int)(p.nclusters < 148u * 16u ? p.nclusters : 148u * 16u), 128, 0, (cudaStream_t)stream>>>(p);
}

void launch_cluster_chi(const ClusterParams& p, void* stream)
{
    if (p.nclusters == 0)
        return;
    cluster_chi_kernel<<<(int)(p.nclusters < 148u * 16u ? p.nclusters : 148u * 16u), 128, 0, (cudaStream_t)stream>>>(p);
}

void launch_cluster_inverse(const ClusterDesc* clusters, uint32_t nclusters, double* work, double* out, int* info, void* stream)
{
    if (nclusters == 0)
        return;
    cluster_inverse_kernel<<<(int)(nclusters < 148u * 8u ? nclusters : 148u * 8u), 256, 0, (cudaStream_t)stream>>>(clusters, nclusters,
                                                                                                                 work, out, info);
}

}  // namespace gadj
