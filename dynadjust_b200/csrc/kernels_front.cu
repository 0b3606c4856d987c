// kernels_front.cu — per-front helper kernels around the tile GEMM:
//   pivot-tile Cholesky + triangular inverse, triangular tile solves, panel GEMV updates for the
//   forward/backward substitution, transposes and the selected-inverse gather.
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace gadj {
namespace {

constexpr int DP = NB + 1;  // shared-memory pitch (doubles): conflict-free for row- and column-wise walks

// ---- pivot tile: L = chol(D), W = L^-1 ----------------------------------------------
// One CTA (NB threads) per tile.  Left-looking column Cholesky in shared memory: thread i owns
// row i.  The inverse is formed column-by-column (thread c solves L x = e_c) into the strictly
// upper triangle of the same buffer (x_k stored at S[c][k]), so one 128 x 129 FP64 tile suffices.
__global__ void __launch_bounds__(NB, 1) diag_kernel(const DiagOp* __restrict__ ops, int* __restrict__ info)
{
    extern __shared__ double S[];
    __shared__ double dinv[NB];
    const DiagOp op = ops[blockIdx.x];
    const int w = op.w, tid = threadIdx.x;
    const int64_t ld = op.ldd;
    // coalesced load of the lower triangle (row-major source)
    for (int idx = tid; idx < w * w; idx += NB) {
        int i = idx / w, j = idx - i * w;
        S[i * DP + j] = (j <= i) ? op.D[i * ld + j] : 0.0;
    }
    __syncthreads();
    if (op.factor) {
        for (int j = 0; j < w; ++j) {
            double s = 0.0;
            if (tid >= j && tid < w) {
                s = S[tid * DP + j];
                for (int k = 0; k < j; ++k)
                    s -= S[tid * DP + k] * S[j * DP + k];
            }
            if (tid == j) {
                if (!(s > 0.0)) {
                    atomicCAS(info, 0, op.front + 1);
                    s = 1.0;
                }
                s = sqrt(s);
                S[j * DP + j] = s;
            }
            __syncthreads();
            if (tid > j && tid < w)
                S[tid * DP + j] = s / S[j * DP + j];
            __syncthreads();
        }
        for (int idx = tid; idx < w * w; idx += NB) {
            int i = idx / w, j = idx - i * w;
            if (j <= i)
                op.D[i * ld + j] = S[i * DP + j];
        }
    }
    if (op.W == nullptr && op.Wt == nullptr)
        return;
    // W = L^-1: thread c owns column c; x_k (k > c) lives at S[c][k]
    if (tid < w) {
        const int c = tid;
        const double xc = 1.0 / S[c * DP + c];
        dinv[c] = xc;
        for (int i = c + 1; i < w; ++i) {
            double s = S[i * DP + c] * xc;
            for (int k = c + 1; k < i; ++k)
                s += S[i * DP + k] * S[c * DP + k];
            S[c * DP + i] = -s / S[i * DP + i];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < w * w; idx += NB) {
        int i = idx / w, j = idx - i * w;
        // W[i][j] (row-major, lower)
        if (op.W)
            op.W[i * op.ldw + j] = (j < i) ? S[j * DP + i] : (j == i ? dinv[i] : 0.0);
        // Wt[i][j] = W[j][i] (upper)
        if (op.Wt)
            op.Wt[i * op.ldwt + j] = (j > i) ? S[i * DP + j] : (j == i ? dinv[i] : 0.0);
    }
}

// ---- triangular solve with one pivot tile ---------------------------------------------
__global__ void __launch_bounds__(NB, 1) tri_kernel(const TriOp* __restrict__ ops, int backward)
{
    extern __shared__ double S[];
    __shared__ double xs[NB];
    const TriOp op = ops[blockIdx.x];
    const int w = op.w, tid = threadIdx.x;
    const int64_t ld = op.ldd;
    for (int idx = tid; idx < w * w; idx += NB) {
        int i = idx / w, j = idx - i * w;
        if (j <= i)
            S[i * DP + j] = op.D[i * ld + j];
    }
    double x = tid < w ? op.x[tid] : 0.0;
    __syncthreads();
    if (!backward) {
        for (int j = 0; j < w; ++j) {
            if (tid == j) {
                x = x / S[j * DP + j];
                xs[j] = x;
            }
            __syncthreads();
            if (tid > j && tid < w)
                x -= S[tid * DP + j] * xs[j];
        }
    } else {
        for (int j = w - 1; j >= 0; --j) {
            if (tid == j) {
                x = x / S[j * DP + j];
                xs[j] = x;
            }
            __syncthreads();
            if (tid < j)
                x -= S[j * DP + tid] * xs[j];
        }
    }
    if (tid < w)
        op.x[tid] = x;
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
__global__ void __launch_bounds__(256) gather_kernel(const GatherOp* __restrict__ ops)
{
    __shared__ double T[GE][GE + 1];
    const GatherOp op = ops[blockIdx.y];
    const int ni = op.nb - op.jb, nj = op.je - op.jb;
    const int ti_n = (ni + GT - 1) / GT, tj_n = (nj + GT - 1) / GT;
    for (int tl = blockIdx.x; tl < ti_n * tj_n; tl += gridDim.x) {
        const int ti = tl / tj_n, tj = tl - ti * tj_n;
        const int i0 = ti * GT, j0 = tj * GT;            // station offsets relative to jb
        if (i0 + GT - 1 < j0)
            continue;                                    // tile wholly above the diagonal (uniform per block)
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

constexpr int TILE_SMEM = NB * DP * 8;

}  // namespace

void launch_diag(const DiagOp* ops, int nops, int* info, void* stream)
{
    if (nops <= 0)
        return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM);
        cudaFuncSetAttribute(tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM);
        configured = true;
    }
    diag_kernel<<<nops, NB, TILE_SMEM, (cudaStream_t)stream>>>(ops, info);
}

void launch_tri(const TriOp* ops, int nops, int backward, void* stream)
{
    if (nops <= 0)
        return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM);
        cudaFuncSetAttribute(tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM);
        configured = true;
    }
    tri_kernel<<<nops, NB, TILE_SMEM, (cudaStream_t)stream>>>(ops, backward);
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

void launch_gather(const GatherOp* ops, int nops, int grid_x, void* stream)
{
    if (nops <= 0)
        return;
    for (int o = 0; o < nops; o += 65535) {
        int n = nops - o < 65535 ? nops - o : 65535;
        gather_kernel<<<dim3(grid_x < 1 ? 1 : (grid_x > 592 ? 592 : grid_x), n), 256, 0, (cudaStream_t)stream>>>(ops + o);
    }
}

}  // namespace gadj
