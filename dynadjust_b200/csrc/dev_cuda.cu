// dev_cuda.cu — CUDA-runtime implementation of dev.h (product build).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "dev.h"
#include "kernels.h"

namespace gadj {
namespace dev {
namespace {

cudaStream_t g_stream = nullptr;
int g_device = -1;
std::string g_async_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

std::string cuda_err(cudaError_t e, const char* what)
{
    return std::string(what) + ": " + cudaGetErrorString(e);
}

}  // namespace

std::string init(int device_ordinal)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return std::string("no CUDA device available (") + cudaGetErrorString(e) +
               "): the adjustment engine has no CPU fallback";
    if (device_ordinal < 0 || device_ordinal >= count)
        return "CUDA device ordinal out of range";
    if (g_stream && g_device == device_ordinal)
        return std::string();
    e = cudaSetDevice(device_ordinal);
    if (e != cudaSuccess)
        return cuda_err(e, "cudaSetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device_ordinal);
    if (e != cudaSuccess)
        return cuda_err(e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return std::string("device '") + prop.name + "' is not sm_100: this library is built for B200 (sm_100a) only";
    if (g_stream)
        cudaStreamDestroy(g_stream);
    e = cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess)
        return cuda_err(e, "cudaStreamCreate");
    g_device = device_ordinal;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
        return "cuTensorMapEncodeTiled is not available from the driver";
    g_encode = (EncodeTiledFn)fn;
    return std::string();
}

bool is_cuda() { return true; }
void* stream() { return (void*)g_stream; }

void* alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 256) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void free_(void* p) { cudaFree(p); }
void* alloc_host_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void free_host_pinned(void* p) { cudaFreeHost(p); }
void zero(void* p, size_t bytes) { cudaMemsetAsync(p, 0, bytes, g_stream); }
void h2d(void* dst, const void* src, size_t bytes) { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream); }
void d2h(void* dst, const void* src, size_t bytes) { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream); }
void d2d(void* dst, const void* src, size_t bytes) { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_stream); }

std::string sync()
{
    cudaError_t e = cudaStreamSynchronize(g_stream);
    if (e == cudaSuccess)
        e = cudaGetLastError();
    if (e != cudaSuccess)
        return cuda_err(e, "CUDA failure on the adjustment stream");
    return std::string();
}

size_t mem_free()
{
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    return f;
}
size_t mem_total()
{
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    return t;
}

void* event_create()
{
    cudaEvent_t e;
    cudaEventCreate(&e);
    return (void*)e;
}
void event_destroy(void* e) { cudaEventDestroy((cudaEvent_t)e); }
void event_record(void* e) { cudaEventRecord((cudaEvent_t)e, g_stream); }
float event_elapsed_ms(void* a, void* b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

bool encode_tma_2d(void* desc128, const double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows)
{
    if (!g_encode)
        return false;
    static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "tensor map size");
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)TILE_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (getenv("GADJ_DEBUG"))
            fprintf(stderr, "gadj: cuTensorMapEncodeTiled failed (%d) base=%p rows=%llu cols=%llu ld=%llu\n", (int)r,
                    (const void*)base, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
        return false;
    }
    std::memcpy(desc128, &tm, sizeof(tm));
    return true;
}

}  // namespace dev
}  // namespace gadj
