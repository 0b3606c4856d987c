// dev_cuda.cu — CUDA-runtime implementation of dev.h (product build).
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>

#include "dev.h"
#include "kernels.h"

namespace gadj {
namespace dev {

struct Device {
    int ordinal = -1;
    cudaStream_t stream = nullptr;
    int sms = 148;
    std::set<int> used_keys;        // one-time per-device set-up (kernel attributes)
    std::set<int> peers_enabled;    // same-process peer access already switched on towards these ordinals
};

namespace {

thread_local Device* t_cur = nullptr;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_mutex;

std::string cuda_err(cudaError_t e, const char* what)
{
    return std::string(what) + ": " + cudaGetErrorString(e);
}

}  // namespace

Device* open(int device_ordinal, std::string& err)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        err = std::string("no CUDA device available (") + cudaGetErrorString(e) + "): the adjustment engine has no CPU fallback";
        return nullptr;
    }
    if (device_ordinal < 0 || device_ordinal >= count) {
        err = "CUDA device ordinal out of range";
        return nullptr;
    }
    e = cudaSetDevice(device_ordinal);
    if (e != cudaSuccess) {
        err = cuda_err(e, "cudaSetDevice");
        return nullptr;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device_ordinal);
    if (e != cudaSuccess) {
        err = cuda_err(e, "cudaGetDeviceProperties");
        return nullptr;
    }
    if (prop.major != 10) {
        err = std::string("device '") + prop.name + "' is not sm_100: this library is built for B200 (sm_100a) only";
        return nullptr;
    }
    Device* d = new Device();
    d->ordinal = device_ordinal;
    d->sms = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        err = cuda_err(e, "cudaStreamCreate");
        delete d;
        return nullptr;
    }
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        if (!g_encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
            if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
                err = "cuTensorMapEncodeTiled is not available from the driver";
                cudaStreamDestroy(d->stream);
                delete d;
                return nullptr;
            }
            g_encode = (EncodeTiledFn)fn;
        }
    }
    t_cur = d;
    return d;
}

void close(Device* d)
{
    if (!d)
        return;
    cudaSetDevice(d->ordinal);
    if (d->stream)
        cudaStreamDestroy(d->stream);
    if (t_cur == d)
        t_cur = nullptr;
    delete d;
}

void use(Device* d)
{
    if (t_cur != d || d == nullptr) {
        t_cur = d;
    }
    if (d) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != d->ordinal)
            cudaSetDevice(d->ordinal);
    }
}

int ordinal() { return t_cur ? t_cur->ordinal : -1; }
int sm_count() { return t_cur ? t_cur->sms : 148; }
bool first_use(int key) { return t_cur ? t_cur->used_keys.insert(key).second : true; }

bool is_cuda() { return true; }
void* stream() { return t_cur ? (void*)t_cur->stream : nullptr; }

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
void zero(void* p, size_t bytes) { cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream()); }
void h2d(void* dst, const void* src, size_t bytes)
{
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream());
}
void d2h(void* dst, const void* src, size_t bytes)
{
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream());
}
void d2d(void* dst, const void* src, size_t bytes)
{
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream());
}

std::string sync()
{
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream());
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
void event_record(void* e) { cudaEventRecord((cudaEvent_t)e, (cudaStream_t)stream()); }
float event_elapsed_ms(void* a, void* b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

// ---- peer memory ------------------------------------------------------------------------------------
void* alloc_shared(size_t bytes) { return alloc(bytes); }
void free_shared(void* p) { free_(p); }
int64_t process_id() { return (int64_t)getpid(); }

bool ipc_export(void* p, size_t, void* handle)
{
    static_assert(sizeof(cudaIpcMemHandle_t) <= IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    std::memset(handle, 0, IPC_HANDLE_BYTES);
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    std::memcpy(handle, &h, sizeof(h));
    return true;
}

void* peer_map(int peer_ordinal, int64_t peer_pid, void* raw, const void* handle, size_t, std::string& err)
{
    if (peer_pid == process_id()) {
        // a rank driven by another thread of this process: its pointers are valid here once peer access is on
        if (t_cur && peer_ordinal != t_cur->ordinal && !t_cur->peers_enabled.count(peer_ordinal)) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, t_cur->ordinal, peer_ordinal);
            if (!can) {
                err = "GPUs " + std::to_string(t_cur->ordinal) + " and " + std::to_string(peer_ordinal) + " have no peer access";
                return nullptr;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(peer_ordinal, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                err = cuda_err(e, "cudaDeviceEnablePeerAccess");
                return nullptr;
            }
            cudaGetLastError();
            t_cur->peers_enabled.insert(peer_ordinal);
        }
        return raw;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        err = cuda_err(e, "cudaIpcOpenMemHandle");
        return nullptr;
    }
    return p;
}

void peer_unmap(void* mapped, int64_t peer_pid)
{
    if (mapped && peer_pid != process_id())
        cudaIpcCloseMemHandle(mapped);
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
