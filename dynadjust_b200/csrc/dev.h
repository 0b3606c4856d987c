// dev.h — minimal device-runtime seam (memory, copies, stream, timing, peer memory).
// dev_cuda.cu implements it on the CUDA runtime for the product library;
// tests/hostsim/dev_host.cpp implements it with malloc/memcpy for the CPU-only
// planner tests.  There is no runtime selection between the two: each binary
// links exactly one.
//
// Every adjustment context owns a Device (device ordinal, its own stream, per-device one-time
// kernel attributes).  The free functions below act on the calling thread's *current* device,
// which every C-ABI entry point sets first (dev::use) — two contexts in one process, on the same
// GPU or on different ones (one host thread per GPU), never share a stream or scratch memory.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace gadj {
namespace dev {

struct Device;

// nullptr on failure (err receives the backend's text)
Device* open(int device_ordinal, std::string& err);
void close(Device* d);
void use(Device* d);                  // make d current for the calling thread
int ordinal();                        // device ordinal of the current device
int sm_count();                       // streaming multiprocessors of the current device
bool first_use(int key);              // true the first time `key` is asked for on the current device

bool is_cuda();                       // false only in the hostsim test build
void* stream();                       // the current device's compute stream handle
void* alloc(size_t bytes);            // nullptr on failure
void free_(void* p);
void zero(void* p, size_t bytes);                       // async on the stream
void h2d(void* dst, const void* src, size_t bytes);     // async on the stream
void d2h(void* dst, const void* src, size_t bytes);     // async on the stream
void d2d(void* dst, const void* src, size_t bytes);     // async on the stream (dst / src may be peer mappings)
std::string sync();                   // wait for the stream; returns error text if any launch failed
size_t mem_free();
size_t mem_total();

// event timing on the compute stream
void* event_create();
void event_destroy(void* e);
void event_record(void* e);
float event_elapsed_ms(void* a, void* b);

// ---- peer memory (multi-GPU: one rank per GPU, threads of one process or separate processes) ----
// Memory that other ranks read / write directly over NVLink.  alloc_shared is cudaMalloc on the device (every
// cudaMalloc allocation can be exported); the hostsim backs it with POSIX shared memory so that CPU tests with one
// process per rank exercise the same exchange of handles.
constexpr size_t IPC_HANDLE_BYTES = 64;
void* alloc_shared(size_t bytes);
void free_shared(void* p);
bool ipc_export(void* p, size_t bytes, void* handle /* IPC_HANDLE_BYTES */);
// maps a peer allocation into this rank: same process -> the raw pointer (peer access enabled between the two devices),
// another process -> the IPC handle is opened.  nullptr on failure.
void* peer_map(int peer_ordinal, int64_t peer_pid, void* raw, const void* handle, size_t bytes, std::string& err);
void peer_unmap(void* mapped, int64_t peer_pid);
int64_t process_id();

// encode a 2-D FP64 row-major tensor map: rows x cols, pitch ld (doubles), box = box_rows x TILE_K, SWIZZLE_128B.
// no-op in the hostsim build.  returns false on failure.
bool encode_tma_2d(void* desc128, const double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

}  // namespace dev
}  // namespace gadj
