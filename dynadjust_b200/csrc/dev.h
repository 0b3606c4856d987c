// dev.h — minimal device-runtime seam (memory, copies, stream, timing).
// dev_cuda.cu implements it on the CUDA runtime for the product library;
// tests/hostsim/dev_host.cpp implements it with malloc/memcpy for the CPU-only
// planner tests.  There is no runtime selection between the two: each binary
// links exactly one.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace gadj {
namespace dev {

// returns empty string on success, else the backend's error text
std::string init(int device_ordinal);
bool is_cuda();                       // false only in the hostsim test build
void* stream();                       // the context's compute stream handle
void* alloc(size_t bytes);            // nullptr on failure
void free_(void* p);
void* alloc_host_pinned(size_t bytes);
void free_host_pinned(void* p);
void zero(void* p, size_t bytes);                       // async on the stream
void h2d(void* dst, const void* src, size_t bytes);     // async on the stream
void d2h(void* dst, const void* src, size_t bytes);     // async on the stream
void d2d(void* dst, const void* src, size_t bytes);     // async on the stream
std::string sync();                   // wait for the stream; returns error text if any launch failed
size_t mem_free();
size_t mem_total();

// event timing on the compute stream
void* event_create();
void event_destroy(void* e);
void event_record(void* e);
float event_elapsed_ms(void* a, void* b);

// encode a 2-D FP64 row-major tensor map: rows x cols, pitch ld (doubles), box = box_rows x TILE_K, SWIZZLE_128B.
// no-op in the hostsim build.  returns false on failure.
bool encode_tma_2d(void* desc128, const double* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

}  // namespace dev
}  // namespace gadj
