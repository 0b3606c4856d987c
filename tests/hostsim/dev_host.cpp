// tests/hostsim/dev_host.cpp — TEST INFRASTRUCTURE.
// malloc/memcpy implementation of csrc/dev.h so that the planner, the index maps and
// the engine's control flow can be exercised on a machine without a GPU.  Linked
// only into tests/hostsim/_build/libgadj_hostsim.so, never into the product library.
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "../../dynadjust_b200/csrc/dev.h"

namespace gadj {
namespace dev {

std::string init(int) { return std::string(); }
bool is_cuda() { return false; }
void* stream() { return nullptr; }
void* alloc(size_t bytes)
{
    // poison fresh device memory with NaNs: a read of anything the engine has not written shows up in the results
    void* p = std::malloc(bytes ? bytes : 1);
    if (p)
        std::memset(p, 0xFF, bytes ? bytes : 1);
    return p;
}
void free_(void* p) { std::free(p); }
void* alloc_host_pinned(size_t bytes) { return std::malloc(bytes); }
void free_host_pinned(void* p) { std::free(p); }
void zero(void* p, size_t bytes) { std::memset(p, 0, bytes); }
void h2d(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
void d2h(void* dst, const void* src, size_t bytes) { std::memcpy(dst, src, bytes); }
void d2d(void* dst, const void* src, size_t bytes) { std::memmove(dst, src, bytes); }
std::string sync() { return std::string(); }
size_t mem_free() { return (size_t)8 << 30; }
size_t mem_total() { return (size_t)8 << 30; }

struct Ev {
    std::chrono::steady_clock::time_point t;
};
void* event_create() { return new Ev(); }
void event_destroy(void* e) { delete (Ev*)e; }
void event_record(void* e) { ((Ev*)e)->t = std::chrono::steady_clock::now(); }
float event_elapsed_ms(void* a, void* b)
{
    return std::chrono::duration<float, std::milli>(((Ev*)b)->t - ((Ev*)a)->t).count();
}
bool encode_tma_2d(void*, const double*, uint64_t, uint64_t, uint64_t, uint32_t) { return true; }

}  // namespace dev
}  // namespace gadj
