// tests/hostsim/dev_host.cpp — TEST INFRASTRUCTURE.
// malloc/memcpy implementation of csrc/dev.h so that the planner, the index maps and
// the engine's control flow can be exercised on a machine without a GPU.  Linked
// only into tests/hostsim/_build/libgadj_hostsim.so, never into the product library.
//
// "Peer memory" (the buffers the ranks of a multi-GPU run read and write directly) is POSIX shared
// memory here, so that the multi-rank CPU tests run one process per rank and pass handles around exactly like the
// CUDA build passes cudaIpc handles; ranks that are threads of one process use the raw pointers.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>

#include "../../dynadjust_b200/csrc/dev.h"

namespace gadj {
namespace dev {

struct Device {
    int ordinal = 0;
    std::set<int> used_keys;
};

namespace {
thread_local Device* t_cur = nullptr;
std::mutex g_mutex;
struct Shm {
    std::string name;
    size_t bytes;
};
std::map<void*, Shm> g_shm;           // allocations of this process
std::map<void*, size_t> g_mapped;     // peers' allocations mapped into this process
std::atomic<uint64_t> g_counter{0};
}  // namespace

Device* open(int device_ordinal, std::string&)
{
    Device* d = new Device();
    d->ordinal = device_ordinal;
    t_cur = d;
    return d;
}
void close(Device* d)
{
    if (t_cur == d)
        t_cur = nullptr;
    delete d;
}
void use(Device* d) { t_cur = d; }
int ordinal() { return t_cur ? t_cur->ordinal : 0; }
int sm_count() { return 4; }
bool first_use(int key) { return t_cur ? t_cur->used_keys.insert(key).second : true; }

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

int64_t process_id() { return (int64_t)getpid(); }

void* alloc_shared(size_t bytes)
{
    if (bytes == 0)
        bytes = 8;
    char name[IPC_HANDLE_BYTES];
    std::snprintf(name, sizeof(name), "/gadj_hostsim_%ld_%llu", (long)getpid(), (unsigned long long)g_counter.fetch_add(1));
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0)
        return nullptr;
    if (ftruncate(fd, (off_t)bytes) != 0) {
        ::close(fd);
        shm_unlink(name);
        return nullptr;
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) {
        shm_unlink(name);
        return nullptr;
    }
    std::memset(p, 0xFF, bytes);
    std::lock_guard<std::mutex> lk(g_mutex);
    g_shm[p] = Shm{name, bytes};
    return p;
}

void free_shared(void* p)
{
    if (!p)
        return;
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = g_shm.find(p);
    if (it == g_shm.end())
        return;
    munmap(p, it->second.bytes);
    shm_unlink(it->second.name.c_str());
    g_shm.erase(it);
}

bool ipc_export(void* p, size_t, void* handle)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = g_shm.find(p);
    if (it == g_shm.end())
        return false;
    std::memset(handle, 0, IPC_HANDLE_BYTES);
    std::memcpy(handle, it->second.name.c_str(), it->second.name.size());
    return true;
}

void* peer_map(int, int64_t peer_pid, void* raw, const void* handle, size_t bytes, std::string& err)
{
    if (peer_pid == process_id())
        return raw;
    char name[IPC_HANDLE_BYTES + 1] = {0};
    std::memcpy(name, handle, IPC_HANDLE_BYTES);
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) {
        err = std::string("shm_open failed for ") + name;
        return nullptr;
    }
    void* p = mmap(nullptr, bytes ? bytes : 8, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) {
        err = "mmap of a peer buffer failed";
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_mutex);
    g_mapped[p] = bytes ? bytes : 8;
    return p;
}

void peer_unmap(void* mapped, int64_t peer_pid)
{
    if (!mapped || peer_pid == process_id())
        return;
    std::lock_guard<std::mutex> lk(g_mutex);
    auto it = g_mapped.find(mapped);
    if (it != g_mapped.end()) {
        munmap(mapped, it->second);
        g_mapped.erase(it);
    }
}

bool encode_tma_2d(void*, const double*, uint64_t, uint64_t, uint64_t, uint32_t) { return true; }

}  // namespace dev
}  // namespace gadj
