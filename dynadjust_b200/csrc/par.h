// par.h — small host-side parallel helpers for the one-off preparation (record scan, sorting, symbolic analysis).
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace gadj {

inline unsigned host_threads()
{
    return std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
}

// fn(begin, end) over [0, n) in contiguous chunks, one per thread
template <class F>
void parallel_for(uint64_t n, F&& fn, uint64_t serial_below = 65536)
{
    const unsigned nt = host_threads();
    if (n < serial_below || nt == 1) {
        fn((uint64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    const uint64_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const uint64_t b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e)
            break;
        th.emplace_back([=, &fn] { fn(b, e); });
    }
    for (auto& t : th)
        t.join();
}

// ascending sort: chunks sorted side by side, then merged pairwise (the merges of a round side by side as well)
template <class T>
void parallel_sort(std::vector<T>& v)
{
    const unsigned nt = host_threads();
    const size_t n = v.size();
    if (n < (1u << 18) || nt == 1) {
        std::sort(v.begin(), v.end());
        return;
    }
    unsigned parts = 1;
    while (parts * 2 <= nt)
        parts *= 2;
    std::vector<size_t> cut(parts + 1);
    for (unsigned p = 0; p <= parts; ++p)
        cut[p] = n * p / parts;
    {
        std::vector<std::thread> th;
        for (unsigned p = 0; p < parts; ++p)
            th.emplace_back([&, p] { std::sort(v.begin() + cut[p], v.begin() + cut[p + 1]); });
        for (auto& t : th)
            t.join();
    }
    for (unsigned width = 1; width < parts; width *= 2) {
        std::vector<std::thread> th;
        for (unsigned p = 0; p + width < parts; p += 2 * width)
            th.emplace_back([&, p, width] {
                std::inplace_merge(v.begin() + cut[p], v.begin() + cut[p + width], v.begin() + cut[std::min(parts, p + 2 * width)]);
            });
        for (auto& t : th)
            t.join();
    }
}

}  // namespace gadj
