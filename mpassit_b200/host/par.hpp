// par.hpp -- split an index range across a few host threads (coordinate tables, byte-order conversion of the grid
// file's geometry arrays).  Each piece is independent; results do not depend on the thread count.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace par {

inline unsigned threads_for(int64_t n, int64_t grain) {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    int64_t want = (n + grain - 1) / grain;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(hw, 16), want));
}

template <typename F>
void range(int64_t n, int64_t grain, F &&fn) {  // fn(begin, end)
    const unsigned nt = threads_for(n, grain);
    if (nt <= 1) {
        fn((int64_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    const int64_t step = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const int64_t b = t * step, e = std::min<int64_t>(n, b + step);
        if (b >= e) break;
        th.emplace_back([&fn, b, e] { fn(b, e); });
    }
    for (auto &x : th) x.join();
}

}  // namespace par
