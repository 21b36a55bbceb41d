// schedule.hpp -- how hare_shoot_batch cuts a device's share of a batch into pipelined chunks (host code; also compiled into
// tests/emu/ for its property test).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace hare {

// Chunk sizes of a device's share of a hare_shoot_batch call.  A traversal launch has a fixed part of about 2 ms (pool fill and drain
// at low lane occupancy, the last CTAs, a cold L2: C3 runs 606 Mrays/s on 4 M rays, 742 on 16 M, 814 on 64 M), so few large launches
// beat many small ones -- but what is copied in before the first kernel and out after the last one is not overlapped with anything.  So the chunks start
// small (an eighth of the share, 2^18 .. 2^20 rays), grow by half each time up to 2^24 rays, and shrink again the same way towards
// the end; the H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap on the three streams.  (Uniform
// 4 M-ray chunks: 163 ms for the 100 M rays of C3 where the kernel alone takes 121.)
inline std::vector<int64_t> shoot_schedule(int64_t n) {
    const int64_t first = std::min<int64_t>(1 << 20, std::max<int64_t>(1 << 18, n / 8)), maxc = 1 << 24;
    std::vector<int64_t> head;
    int64_t used = 0;
    for (int64_t c = first; c < maxc && 2 * used + 2 * c < n; c += c / 2) { head.push_back(c); used += c; }
    const int64_t mid = n - 2 * used, nm = std::max<int64_t>(1, (mid + maxc - 1) / maxc);
    std::vector<int64_t> v(head);
    for (int64_t k = 0; k < nm; ++k) v.push_back(mid * (k + 1) / nm - mid * k / nm);
    v.insert(v.end(), head.rbegin(), head.rend());
    return v;
}

}  // namespace hare
