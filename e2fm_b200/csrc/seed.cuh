// seed.cuh — the seed table: the SA interval of the LAST j super-characters of a shifted pattern in ONE sector gather.
//
// EFMIndex::exactMatch(interval) (EFMIndex.cpp:1901-1963) starts from [C[c], C[c+1]) and takes one backwardStep
// (:1966-2001) per super-character, i.e. 2 rank lookups each. The rows of the index are sorted by their rotations, so the
// interval reached after the last j super-characters c_0 .. c_{j-1} (text order) is "the rows whose rotation starts with
// c_0 .. c_{j-1}": its start is the number of rows whose first j super-characters compare smaller, its length the number
// of rows that start with exactly these j. Both are tabulated at load from the dense sa[] / text[] tables:
//
//   seed = (c_0, .., c_{j-2} | c_{j-1}):  hi = ((c_0 * A + c_1) * A + ..) + c_{j-2}   (all but the last super-character)
//   group g = hi * gph + c_{j-1} / 24,  slot = c_{j-1} % 24,  gph = ceil(A / 24)
//   entry g (16 bytes): w0 = rows whose seed is smaller than the group's first seed (bit 31: a counter of the group is saturated)
//                       w1..w3 = 24 4-bit counters: rows per seed of the group, 15 = "15 or more" (not exact)
//   interval = [w0 + sum of the counters below the slot, + counter of the slot)
//
// ONE 16-byte load answers a lookup (the DRAM moves a 64-byte burst for it either way; what the 16-byte entry saves against a
// 32-byte sector of 56 counters is L2->SM traffic, decode work and four registers per gather in flight, which is what lets a thread
// keep four lookups outstanding). A lookup in a group that holds a saturated counter falls back to the backward steps the table
// replaces. j is the smallest number of super-characters with A^j >= n/2 (about one row per seed or fewer: for the 3.1 Gbp
// collection j = 4, 3.0 GB), so that a shift that does not occur in the text is rejected by this one gather with probability
// 1 - n/A^j.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "device_index.cuh"

namespace e2fm {

constexpr uint32_t SEED_GROUP = 24;     // seeds per 16-byte entry (3 words x 8 nibbles)
constexpr uint32_t SEED_SAT = 15;       // saturated counter

struct SeedHit {
    uint32_t s, cnt;    // interval [s, s + cnt)
    bool exact;         // false: a counter of the group is saturated, the interval is unknown
};

__device__ __forceinline__ size_t seed_sector(const DevIndex &ix, uint32_t hi, uint32_t c_last, uint32_t &slot) {
    const uint32_t q = c_last / SEED_GROUP;
    slot = c_last - q * SEED_GROUP;
    return (size_t)hi * ix.seed_gph + q;
}

__device__ __forceinline__ uint4 seed_load(const DevIndex &ix, size_t entry) {
    uint4 r;
    asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ix.seed_tab + entry));
    return r;
}

__device__ __forceinline__ SeedHit seed_decode(const uint4 &e, uint32_t slot) {
    const uint32_t w[3] = {e.y, e.z, e.w};
    const uint32_t wi = slot >> 3, sh = (slot & 7u) * 4u;
    uint32_t acc = 0, own = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        // the counters of word i that lie below the slot: all of them when i < wi, the low ones when i == wi
        uint32_t x = w[i];
        if ((uint32_t)i == wi) { own = x; x &= (1u << sh) - 1u; }
        else if ((uint32_t)i > wi) x = 0u;
        acc = __dp4a(x & 0x0F0F0F0Fu, 0x01010101u, acc);           // sum of the four low nibbles
        acc = __dp4a((x >> 4) & 0x0F0F0F0Fu, 0x01010101u, acc);    // ... and of the four high ones
    }
    SeedHit h;
    h.cnt = (own >> sh) & 15u;
    h.s = (e.x & 0x7FFFFFFFu) + acc;
    h.exact = (e.x >> 31) == 0u;           // bit 31 of the base (rows are below 2^31): a counter of this group is saturated
    return h;
}

}  // namespace e2fm
