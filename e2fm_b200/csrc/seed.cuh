// seed.cuh — the seed table: the SA interval of the LAST j super-characters of a shifted pattern in ONE sector gather.
//
// EFMIndex::exactMatch(interval) (EFMIndex.cpp:1901-1963) starts from [C[c], C[c+1]) and takes one backwardStep
// (:1966-2001) per super-character, i.e. 2 rank lookups each. The rows of the index are sorted by their rotations, so the
// interval reached after the last j super-characters c_0 .. c_{j-1} (text order) is "the rows whose rotation starts with
// c_0 .. c_{j-1}": its start is the number of rows whose first j super-characters compare smaller, its length the number
// of rows that start with exactly these j. Both are tabulated at load from the dense sa[] / text[] tables:
//
//   seed = (c_0, .., c_{j-2} | c_{j-1}):  hi = ((c_0 * A + c_1) * A + ..) + c_{j-2}   (all but the last super-character)
//   group g = hi * gph + c_{j-1} / 56,  slot = c_{j-1} % 56,  gph = ceil(A / 56)
//   sector g (32 bytes): w0 = rows whose seed is smaller than the group's first seed
//                        w1..w7 = 56 4-bit counters: rows per seed of the group, 15 = "15 or more" (not exact)
//   interval = [w0 + sum of the counters below the slot, + counter of the slot)
//
// A lookup that meets a saturated counter at or below its slot falls back to the backward steps it replaces. j is the
// smallest number of super-characters with A^j >= n/2 (about one row per seed or fewer: for the 3.1 Gbp collection j = 4,
// 2.75 GB), so that a shift that does not occur in the text is rejected by this one gather with probability 1 - n/A^j.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "device_index.cuh"

namespace e2fm {

constexpr uint32_t SEED_GROUP = 56;     // seeds per sector (7 words x 8 nibbles)
constexpr uint32_t SEED_SAT = 15;       // saturated counter

struct SeedHit {
    uint32_t s, cnt;    // interval [s, s + cnt)
    bool exact;         // false: a saturated counter was met, the interval is unknown
};

__device__ __forceinline__ size_t seed_sector(const DevIndex &ix, uint32_t hi, uint32_t c_last, uint32_t &slot) {
    const uint32_t q = c_last / SEED_GROUP;
    slot = c_last - q * SEED_GROUP;
    return (size_t)hi * ix.seed_gph + q;
}

__device__ __forceinline__ Sector seed_load(const DevIndex &ix, size_t sector) {
    return ld_sector(ix.seed_tab + sector * 2);
}

__device__ __forceinline__ SeedHit seed_decode(const Sector &sec, uint32_t slot) {
    const uint32_t w[7] = {sec.lo.y, sec.lo.z, sec.lo.w, sec.hi.x, sec.hi.y, sec.hi.z, sec.hi.w};
    const uint32_t wi = slot >> 3, sh = (slot & 7u) * 4u;
    uint32_t acc = 0, sat = 0, own = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const uint32_t m = (uint32_t)i < wi ? 0xFFFFFFFFu : ((uint32_t)i == wi ? ((1u << sh) - 1u) : 0u);
        const uint32_t x = w[i] & m;
        sat |= x & (x >> 1) & (x >> 2) & (x >> 3);
        acc += (x & 0x0F0F0F0Fu) + ((x >> 4) & 0x0F0F0F0Fu);     // byte sums: at most 7 words x 30
        if ((uint32_t)i == wi) own = w[i];
    }
    SeedHit h;
    h.cnt = (own >> sh) & 15u;
    h.s = sec.lo.x + (acc & 0xFFu) + ((acc >> 8) & 0xFFu) + ((acc >> 16) & 0xFFu) + (acc >> 24);
    h.exact = (sat & 0x11111111u) == 0u && h.cnt != SEED_SAT;
    return h;
}

}  // namespace e2fm
