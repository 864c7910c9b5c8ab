// device_index.cuh — the flattened, device-resident layout of one E2FM index and the rank / LF
// primitives every search kernel uses.
//
// Layout (all in HBM; sizes for n super-characters, compact alphabet A):
//   rec[n]        u64 per BWT row:  bit 63 marked | bits 32..47 compact code | bits 0..31 LF(row)
//                 (for a marked row the low word holds the marked value = textPos/markingRate instead)
//   occ[nBlk][A]  one 32-byte SECTOR per (rank block, symbol):
//                   u32 base  = C[c] + #{rows < blockStart : bwt[row]==c}
//                   u16 off[14] ascending in-block offsets of the occurrences of c, 0xFFFF padded
//                 so that  C[c] + occ(c,row) = base + #{off <= row % BLK}  costs ONE sector gather.
//                 A block with more than 14 occurrences sets bit 15 of off[0]; off[0..11] hold the
//                 first 12 offsets and off[12..13] the index (u32) of the rest in the `ovf` pool.
//   markTab[mrn]  uint2 {row, LF(row)} of the row whose marked value is v  (reverse of rec's marks;
//                 replaces markedRowsBuckets + Bucket::getMarkedBwtPosition, Bucket.cpp:794)
// The reference's two-level counters + per-bucket position lists (SuperBucket.cpp:82, Bucket.cpp:656-779)
// answer the same queries with 5-7 dependent sector reads; here memory is traded for latency because a
// B200 has 180 GB of HBM and the path is bound by dependent random gathers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace e2fm {

constexpr uint64_t REC_MARKED = 1ull << 63;
constexpr int SECTOR_SLOTS = 14;
constexpr int SECTOR_INLINE_OVF = 12;

struct DevIndex {
    uint64_t n;                 // rows
    uint64_t initial_n;
    uint64_t mrn;               // number of marked rows
    uint32_t k, n_sym, A, rate;
    uint32_t blk_shift;         // log2(rows per rank block)
    uint32_t n_blk;
    uint32_t sigma_k;
    uint32_t n_term;
    const uint64_t *rec;
    const uint4 *occ;           // 2 x uint4 per sector
    const uint16_t *ovf;
    const uint2 *mark_tab;
    const uint64_t *C;          // [A+1]
    const uint16_t *compact_of_natural;   // [sigma_k]
    const uint32_t *natural_of_compact;   // [A]
    const uint8_t *sym_index;   // [256] ASCII -> index in the original alphabet, 0xFF if absent
    const int64_t *terminators; // [n_term]
};

struct OpCounters { unsigned long long rank, lf, mark; };

__device__ __forceinline__ uint32_t rec_code(uint64_t r) { return (uint32_t)(r >> 32) & 0xFFFFu; }
__device__ __forceinline__ uint32_t rec_low(uint64_t r) { return (uint32_t)r; }
__device__ __forceinline__ bool rec_marked(uint64_t r) { return (r >> 63) != 0; }

// number of 16-bit lanes of w that are <= x (x < 0x8000), both halves
__device__ __forceinline__ uint32_t count_le2(uint32_t w, uint32_t xx) {
    return __popc(__vcmpleu2(w, xx)) >> 4;
}

// C[c] + #{rows <= row : bwt[row]==c}  — one 32-byte gather (two 16-byte loads of the same sector).
__device__ __forceinline__ uint32_t rank_c(const DevIndex &ix, uint32_t c, uint32_t row) {
    const uint32_t blk = row >> ix.blk_shift;
    const uint32_t x = row & ((1u << ix.blk_shift) - 1);
    const uint4 *s = ix.occ + ((size_t)blk * ix.A + c) * 2;
    const uint4 lo = __ldg(s), hi = __ldg(s + 1);
    const uint32_t xx = x | (x << 16);
    const uint32_t first = lo.y & 0xFFFFu;
    if ((first & 0x8000u) == 0 || first == 0xFFFFu) {
        return lo.x + count_le2(lo.y, xx) + count_le2(lo.z, xx) + count_le2(lo.w, xx) + count_le2(hi.x, xx) +
               count_le2(hi.y, xx) + count_le2(hi.z, xx) + count_le2(hi.w, xx);
    }
    // overflow sector: off[0] carries the flag, off[12..13] the pool index
    uint32_t cnt = ((first & 0x7FFFu) <= x) + (((lo.y >> 16)) <= x);
    cnt += count_le2(lo.z, xx) + count_le2(lo.w, xx) + count_le2(hi.x, xx) + count_le2(hi.y, xx) + count_le2(hi.z, xx);
    if (cnt == SECTOR_INLINE_OVF) {
        const uint16_t *p = ix.ovf + hi.w;
        while (true) {
            uint32_t o = __ldg(p++);
            if (o > x) break;   // 0xFFFF terminator compares greater than any offset
            cnt++;
        }
    }
    return lo.x + cnt;
}

// LF(row) given the row's record. Unmarked rows carry it; a marked row's low word is its marked value,
// so its LF is recomputed from the rank sector (2% of rows at the default marking rate).
__device__ __forceinline__ uint32_t lf_of(const DevIndex &ix, uint64_t r, uint32_t row) {
    if (!rec_marked(r)) return rec_low(r);
    return rank_c(ix, rec_code(r), row) - 1;
}

__device__ __forceinline__ uint64_t ld_rec(const DevIndex &ix, uint32_t row) { return __ldg(ix.rec + row); }

}  // namespace e2fm
