// device_index.cuh — the flattened, device-resident layout of one E2FM index and the rank / LF
// primitives every search kernel uses.
//
// Layout (all in HBM; sizes for n super-characters, compact alphabet A):
//   rec[n]        u64 per BWT row:  bit 63 marked | bits 32..47 compact code | bits 0..31 LF(row)
//                 (for a marked row the low word holds the marked value = textPos/markingRate instead)
//   occ[nBlk][A]  one 32-byte SECTOR per (rank block, symbol), 8 words:
//                   w0     = C[c] + #{rows < blockStart : bwt[row]==c}
//                   w1..w7 = 14 u16 slots (slot j in word 1+j/2, half j&1): ascending in-block offsets of the
//                            occurrences of c, padded with 0x7FFF (> any offset: blocks hold <= 16384 rows)
//                 so that  C[c] + occ(c,row) = w0 + #{slot <= row % BLK}  costs ONE sector gather and a SWAR
//                 compare. A block with more than 14 occurrences sets bit 31 of w1 (bit 15 of slot 1): slots
//                 0..11 hold the first 12 offsets and w7 the index of the rest in the `ovf` pool (0xFFFF-terminated).
//   sa[n], text[n]  (dense locate tables, built once at load by walking LF from every marked row — the marked-SA-sample
//                 walk of EFMIndex::getRowPosition done in bulk for ALL rows instead of per query): sa[row] = super-text
//                 position of the row's suffix, text[p] = compact code of super-character p. With them locate is one
//                 gather per occurrence and hat verification (EFMIndex::extractSuperCharacter) one more. 6 bytes per row;
//                 skipped when HBM is short (the per-query walk kernel then does the same work on demand).
//   bwt16[n]      compact BWT code per row, 2 bytes (built with the dense tables): the '?'-prefixed first super-character
//                 makes the search test EVERY row of an interval (EFMIndex::backwardStepIfMatch); rows of an interval are
//                 consecutive, so that test streams 2 bytes per row instead of gathering 8-byte records.
//   pair_tab[A*A] uint2 {s,e}: the SA interval after the last two super-characters of a shift (start interval from C[] +
//                 one backwardStep, EFMIndex.cpp:1728-1748, :1966) in one gather; built at load with A*A rank pairs.
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
constexpr uint32_t VREC_LEFT = 10;      // super-characters before the row's text position held by a verification record
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
    const uint32_t *sa;         // [n] or nullptr
    const uint16_t *text;       // [n] or nullptr
    const uint16_t *bwt16;      // [n] or nullptr: compact BWT code of every row (2 bytes: wildcard rows are scanned in row order)
    const uint2 *pair_tab;      // [A*A] or nullptr: SA interval [s,e) after the last two super-characters (a = last, b = previous)
    const uint64_t *C;          // [A+1]
    const uint16_t *compact_of_natural;   // [sigma_k]
    const uint32_t *natural_of_compact;   // [A]
    const uint8_t *sym_index;   // [256] ASCII -> index in the original alphabet, 0xFF if absent
    const uint8_t *symbols;     // [n_sym] index -> ASCII
    // per compact code c and number of fixed symbols f in 1..k-1 (tables of (k-1)*A entries, index (f-1)*A + c):
    const uint32_t *kmer_low;   // natural(c) mod n_sym^f      — the f low-order symbols  ('?'-prefixed first super-character)
    const uint32_t *kmer_high;  // natural(c) div n_sym^(k-f)  — the f high-order symbols ('?'-suffixed last super-character)
    const int64_t *terminators; // [n_term]
    // seed table (seed.cuh): the interval of the last seed_chars super-characters of a shift in one gather; nullptr = none
    const uint4 *seed_tab;
    uint32_t seed_chars;        // j
    uint32_t seed_gph;          // 16-byte entries per value of the first j-1 super-characters = ceil(A / 24)
    // verification records (index_build.cu, vrec_build_kernel): [n] 32-byte sectors {sa[row], the VREC_LEFT super-characters before
    // that text position, the ones 2, 3 and 4 behind it}; nullptr = not built (HBM short): seed_verify_kernel reads sa[] / text[]
    const uint4 *vrec;
    // 2-bit packed patterns (A C G T = 0 1 2 3, base i of a k-mer at bits 2i): packed k-mer -> compact code, 0xFFFF when the
    // k-mer does not occur in the text or one of its symbols is not in the index's alphabet; nullptr when k > 7
    const uint16_t *packed_code; // [4^k]
    uint8_t sym4[4];            // index of A, C, G, T in the original alphabet (0xFF: not a symbol of this index)
};

struct OpCounters { unsigned long long rank, lf, mark; };

__device__ __forceinline__ uint32_t rec_code(uint64_t r) { return (uint32_t)(r >> 32) & 0xFFFFu; }
__device__ __forceinline__ uint32_t rec_low(uint64_t r) { return (uint32_t)r; }
__device__ __forceinline__ bool rec_marked(uint64_t r) { return (r >> 63) != 0; }

constexpr uint32_t SLOT_PAD = 0x7FFFu;
constexpr uint32_t SWAR_M = 0x80008000u;

struct Sector { uint4 lo, hi; };

// One aligned 32-byte sector in ONE load instruction (LDG.E.256, sm_100): measured on a B200 (profiles/exp/gather_variants.py),
// random sectors read with two 16-byte loads come in at 37 G sectors/s, with one 32-byte load at 47 G/s — the same rate as 8-byte
// loads, i.e. the row-activation limit of the HBM.
// Without a hint the L2 fetches the whole 128-byte line of a missing sector from DRAM (127 B of DRAM traffic per random gather,
// measured); L2::64B halves that. The gather RATE of the microbenchmark does not change (it is bound by row activations), but a
// kernel that mixes gathers with streaming traffic gets its bandwidth back.
__device__ __forceinline__ Sector ld_sector(const uint4 *p) {
    Sector r;
    asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w), "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w)
                 : "l"(p));
    return r;
}

// a lone 4-byte gather with the same hint (sa[])
__device__ __forceinline__ uint32_t ld_gather_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ Sector load_sector(const DevIndex &ix, uint32_t c, uint32_t blk) {
    return ld_sector(ix.occ + ((size_t)blk * ix.A + c) * 2);
}

// bits 15 and 31 of the result are set iff the low / high u16 slot of w is <= x (slots and x below 0x8000):
// (x | 0x8000) - slot keeps bit 15 exactly when no borrow is needed, and the set top bit stops the borrow
// from leaving its half.
__device__ __forceinline__ uint32_t swar_le(uint32_t w, uint32_t xm) { return xm - w; }

// C[c] + #{rows of the sector's block with in-block offset <= x that hold c}
__device__ __forceinline__ uint32_t sector_rank(const DevIndex &ix, const Sector &s, uint32_t x) {
    const uint32_t xm = (x | (x << 16)) | SWAR_M;
    if ((int32_t)s.lo.y >= 0) {
        uint32_t acc = swar_le(s.lo.y, xm) & SWAR_M;
        acc |= (swar_le(s.lo.z, xm) & SWAR_M) >> 1;
        acc |= (swar_le(s.lo.w, xm) & SWAR_M) >> 2;
        acc |= (swar_le(s.hi.x, xm) & SWAR_M) >> 3;
        acc |= (swar_le(s.hi.y, xm) & SWAR_M) >> 4;
        acc |= (swar_le(s.hi.z, xm) & SWAR_M) >> 5;
        acc |= (swar_le(s.hi.w, xm) & SWAR_M) >> 6;
        return s.lo.x + __popc(acc);
    }
    // overflow sector: 12 inline slots, then the pool
    uint32_t acc = swar_le(s.lo.y & 0x7FFFFFFFu, xm) & SWAR_M;
    acc |= (swar_le(s.lo.z, xm) & SWAR_M) >> 1;
    acc |= (swar_le(s.lo.w, xm) & SWAR_M) >> 2;
    acc |= (swar_le(s.hi.x, xm) & SWAR_M) >> 3;
    acc |= (swar_le(s.hi.y, xm) & SWAR_M) >> 4;
    acc |= (swar_le(s.hi.z, xm) & SWAR_M) >> 5;
    uint32_t cnt = __popc(acc);
    if (cnt == SECTOR_INLINE_OVF) {
        const uint16_t *p = ix.ovf + s.hi.w;
        while (true) {
            const uint32_t o = __ldg(p++);
            if (o > x) break;   // the 0xFFFF terminator compares greater than any offset
            cnt++;
        }
    }
    return s.lo.x + cnt;
}

// C[c] + #{rows <= row : bwt[row]==c}  — one 32-byte gather (two 16-byte loads of the same sector).
__device__ __forceinline__ uint32_t rank_c(const DevIndex &ix, uint32_t c, uint32_t row) {
    const Sector s = load_sector(ix, c, row >> ix.blk_shift);
    return sector_rank(ix, s, row & ((1u << ix.blk_shift) - 1));
}

__device__ __forceinline__ uint64_t ld_rec(const DevIndex &ix, uint32_t row) { return __ldg(ix.rec + row); }

}  // namespace e2fm
