// search.cu — batched exact-pattern search over the flattened device index.
//
//   encode_patterns_kernel       k-mer compact code at every pattern byte (ScrambledSuperAlphabet::getCode + charactersRemappings,
//                                EFMIndex.cpp:2186-2191), so that a backward step costs one 2-byte load.
//   K3  backward_search_kernel   EFMIndex::computeSuperPatterns + exactMatch(interval) + backwardStep
//                                (EFMIndex.cpp:2055-2207, :1901-1963, :1966-2001) for every (pattern, shift):
//                                one (pattern, shift) per LANE, lanes refilled from a warp-local slice of a global
//                                work queue; start interval + first step from the pair table, then one 32-byte rank
//                                sector per step instead of the reference's super-bucket + bucket + position-list
//                                chain (SuperBucket.cpp:82, Bucket.cpp:656-779).
//   scan / expand                the interval each shift ends with -> row tasks {row, item}
//   firstcheck_kernel            EFMIndex::backwardStepIfMatch (:2011-2041) over whole long intervals: streams bwt16[]
//   K4  resolve_kernel           row tasks answered from the dense sa[] / text[] tables: getRowPosition (:2468-2497),
//                                findWholePatternMatches + extractSuperCharacter (:1862-1899, :2216-2269)
//   K4' walk_kernel              the same row tasks by per-query walks from the marked rows (dense tables off): one
//                                dependent 8-byte gather per LF step, Bucket::getMarkedTextPosition (Bucket.cpp:823)
//   result pipeline              per-pattern CSR (scan + scatter), segmented sort + adjacent de-duplication
//                                (EFMIndex::locateOccurrences, :2281-2301), position -> (sequence, offset)
//                                (EFMCollection::search, EFMCollection.cpp:124-148).
//   extract_kernel               EFMIndex::extractSnippet (:2325-2452) from text[]
//   gather_bench_kernel          the random-gather roofline measurement
#include "search.cuh"

#include <algorithm>
#include <cstdlib>

#include "../../include/e2fm_b200.h"

namespace e2fm {

static constexpr unsigned FULL = 0xFFFFFFFFu;

// ---------------------------------------------------------------- helpers

__device__ __forceinline__ void pattern_span(const PatternView &pv, uint32_t pl, const uint8_t *&ptr, uint32_t &m) {
    const uint64_t p = pv.first + pl;
    if (pv.off) {
        const uint64_t a = __ldg(pv.off + p), b = __ldg(pv.off + p + 1);
        ptr = pv.bytes + a;
        m = (uint32_t)(b - a);
    } else {
        ptr = pv.bytes + p * pv.fixed_len;
        m = pv.fixed_len;
    }
}

// natural value (base n_sym, most significant symbol first) of `cnt` pattern bytes; 0xFFFFFFFF if a byte is
// not one of the index's original symbols (EFMIndex.cpp:1702-1717 rejects such patterns)
__device__ __forceinline__ uint32_t natural_value(const DevIndex &ix, const uint8_t *p, uint32_t cnt) {
    uint32_t v = 0;
    bool bad = false;
    for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t s = __ldg(ix.sym_index + __ldg(p + j));
        bad |= (s == 0xFFu);
        v = v * ix.n_sym + s;
    }
    return bad ? 0xFFFFFFFFu : v;
}

__device__ __forceinline__ uint32_t ipow(uint32_t b, uint32_t e) {
    uint32_t r = 1;
    for (uint32_t i = 0; i < e; i++) r *= b;
    return r;
}

// LF of a row whose record is r (marked rows keep their LF in the marked-row table)
__device__ __forceinline__ uint32_t lf_from_rec(const DevIndex &ix, uint64_t r) {
    const uint32_t low = rec_low(r);
    return rec_marked(r) ? __ldg(&ix.mark_tab[low].y) : low;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// ---------------------------------------------------------------- pattern pre-pass

// codes[j] = compact code of the k-mer that starts at byte j of the batch, 0xFFFF when it would cross the end of
// its pattern, holds a symbol outside the index's alphabet (EFMIndex.cpp:1702-1717) or never occurs in the text
// (EFMIndex.cpp:2188-2197). Every backward step then costs one 2-byte load instead of
// ScrambledSuperAlphabet::getCode + charactersRemappings (EFMIndex.cpp:2186-2191).
// One thread per 8 consecutive bytes: two 8-byte loads, symbol and k-mer tables in shared memory, one 16-byte store.
static constexpr int ENC_PER = 8;
static constexpr uint32_t ENC_SMEM_KMERS = 20480;   // k-mer table entries kept in shared memory (else read through L1)

// KT = compile-time super-alphabet order, 1..8 (0 = generic, any k, slower)
template <int KT, bool RAGGED, bool KSMEM>
__global__ void __launch_bounds__(256) encode_patterns_kernel(DevIndex ix, PatternView pv, uint64_t j_begin, uint64_t n_bytes, uint16_t *__restrict__ codes) {
    extern __shared__ uint16_t enc_smem[];
    uint8_t *sym_s = reinterpret_cast<uint8_t *>(enc_smem);            // [256]
    uint16_t *kmer_s = enc_smem + 128;                                  // [sigma_k] when it fits
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) sym_s[i] = ix.sym_index[i];
    if (KSMEM)
        for (uint32_t i = threadIdx.x; i < ix.sigma_k; i += blockDim.x) kmer_s[i] = ix.compact_of_natural[i];
    __syncthreads();
    const uint32_t k = KT ? (uint32_t)KT : ix.k, no = ix.n_sym;
    const bool aligned = (reinterpret_cast<uintptr_t>(pv.bytes) & 7) == 0;
    // persistent blocks: the tables above are filled once per CTA, not once per 2 KB of patterns
    for (uint64_t j0 = j_begin + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * ENC_PER; j0 < n_bytes;
         j0 += (uint64_t)gridDim.x * blockDim.x * ENC_PER) {
        // symbols of bytes j0 .. j0+15 (0xFF = not in the index's alphabet)
        uint32_t sy[16];
        if (aligned && j0 + 16 <= n_bytes) {
            const uint2 *q = reinterpret_cast<const uint2 *>(pv.bytes + j0);
            const uint2 lo = __ldg(q), hi = __ldg(q + 1);
            const uint32_t w[4] = {lo.x, lo.y, hi.x, hi.y};
#pragma unroll
            for (int i = 0; i < 16; i++) sy[i] = sym_s[(w[i >> 2] >> (8 * (i & 3))) & 0xFFu];
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) sy[i] = j0 + i < n_bytes ? sym_s[__ldg(pv.bytes + j0 + i)] : 0xFFu;
        }
        uint32_t badmask = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) badmask |= (sy[i] == 0xFFu ? 1u : 0u) << i;
        // bytes left in the pattern that owns byte j0 (room); patterns are shorter than 4 GB
        uint64_t pat = 0;
        uint32_t room;
        if (RAGGED) {
            uint64_t lo = pv.first, hi = pv.first + pv.count;
            while (lo + 1 < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (__ldg(pv.off + mid) <= j0) lo = mid; else hi = mid;
            }
            pat = lo;
            room = (uint32_t)(__ldg(pv.off + pat + 1) - j0);
        } else {
            const uint32_t len = pv.fixed_len;
            room = len - ((n_bytes >> 32) ? (uint32_t)(j0 % len) : (uint32_t)j0 % len);
        }
        uint32_t out[ENC_PER];
#pragma unroll
        for (int q = 0; q < ENC_PER; q++) {
            if (RAGGED) {
                while (room == 0 && j0 + q < n_bytes) {                  // step into the next pattern (empty ones included)
                    pat++;
                    room = (uint32_t)(__ldg(pv.off + pat + 1) - __ldg(pv.off + pat));
                }
            } else if (room == 0) room = pv.fixed_len;
            uint32_t c = 0xFFFFu;
            if (k <= room) {
                uint32_t nat = 0;
                bool bad;
                if (KT) {
#pragma unroll
                    for (int i = 0; i < (KT ? KT : 1); i++) nat = nat * no + sy[q + i];
                    bad = ((badmask >> q) & ((1u << KT) - 1)) != 0;
                } else {
                    bad = false;
                    for (uint32_t i = 0; i < k; i++) {
                        const uint32_t v = sym_s[__ldg(pv.bytes + j0 + q + i)];
                        bad |= v == 0xFFu;
                        nat = nat * no + v;
                    }
                }
                if (!bad) c = KSMEM ? kmer_s[nat] : __ldg(ix.compact_of_natural + nat);
            }
            out[q] = c;
            room--;
        }
        if (j0 + ENC_PER <= n_bytes) {
            uint4 v = make_uint4(out[0] | (out[1] << 16), out[2] | (out[3] << 16), out[4] | (out[5] << 16), out[6] | (out[7] << 16));
            *reinterpret_cast<uint4 *>(codes + j0) = v;
        } else {
            for (int q = 0; q < ENC_PER && j0 + q < n_bytes; q++) codes[j0 + q] = (uint16_t)out[q];
        }
    }
}

void launch_encode_patterns(const DevIndex &ix, const PatternView &pv, uint64_t j_begin, uint64_t j_end, uint64_t n_bytes, uint16_t *codes,
                            int n_sm, cudaStream_t s) {
    j_begin &= ~(uint64_t)(ENC_PER - 1);    // threads own aligned groups of 8 bytes; re-encoding a few earlier bytes is harmless
    if (j_end > n_bytes) j_end = n_bytes;
    if (j_end <= j_begin) return;
    const uint64_t threads = (j_end - j_begin + ENC_PER - 1) / ENC_PER;
    n_bytes = j_end;   // a chunk ends at a pattern end: nothing beyond it is read (later chunks may not be resident yet)
    const size_t smem = 256 + (ix.sigma_k <= ENC_SMEM_KMERS ? (size_t)ix.sigma_k * 2 : 0);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((threads + 255) / 256, (uint64_t)n_sm * 8);
    const bool ksmem = ix.sigma_k <= ENC_SMEM_KMERS, ragged = pv.off != nullptr;
#define E2FM_ENC(KT, RG, KS) encode_patterns_kernel<KT, RG, KS><<<blocks, 256, smem, s>>>(ix, pv, j_begin, n_bytes, codes)
    if (!ksmem) { if (ragged) E2FM_ENC(0, true, false); else E2FM_ENC(0, false, false); return; }
    switch (ix.k) {
        case 2: if (ragged) E2FM_ENC(2, true, true); else E2FM_ENC(2, false, true); break;
        case 3: if (ragged) E2FM_ENC(3, true, true); else E2FM_ENC(3, false, true); break;
        case 4: if (ragged) E2FM_ENC(4, true, true); else E2FM_ENC(4, false, true); break;
        case 5: if (ragged) E2FM_ENC(5, true, true); else E2FM_ENC(5, false, true); break;
        default: if (ragged) E2FM_ENC(0, true, true); else E2FM_ENC(0, false, true); break;
    }
#undef E2FM_ENC
}

// ---------------------------------------------------------------- K3: backward search

static constexpr uint32_t IV_LONG = 0x80000000u;   // iv[item].y flag: long wildcard interval, scanned by firstcheck_kernel, not expanded
static constexpr uint32_t FC_MIN_LEN = 32;  // wildcard intervals shorter than this are expanded into row tasks like any other interval
static constexpr uint32_t BS_GRAB = 256;   // items a warp takes from the global queue at a time
static constexpr int BS_BURST = 4;         // backward steps between two refills

// KT = compile-time super-alphabet order (0 = read it from the index)
// MINB = resident CTAs per SM the register allocation is tuned for: 8 (32 registers, a few spills) wins while the index is
// L2- or nearly L2-resident, 6 (40 registers, no spills) wins once every step is a DRAM gather (3.1 Gbp: 4.9 vs 5.5 ms per batch)
template <int KT, int MINB>
__global__ void __launch_bounds__(256, MINB) backward_search_kernel(SearchArgs a) {
    const DevIndex &ix = a.ix;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const uint32_t k = KT ? (uint32_t)KT : ix.k;
    const uint32_t n_items = a.pv.count * k;
    const uint32_t blk_shift = ix.blk_shift, blk_mask = (1u << blk_shift) - 1;
    uint32_t wnext = 0, wend = 0;
    bool exhausted = false, active = false, owe_iv = false;
    uint32_t item = 0, s = 0, e = 0, flags = 0;   // SA interval [s, e); flags: bit0 first super-character variable, bit1 hat
    int i = 0, lo = 0;                           // next super-character to process, and the last fixed one
    const uint16_t *cptr = nullptr;              // codes of this pattern, shifted so that character c is cptr[c*k]
    uint32_t n_rank = 0, n_sect = 0;
    

    for (;;) {
        if (wnext >= wend && !exhausted) {
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(a.stats + ST_QUEUE, (unsigned long long)BS_GRAB);
            b = __shfl_sync(FULL, b, 0);
            if (b >= n_items) exhausted = true;
            else { wnext = (uint32_t)b; wend = min((uint32_t)b + BS_GRAB, n_items); }
        }
        // ---- refill idle lanes: EFMIndex::computeSuperPatterns (:2055-2207) + start interval (:1728-1748) of one shift
        const unsigned need = __ballot_sync(FULL, !active);
        if (!active && wnext < wend) {
            const uint32_t idx = wnext + __popc(need & lt);
            if (idx < wend) {
                item = idx;
                owe_iv = true;                       // every item writes iv[item] exactly once (no memset of the array)
                const uint32_t pl = idx / k, sa = idx - pl * k;
                uint64_t base;
                uint32_t m;
                if (a.pv.off) {
                    base = __ldg(a.pv.off + a.pv.first + pl);
                    m = (uint32_t)(__ldg(a.pv.off + a.pv.first + pl + 1) - base);
                } else {
                    m = a.pv.fixed_len;
                    base = (a.pv.first + pl) * m;
                }
                if (m > 0) {
                    const uint32_t L = (m + sa + k - 1) / k;
                    const bool last_var = (m + sa) % k != 0;
                    int l = -1;
                    if (!last_var) { if (!(L == 1 && sa > 0)) l = (int)L; }              // last super-character fixed
                    else if (L > 1 && (L > 2 || sa == 0)) l = (int)L - 1;                // hat: start from its predecessor
                    // otherwise the shift yields nothing (variable last character without fixed predecessor, :2132/:1745)
                    if (l > 0) {
                        cptr = a.codes + base - sa;
                        lo = sa > 0 ? 1 : 0;
                        flags = (sa > 0 ? 1u : 0u) | (last_var ? 2u : 0u);
                        const uint32_t cc = __ldg(cptr + (size_t)(l - 1) * k);
                        // the code of the next super-character is independent of the first: both loads are in flight together
                        const uint32_t c1 = (l - 2 >= lo) ? (uint32_t)__ldg(cptr + (size_t)(l - 2) * k) : 0xFFFEu;
                        if (cc != 0xFFFFu && c1 != 0xFFFFu) {
                            if (ix.pair_tab && c1 != 0xFFFEu && !(a.split_long & 4u)) {
                                // start interval (:1728-1748) + first backwardStep (:1966) in one gather of the pair table
                                const uint2 p = __ldg(ix.pair_tab + (size_t)cc * ix.A + c1);
                                s = p.x;
                                e = p.y;
                                i = l - 3;
                                n_rank += 2;
                                n_sect++;
                            } else {
                                s = (uint32_t)__ldg(ix.C + cc);
                                e = (uint32_t)__ldg(ix.C + cc + 1);
                                i = l - 2;
                            }
                            active = e > s;
                        }
                    }
                }
            }
        }
        if (owe_iv && !active) {                     // nothing to search for this shift: an empty interval
            a.iv[item] = make_uint2(0u, 0u);
            owe_iv = false;
        }
        wnext = min(wend, wnext + (uint32_t)__popc(need));
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        // ---- EFMIndex::backwardStep (:1966-2001), a few steps per refill
#pragma unroll 1
        for (int it = 0; it < BS_BURST; it++) {
            const bool go = active && i >= lo;
            if (!__any_sync(FULL, go)) break;
            if (go) {
                const uint32_t cc = __ldg(cptr + (size_t)i * k);
                if (cc == 0xFFFFu) active = false;
                else if (e - s == 1 && !(a.split_long & 2u)) {
                    // one row left: occ(c,s) - occ(c,s-1) is 1 iff the row's BWT symbol is c, and then the new row is LF(s):
                    // one 8-byte gather and a compare instead of a rank sector and two SWAR counts
                    const uint64_t x = __ldg(ix.rec + s);
                    n_sect++;
                    n_rank += 2;
                    i--;
                    if (rec_code(x) == cc) { s = lf_from_rec(ix, x); e = s + 1; }
                    else active = false;
                } else {
                    const uint32_t re = e - 1;
                    const Sector se = load_sector(ix, cc, re >> blk_shift);
                    uint32_t ns;
                    if (s == 0) ns = (uint32_t)__ldg(ix.C + cc);
                    else {
                        const uint32_t rs = s - 1;
                        if ((rs >> blk_shift) == (re >> blk_shift)) ns = sector_rank(ix, se, rs & blk_mask);
                        else {
                            const Sector ss = load_sector(ix, cc, rs >> blk_shift);
                            ns = sector_rank(ix, ss, rs & blk_mask);
                            n_sect++;
                        }
                    }
                    e = sector_rank(ix, se, re & blk_mask);
                    s = ns;
                    n_sect++;
                    n_rank += 2;
                    i--;
                    active = e > s;
                }
            }
        }
        // ---- end of EFMIndex::exactMatch(interval) (:1901-1963): hand the surviving interval on
        if (active && i < lo) {
            const uint32_t len = e - s;
            if (a.mode == E2FM_MODE_COUNT && flags == 0) {
                atomicAdd(a.counts + a.pv.first + item / k, (unsigned long long)len);   // countOccurrences (:2273)
            } else if ((a.split_long & 1u) && (flags & 1u) && len >= FC_MIN_LEN && len < IV_LONG) {
                // '?'-prefixed first super-character and many rows: tested row by row, in row order, by firstcheck_kernel.
                // What those rows are checked against is computed here, once per (pattern, shift) instead of once per row: the
                // fixed symbols of the '?'-prefixed first and of the '?'-suffixed last super-character.
                a.iv[item] = make_uint2(s, len | IV_LONG);
                owe_iv = false;
                a.stats[ST_ANY_LONG] = 1;
                const uint32_t pl = item / k, sa = item - pl * k;
                const uint8_t *pp;
                uint32_t pm;
                pattern_span(a.pv, pl, pp, pm);
                a.firstw[item] = natural_value(ix, pp, k - sa);
                uint32_t hw = 0xFFFFFFFFu;
                if (flags & 2u) {
                    const uint32_t L = (pm + sa + k - 1) / k, from = (L - 1) * k - sa, f = pm - from;
                    const uint32_t want = natural_value(ix, pp + from, f);
                    hw = want == 0xFFFFFFFFu ? 0xFFFFFFFEu : ((f << 28) | want);     // 0xFFFFFFFE: foreign symbol, nothing matches
                }
                a.hatw[item] = hw;
            } else {
                a.iv[item] = make_uint2(s, len);
                owe_iv = false;
            }
            active = false;
        }
        if (owe_iv && !active) {                     // died, dropped out, or counted directly: an empty interval
            a.iv[item] = make_uint2(0u, 0u);
            owe_iv = false;
        }
    }
    unsigned long long r = warp_sum((unsigned long long)n_rank), q = warp_sum((unsigned long long)n_sect);
    if (lane == 0) {
        if (q) atomicAdd(a.stats + ST_SECT_BS, q);
        if (r) atomicAdd(a.stats + ST_RANK, r);
    }
}

void launch_backward_search(const SearchArgs &a, int n_sm, cudaStream_t s) {
    const uint64_t n_items = (uint64_t)a.pv.count * a.ix.k;
    if (n_items == 0) return;
    uint64_t blocks = (n_items + 255) / 256;
    const uint64_t cap = (uint64_t)n_sm * 8;   // persistent: up to 8 x 256 threads per SM
    if (blocks > cap) blocks = cap;
    static const char *ov = getenv("E2FM_K3_MINB");
    const bool big = ov ? atoi(ov) == 6 : (uint64_t)a.ix.n * 12 > (1ull << 30);
    if (a.ix.k == 4) {
        if (big) backward_search_kernel<4, 6><<<(uint32_t)blocks, 256, 0, s>>>(a);
        else backward_search_kernel<4, 8><<<(uint32_t)blocks, 256, 0, s>>>(a);
    } else {
        if (big) backward_search_kernel<0, 6><<<(uint32_t)blocks, 256, 0, s>>>(a);
        else backward_search_kernel<0, 8><<<(uint32_t)blocks, 256, 0, s>>>(a);
    }
}

// ---------------------------------------------------------------- wildcard first super-character

// EFMIndex::backwardStepIfMatch (:2011-2041) for whole intervals: when a shifted pattern starts with a '?'-prefixed
// super-character, EVERY row of the interval reached so far has to be tested against the fixed symbols of that k-mer.
// The rows are consecutive, so the test streams bwt16[] (2 bytes per row, coalesced) instead of expanding the interval
// into row tasks that each gather an 8-byte record. Rows that pass become row tasks (flagged: their position is
// sa[row] - 1), or are counted on the spot when nothing else has to be verified.
static constexpr uint32_t TASK_FIRST_DONE = 0x80000000u;

static constexpr int FC_GROUP = 2;        // warp batches (of 32 items) per task-slot reservation: same-address atomics are the
                                          // scarce resource (about one per nanosecond device-wide), so they are amortised over 256 items

// low[(f-1)*A + code] = the f low-order symbols of the k-mer of compact code `code` (its natural value mod n_sym^f): what a
// '?'-prefixed first super-character with f fixed symbols has to equal. In shared memory when it fits, else recomputed.
struct FcTable {
    const uint32_t *low;   // DevIndex::kmer_low (a few KB, L1-resident): entry (f-1)*A + code
    uint32_t A;
};

// 8-bit mask of the matching rows among the 8 consecutive rows [c, c+8) (c a multiple of 8, loaded as v) that fall in [s0, s0+ln)
__device__ __forceinline__ uint32_t fc_eval(const FcTable &tab, const uint4 v, uint32_t c, uint32_t s0, uint32_t ln, uint32_t f,
                                            uint32_t pw, uint32_t want) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const uint32_t *row = tab.low + (f - 1) * tab.A;
    uint32_t m = 0;
    (void)pw;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t code = (w[t >> 1] >> (16 * (t & 1))) & 0xFFFFu;
        m |= (__ldg(row + code) == want ? 1u : 0u) << t;
    }
    // rows of the group that belong to the interval: [max(s0,c), min(s0+ln, c+8))
    const uint32_t lo = s0 > c ? s0 - c : 0u, hi = min(s0 + ln - c, 8u);
    return m & ((1u << hi) - 1u) & ~((1u << lo) - 1u);
}
static constexpr int FC_UNROLL = 4;   // 16-byte loads in flight per lane: an interval of up to 1024 rows costs one DRAM round trip

template <int KT>
__global__ void __launch_bounds__(256) firstcheck_kernel(SearchArgs a) {
    if (a.stats[ST_ANY_LONG] == 0) return;              // no long wildcard interval in this chunk (the usual case for long patterns)
    const DevIndex &ix = a.ix;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const uint32_t k = KT ? (uint32_t)KT : ix.k, no = ix.n_sym;
    const uint32_t n_items = a.pv.count * k;
    const FcTable tab{ix.kmer_low, ix.A};
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    unsigned long long n_rows = 0;
    for (uint32_t g0 = warp_global * (32u * FC_GROUP); g0 < n_items; g0 += n_warps * (32u * FC_GROUP)) {
        // this lane's interval in each batch of the group: rows [s, s+len), f fixed symbols that must equal `want`
        uint32_t S[FC_GROUP], LEN[FC_GROUP], WANT[FC_GROUP], CNT[FC_GROUP];
        uint32_t total_mine = 0;
#pragma unroll
        for (int b = 0; b < FC_GROUP; b++) {
            const uint32_t item = g0 + b * 32 + lane;
            const uint2 v = item < n_items ? a.iv[item] : make_uint2(0u, 0u);
            S[b] = v.x;
            LEN[b] = (v.y & IV_LONG) ? (v.y & ~IV_LONG) : 0u;
        }
        // ---- pass 1: count the matching rows of every interval; the whole warp scans one interval, 256 rows per step
#pragma unroll
        for (int b = 0; b < FC_GROUP; b++) {
            const uint32_t item = g0 + b * 32 + lane;
            const uint32_t sa = item % k, f = k - sa, pw = ipow(no, f);
            uint32_t want = 0xFFFFFFFFu, len = LEN[b];
            bool count_only = false;
            if (len) {
                want = a.firstw[item];
                count_only = a.mode == E2FM_MODE_COUNT && a.hatw[item] == 0xFFFFFFFFu;
                if (want == 0xFFFFFFFFu) len = 0;                   // foreign symbol: no row can match
                n_rows += len;
            }
            uint32_t hits = 0;
            for (unsigned todo = __ballot_sync(FULL, len > 0); todo;) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t s0 = __shfl_sync(FULL, S[b], src), ln = __shfl_sync(FULL, len, src);
                const uint32_t w = __shfl_sync(FULL, want, src), ff = __shfl_sync(FULL, f, src), p = __shfl_sync(FULL, pw, src);
                uint32_t h = 0;
                for (uint32_t c0 = (s0 & ~7u) + 8 * lane; c0 < s0 + ln; c0 += 256 * FC_UNROLL) {
                    uint4 v[FC_UNROLL];
#pragma unroll
                    for (int u = 0; u < FC_UNROLL; u++)
                        v[u] = c0 + 256 * u < s0 + ln ? __ldg(reinterpret_cast<const uint4 *>(ix.bwt16 + c0 + 256 * u)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < FC_UNROLL; u++)
                        if (c0 + 256 * u < s0 + ln) h += __popc(fc_eval(tab, v[u], c0 + 256 * u, s0, ln, ff, p, w));
                }
                h = (uint32_t)warp_sum((unsigned long long)h);
                if ((int)lane == src) hits = h;
            }
            if (count_only) {
                if (hits) atomicAdd(a.counts + a.pv.first + item / k, (unsigned long long)hits);
                hits = 0;
            }
            LEN[b] = len;
            WANT[b] = want;
            CNT[b] = hits;
            total_mine += hits;
        }
        // ---- one reservation for the whole group
        uint32_t incl = total_mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, d);
            if ((int)lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.task_cursor, (unsigned long long)total);
        base = __shfl_sync(FULL, base, 0);
        unsigned long long dst = base + incl - total_mine;          // this lane's slots, its batches in order
        // ---- pass 2: write the matching rows as flagged row tasks
#pragma unroll
        for (int b = 0; b < FC_GROUP; b++) {
            const uint32_t item = g0 + b * 32 + lane;
            const uint32_t sa = item % k, f = k - sa, pw = ipow(no, f);
            for (unsigned todo = __ballot_sync(FULL, CNT[b] > 0); todo;) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t it = __shfl_sync(FULL, item, src), s0 = __shfl_sync(FULL, S[b], src), ln = __shfl_sync(FULL, LEN[b], src);
                const uint32_t w = __shfl_sync(FULL, WANT[b], src), ff = __shfl_sync(FULL, f, src), p = __shfl_sync(FULL, pw, src);
                unsigned long long d = __shfl_sync(FULL, dst, src);
                for (uint32_t c0 = (s0 & ~7u) + 8 * lane; c0 - 8 * lane < s0 + ln; c0 += 256 * FC_UNROLL) {
                    uint4 v[FC_UNROLL];
                    uint32_t m[FC_UNROLL];
#pragma unroll
                    for (int u = 0; u < FC_UNROLL; u++)
                        v[u] = c0 + 256 * u < s0 + ln ? __ldg(reinterpret_cast<const uint4 *>(ix.bwt16 + c0 + 256 * u)) : make_uint4(0, 0, 0, 0);
                    uint32_t mine = 0;
#pragma unroll
                    for (int u = 0; u < FC_UNROLL; u++) {
                        m[u] = c0 + 256 * u < s0 + ln ? fc_eval(tab, v[u], c0 + 256 * u, s0, ln, ff, p, w) : 0u;
                        mine += __popc(m[u]);
                    }
                    uint32_t inc = mine;
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, inc, dd);
                        if ((int)lane >= dd) inc += t;
                    }
                    unsigned long long slot = d + inc - mine;
#pragma unroll
                    for (int u = 0; u < FC_UNROLL; u++) {
                        uint32_t mm = m[u];
                        while (mm) {
                            const uint32_t t = __ffs(mm) - 1;
                            mm &= mm - 1;
                            if (slot < a.task_cap) a.tasks[slot] = make_uint2(c0 + 256 * u + t, it | TASK_FIRST_DONE);
                            slot++;
                        }
                    }
                    d += __shfl_sync(FULL, inc, 31);
                }
            }
            dst += CNT[b];
        }
    }
    (void)lt;
    n_rows = warp_sum(n_rows);
    if (lane == 0 && n_rows) atomicAdd(a.stats + ST_SCAN_ROWS, n_rows);
}

void launch_firstcheck(const SearchArgs &a, int n_sm, cudaStream_t s) {
    const uint64_t n_items = (uint64_t)a.pv.count * a.ix.k;
    if (n_items == 0 || a.ix.k < 2) return;
    const uint64_t groups = (n_items + 32 * FC_GROUP - 1) / (32 * FC_GROUP);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((groups + 7) / 8, (uint64_t)n_sm * 8);
    if (a.ix.k == 4) firstcheck_kernel<4><<<blocks, 256, 0, s>>>(a);
    else firstcheck_kernel<0><<<blocks, 256, 0, s>>>(a);
}

// ---------------------------------------------------------------- scans

static constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_B = SCAN_T * SCAN_E;

struct LoadIvLen { const uint2 *p; __device__ unsigned long long operator()(uint64_t i) const { const uint32_t y = p[i].y; return (y & IV_LONG) ? 0u : y; } };
struct LoadU64 { const unsigned long long *p; __device__ unsigned long long operator()(uint64_t i) const { return p[i]; } };

template <class Load>
__global__ void __launch_bounds__(SCAN_T) scan_block_sums_kernel(Load in, uint64_t n, uint64_t *sums) {
    __shared__ unsigned long long ws[SCAN_T / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_B;
    unsigned long long v = 0;
    for (int j = 0; j < SCAN_E; j++) {
        const uint64_t i = base + (uint64_t)j * SCAN_T + threadIdx.x;
        if (i < n) v += in(i);
    }
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < SCAN_T / 32; w++) t += ws[w];
        sums[blockIdx.x] = t;
    }
}

// single block: in-place exclusive scan of sums[0..nb), total to sums[nb]
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint64_t *sums, uint64_t nb) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t base = 0; base < nb; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < nb ? sums[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(FULL, inc, d);
            if ((int)lane >= d) inc += t;
        }
        if (lane == 31) ws[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = ws[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(FULL, winc, d);
                if ((int)lane >= d) winc += t;
            }
            ws[lane] = winc - w;
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        if (i < nb) sums[i] = carry + ws[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + ws[warp] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = carry_s;
}

template <class Load>
__global__ void __launch_bounds__(SCAN_T) scan_apply_kernel(Load in, uint64_t n, const uint64_t *sums, uint64_t nb, uint64_t *out) {
    __shared__ unsigned long long ws[SCAN_T / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_B + (uint64_t)threadIdx.x * SCAN_E;
    unsigned long long v[SCAN_E], tot = 0;
#pragma unroll
    for (int j = 0; j < SCAN_E; j++) {
        v[j] = base + j < n ? in(base + j) : 0;
        tot += v[j];
    }
    unsigned long long inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(FULL, inc, d);
        if ((int)lane >= d) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    unsigned long long run = sums[blockIdx.x] + inc - tot;
    for (uint32_t w = 0; w < warp; w++) run += ws[w];
#pragma unroll
    for (int j = 0; j < SCAN_E; j++) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[nb];
}

// Fast path: the apply pass of the interval scan writes the row tasks directly (no offset array, no separate expansion
// kernel). Thread t owns SCAN_E consecutive (pattern, shift) items; its first task slot is the scanned prefix. The chunk's task
// count goes to *n_tasks_out; when it exceeds `cap` nothing is written and the host redoes the chunk through the windowed path.
__global__ void __launch_bounds__(SCAN_T) scan_expand_kernel(const uint2 *__restrict__ iv, uint64_t n, const uint64_t *sums, uint64_t nb,
                                                             uint2 *__restrict__ tasks, uint64_t cap, unsigned long long *n_tasks_out) {
    __shared__ unsigned long long ws[SCAN_T / 32], red[2][SCAN_T / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // prefix of this block and the chunk's total, straight from the block sums (a few thousand at most): no separate launch
    unsigned long long pre = 0, all = 0;
    for (uint64_t i = threadIdx.x; i < nb; i += SCAN_T) {
        const unsigned long long v = sums[i];
        all += v;
        if (i < blockIdx.x) pre += v;
    }
    pre = warp_sum(pre);
    all = warp_sum(all);
    if (lane == 0) { red[0][warp] = pre; red[1][warp] = all; }
    __syncthreads();
    unsigned long long block_prefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_T / 32; w++) { block_prefix += red[0][w]; total += red[1][w]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_tasks_out = total;
    if (total > cap) return;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_B + (uint64_t)threadIdx.x * SCAN_E;
    uint2 it[SCAN_E];
    unsigned long long tot = 0;
#pragma unroll
    for (int j = 0; j < SCAN_E; j++) {
        it[j] = base + j < n ? iv[base + j] : make_uint2(0u, 0u);
        if (it[j].y & IV_LONG) it[j].y = 0;         // long wildcard intervals are scanned by firstcheck_kernel
        tot += it[j].y;
    }
    unsigned long long inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(FULL, inc, d);
        if ((int)lane >= d) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    unsigned long long run = block_prefix + inc - tot;
    for (uint32_t w = 0; w < warp; w++) run += ws[w];
#pragma unroll
    for (int j = 0; j < SCAN_E; j++) {
        const uint32_t len = it[j].y, item = (uint32_t)(base + j);
        if (len && len <= 8)
            for (uint32_t q = 0; q < len; q++) tasks[run + q] = make_uint2(it[j].x + q, item);
        unsigned mask = __ballot_sync(FULL, len > 8);
        while (mask) {                              // long intervals are written by the whole warp
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const uint32_t r0 = __shfl_sync(FULL, it[j].x, src), ln = __shfl_sync(FULL, len, src), im = __shfl_sync(FULL, item, src);
            const unsigned long long d = __shfl_sync(FULL, run, src);
            for (uint32_t q = lane; q < ln; q += 32) tasks[d + q] = make_uint2(r0 + q, im);
        }
        run += len;
    }
}

void launch_scan_expand(const uint2 *iv, uint32_t n_items, uint64_t *scratch, uint2 *tasks, uint64_t cap, unsigned long long *n_tasks_out,
                        cudaStream_t s) {
    if (n_items == 0) return;
    const uint64_t nb = ((uint64_t)n_items + SCAN_B - 1) / SCAN_B;
    scan_block_sums_kernel<<<(uint32_t)nb, SCAN_T, 0, s>>>(LoadIvLen{iv}, (uint64_t)n_items, scratch);
    scan_expand_kernel<<<(uint32_t)nb, SCAN_T, 0, s>>>(iv, (uint64_t)n_items, scratch, nb, tasks, cap, n_tasks_out);
}

size_t scan_scratch_elems(uint64_t n) { return (size_t)((n + SCAN_B - 1) / SCAN_B + 2); }

template <class Load>
static void scan_impl(Load in, uint64_t *off, uint64_t n, uint64_t *scratch, cudaStream_t s) {
    const uint64_t nb = (n + SCAN_B - 1) / SCAN_B;
    if (n == 0) {
        cudaMemsetAsync(off, 0, 8, s);
        return;
    }
    scan_block_sums_kernel<<<(uint32_t)nb, SCAN_T, 0, s>>>(in, n, scratch);
    scan_sums_kernel<<<1, 1024, 0, s>>>(scratch, nb);
    scan_apply_kernel<<<(uint32_t)nb, SCAN_T, 0, s>>>(in, n, scratch, nb, off);
}

void launch_task_scan(const uint2 *iv, uint64_t *off, uint32_t n_items, uint64_t *scratch, cudaStream_t s) {
    scan_impl(LoadIvLen{iv}, off, n_items, scratch, s);
}
void launch_scan_u64(const unsigned long long *in, uint64_t *off, uint64_t n, uint64_t *scratch, cudaStream_t s) {
    scan_impl(LoadU64{in}, off, n, scratch, s);
}

// ---------------------------------------------------------------- interval expansion

// one (pattern, shift) interval per lane; rows [sp, sp+len) become row tasks at off[item] - w0 (clipped to the
// task window [w0, w1)). Long intervals are written by the whole warp.
__global__ void __launch_bounds__(256) expand_kernel(const uint2 *__restrict__ iv, const uint64_t *__restrict__ off, uint32_t n_items,
                                                      uint64_t w0, uint64_t w1, uint2 *__restrict__ tasks, unsigned long long *n_tasks_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_tasks_out) {
        // fast path (no host round trip): the window is [0, capacity); publish the chunk's task count and write nothing
        // when it does not fit — the host then redoes the chunk through the windowed path
        const uint64_t total = off[n_items];
        if (item == 0) *n_tasks_out = total;
        if (total > w1) return;
    }
    uint32_t row0 = 0, len = 0;
    uint64_t dst = 0;
    if (item < n_items) {
        uint2 v = iv[item];
        if (v.y & IV_LONG) v.y = 0;                 // long wildcard intervals are scanned by firstcheck_kernel
        uint64_t a = off[item], b = a + v.y;
        const uint64_t ca = a < w0 ? w0 : a, cb = b > w1 ? w1 : b;
        if (cb > ca) {
            row0 = v.x + (uint32_t)(ca - a);
            len = (uint32_t)(cb - ca);
            dst = ca - w0;
        }
    }
    const bool longer = len > 8;
    if (!longer)
        for (uint32_t q = 0; q < len; q++) tasks[dst + q] = make_uint2(row0 + q, item);
    unsigned mask = __ballot_sync(FULL, longer);
    while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t r0 = __shfl_sync(FULL, row0, src), ln = __shfl_sync(FULL, len, src), it = __shfl_sync(FULL, item, src);
        const uint64_t d = __shfl_sync(FULL, dst, src);
        for (uint32_t q = lane; q < ln; q += 32) tasks[d + q] = make_uint2(r0 + q, it);
    }
}

void launch_expand(const uint2 *iv, const uint64_t *off, uint32_t n_items, uint64_t w0, uint64_t w1, uint2 *tasks,
                   unsigned long long *n_tasks_out, cudaStream_t s) {
    if (n_items == 0 || w1 <= w0) return;
    expand_kernel<<<(n_items + 255) / 256, 256, 0, s>>>(iv, off, n_items, w0, w1, tasks, n_tasks_out);
}

// ---------------------------------------------------------------- K4: row walks

// phases of one row task; WALK and HAT_WALK are the hot ones (one dependent 8-byte gather per step, nothing else)
enum : uint32_t { PH_FIRST = 0, PH_WALK = 1, PH_ATMARK = 2, PH_HAT_START = 3, PH_HAT_WALK = 4, PH_HAT_FINAL = 5 };
static constexpr uint32_t WALK_GRAB = 128;
static constexpr int WALK_BURST = 8;       // LF steps between two passes over the rare phases / refills

template <int KT>
__global__ void __launch_bounds__(256) walk_kernel(WalkArgs a) {
    const DevIndex &ix = a.ix;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const uint32_t k = KT ? (uint32_t)KT : ix.k, no = ix.n_sym;
    const uint32_t rate = ix.rate;
    const uint64_t *__restrict__ rec = ix.rec;
    const uint64_t *mark_tab64 = reinterpret_cast<const uint64_t *>(ix.mark_tab);
    uint32_t wnext = 0, wend = 0;
    bool exhausted = false, active = false;
    uint32_t phase = 0, row = 0, t = 0, item = 0, tp = 0;
    uint32_t n_lf = 0, n_mark = 0, n_gather = 0;
    if (a.n_tasks_dev) {                                   // fast path: the count lives on the device (expand_kernel)
        const unsigned long long nt = *a.n_tasks_dev;
        a.n_tasks = nt > a.n_tasks ? 0u : (uint32_t)nt;   // a.n_tasks holds the capacity; an overflowing chunk is redone by the host
    }

    for (;;) {
        if (wnext >= wend && !exhausted) {
            unsigned long long b = 0;
            if (lane == 0) b = atomicAdd(a.stats + ST_QUEUE, (unsigned long long)WALK_GRAB);
            b = __shfl_sync(FULL, b, 0);
            if (b >= a.n_tasks) exhausted = true;
            else { wnext = (uint32_t)b; wend = min((uint32_t)b + WALK_GRAB, a.n_tasks); }
        }
        const unsigned need = __ballot_sync(FULL, !active);
        if (!active && wnext < wend) {
            const uint32_t idx = wnext + __popc(need & lt);
            if (idx < wend) {
                const uint2 tk = __ldg(a.tasks + idx);
                row = tk.x;
                item = tk.y;
                phase = (item % k) ? PH_FIRST : PH_WALK;   // a shifted pattern starts with a '?'-prefixed super-character
                t = 0;
                active = true;
            }
        }
        wnext = min(wend, wnext + (uint32_t)__popc(need));
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        // ---- rare phases: at most one gather per lane per pass
        bool emit = false;
        if (active && phase != PH_WALK && phase != PH_HAT_WALK) {
            const uint32_t pl = item / k, sa = item - pl * k;
            const uint8_t *ptr;
            uint32_t m;
            pattern_span(a.pv, pl, ptr, m);
            const bool hat = (m + sa) % k != 0;
            if (phase == PH_ATMARK) {
                // ---- EFMIndex::getRowPosition (:2468-2497) reached a marked row: `row` holds its marked value
                n_mark += t + 1;
                n_lf += t;
                n_gather += t + 1;
                tp = (uint32_t)(((uint64_t)row * rate + t) % ix.n);
                if (hat) {
                    // ---- EFMIndex::extractSuperCharacter (:2216-2269): start from the next marked row
                    const uint32_t L = (m + sa + k - 1) / k;
                    const uint64_t p = (uint64_t)tp + L - 1;
                    const uint64_t nv = p / rate + 1;
                    uint64_t nxt = nv * rate;
                    if (nxt > ix.n - 1) nxt = ix.n;
                    row = (uint32_t)(nv % ix.mrn);          // marked value whose row is needed
                    t = (uint32_t)(nxt - p - 1);            // LF steps to take from it
                    n_lf += t + 1;
                    n_gather += t ? t + 1 : 2;
                    phase = PH_HAT_START;
                } else {
                    emit = true;
                    active = false;
                }
            }
            if (active) {
                if (phase == PH_FIRST) {
                    // ---- EFMIndex::backwardStepIfMatch (:2011-2041): '?'-prefixed first super-character
                    const uint64_t x = __ldg(rec + row);
                    n_lf++;
                    n_gather++;
                    const uint32_t f = k - sa;                      // fixed symbols at the low end of the k-mer
                    const uint32_t want = natural_value(ix, ptr, f);
                    const uint32_t nat = __ldg(ix.natural_of_compact + rec_code(x));
                    if (want != 0xFFFFFFFFu && nat % ipow(no, f) == want) {
                        row = lf_from_rec(ix, x);
                        if (a.mode == E2FM_MODE_COUNT && !hat) {
                            emit = true;
                            active = false;
                        } else {
                            phase = PH_WALK;
                            t = 0;
                        }
                    } else active = false;
                } else if (phase == PH_HAT_START) {
                    const uint64_t x = __ldg(mark_tab64 + row);
                    const uint32_t mrow = (uint32_t)x, mlf = (uint32_t)(x >> 32);
                    if (mrow >= ix.n) active = false;               // no such marked value (corrupt index)
                    else if (t == 0) { row = mrow; phase = PH_HAT_FINAL; }
                    else {
                        row = mlf;
                        t--;
                        phase = t ? PH_HAT_WALK : PH_HAT_FINAL;
                    }
                } else if (phase == PH_HAT_FINAL) {
                    // ---- EFMIndex::findWholePatternMatches (:1862-1899): compare the '?'-suffixed last super-character
                    const uint64_t x = __ldg(rec + row);
                    const uint32_t L = (m + sa + k - 1) / k;
                    const uint32_t from = (L - 1) * k - sa;         // first pattern byte of the last super-character
                    const uint32_t f = m - from;                    // fixed symbols at the high end of the k-mer
                    const uint32_t want = natural_value(ix, ptr + from, f);
                    const uint32_t nat = __ldg(ix.natural_of_compact + rec_code(x));
                    emit = want != 0xFFFFFFFFu && nat / ipow(no, k - f) == want;
                    active = false;
                }
            }
        }
        // ---- warp-aggregated result compaction
        const unsigned em = __ballot_sync(FULL, emit);
        if (em) {
            const uint32_t pl = item / k, sa = item - pl * k;
            if (emit) atomicAdd(a.counts + a.pv.first + pl, 1ull);
            if (a.mode == E2FM_MODE_LOCATE) {
                unsigned long long base = 0;
                const int leader = __ffs(em) - 1;
                if ((int)lane == leader) base = atomicAdd(a.stats + ST_RESULTS, (unsigned long long)__popc(em));
                base = __shfl_sync(FULL, base, leader);
                if (emit) {
                    const unsigned long long slot = base + __popc(em & lt);
                    if (slot < a.res_cap) {
                        a.res_pat[slot] = (uint32_t)(a.pv.first + pl);
                        a.res_pos[slot] = (uint64_t)tp * k + sa + ix.initial_n;   // EFMIndex.cpp:2312-2319
                    } else atomicAdd(a.stats + ST_RES_OVF, 1ull);
                }
            }
        }
        // ---- hot phases: a burst of LF steps, one dependent gather each
#pragma unroll 1
        for (int it = 0; it < WALK_BURST; it++) {
            const bool fw = active && phase == PH_WALK, fh = active && phase == PH_HAT_WALK;
            if (!__any_sync(FULL, fw || fh)) break;
            if (fw || fh) {
                const uint64_t x = __ldg(rec + row);
                const uint32_t low = rec_low(x);
                if (fw) {
                    row = low;                                      // LF(row), or the marked value when the walk ends
                    if (rec_marked(x)) phase = PH_ATMARK;
                    else if (++t > rate) active = false;            // a valid index marks every rate-th text position
                } else {
                    row = rec_marked(x) ? __ldg(&ix.mark_tab[low].y) : low;
                    if (--t == 0) phase = PH_HAT_FINAL;
                }
            }
        }
    }
    const unsigned long long l = warp_sum((unsigned long long)n_lf), mk = warp_sum((unsigned long long)n_mark),
                             g = warp_sum((unsigned long long)n_gather);
    if (lane == 0) {
        if (g) atomicAdd(a.stats + ST_GATHER, g);
        if (l) atomicAdd(a.stats + ST_LF, l);
        if (mk) atomicAdd(a.stats + ST_MARK, mk);
    }
}

void launch_walk(const WalkArgs &a, int n_sm, cudaStream_t s) {
    if (a.n_tasks == 0) return;
    uint64_t blocks = a.n_tasks_dev ? ~0ull : ((uint64_t)a.n_tasks + 255) / 256;
    const uint64_t cap = (uint64_t)n_sm * 8;
    if (blocks > cap) blocks = cap;
    if (a.ix.k == 4) walk_kernel<4><<<(uint32_t)blocks, 256, 0, s>>>(a);
    else walk_kernel<0><<<(uint32_t)blocks, 256, 0, s>>>(a);
}

// ---------------------------------------------------------------- K4 (dense tables): resolve row tasks without walking

// Same row tasks as walk_kernel, answered from the dense tables built at load (device_index.cuh): the position of a
// row is sa[row] (what EFMIndex::getRowPosition walks to), the super-character at a text position is text[p] (what
// EFMIndex::extractSuperCharacter walks to). 1-3 independent gathers per task, one task per thread. What a row has to
// match was computed once per (pattern, shift) by the backward search (firstw / hatw); the per-symbol parts of every
// k-mer come from two small shared-memory tables, so a task touches no pattern bytes.
static constexpr uint32_t RS_FLUSH = 1024, RS_CAP = RS_FLUSH + 256;   // CTA staging buffer of resolve_kernel

template <int KT>
__global__ void __launch_bounds__(256) resolve_kernel(WalkArgs a) {
    const DevIndex &ix = a.ix;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const uint32_t k = KT ? (uint32_t)KT : ix.k, no = ix.n_sym, A = ix.A;
    const uint32_t *low = ix.kmer_low, *high = ix.kmer_high;   // (f-1)*A + code; a few KB, L1-resident
    if (a.n_tasks_dev) {
        const unsigned long long nt = *a.n_tasks_dev;
        a.n_tasks = nt > a.n_tasks ? 0u : (uint32_t)nt;
    }
    uint32_t n_g = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    // Located occurrences are staged per CTA in shared memory and handed over RS_FLUSH at a time: a same-address returning
    // atomic per warp iteration (10^5..10^7 per launch) serialises at about one per nanosecond and was most of this kernel.
    __shared__ uint32_t rs_pat[RS_CAP];
    __shared__ uint64_t rs_pos[RS_CAP];
    __shared__ uint32_t rs_cnt;
    __shared__ unsigned long long rs_base;
    const bool locate = a.mode == E2FM_MODE_LOCATE;
    if (threadIdx.x == 0) rs_cnt = 0;
    __syncthreads();
    auto flush = [&]() {                                   // called by the whole CTA, rs_cnt stable
        const uint32_t c = rs_cnt;
        if (threadIdx.x == 0) rs_base = atomicAdd(a.stats + ST_RESULTS, (unsigned long long)c);
        __syncthreads();
        const unsigned long long b0 = rs_base;
        uint32_t lost = 0;
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
            const unsigned long long slot = b0 + i;
            if (slot < a.res_cap) { a.res_pat[slot] = rs_pat[i]; a.res_pos[slot] = rs_pos[i]; }
            else lost++;
        }
        if (lost) atomicAdd(a.stats + ST_RES_OVF, (unsigned long long)lost);
        __syncthreads();
        if (threadIdx.x == 0) rs_cnt = 0;
        __syncthreads();
    };
    // the whole CTA iterates together (uniform trip count: the staging buffer is flushed between iterations)
    for (uint32_t base = blockIdx.x * blockDim.x; base < a.n_tasks; base += stride) {
        const uint32_t idx = base + threadIdx.x;
        bool emit = false;
        uint32_t item = 0, tp = 0;
        if (idx < a.n_tasks) {
            const uint2 tk = __ldg(a.tasks + idx);
            const uint32_t row = tk.x;
            const bool first_done = (tk.y & 0x80000000u) != 0;   // firstcheck_kernel already matched the row's BWT symbol
            item = tk.y & 0x7FFFFFFFu;
            const uint32_t pl = item / k, sa = item - pl * k;
            const uint8_t *ptr;
            uint32_t m;
            pattern_span(a.pv, pl, ptr, m);
            const bool hat = (m + sa) % k != 0;
            const bool need_tp = hat || a.mode == E2FM_MODE_LOCATE;
            const uint32_t L = (m + sa + k - 1) / k;
            bool ok = true;
            uint32_t hat_f = 0, hat_want = 0;
            if (first_done) {
                // rows of a long wildcard interval: their checks were prepared per (pattern, shift) by the backward search
                const uint32_t hw = __ldg(a.hatw + item);
                hat_f = hw >> 28;
                hat_want = hw & 0x0FFFFFFFu;
                if (hw == 0xFFFFFFFEu) ok = false;
            } else {
                if (sa > 0) {
                    // EFMIndex::backwardStepIfMatch (:2011-2041): the row's BWT character against the '?'-prefixed first one
                    const uint32_t code = rec_code(__ldg(ix.rec + row));
                    n_g++;
                    const uint32_t f = k - sa, want = natural_value(ix, ptr, f);
                    ok = want != 0xFFFFFFFFu &&
                         __ldg(low + (f - 1) * A + code) == want;
                }
                if (ok && hat) {
                    const uint32_t from = (L - 1) * k - sa;
                    hat_f = m - from;
                    hat_want = natural_value(ix, ptr + from, hat_f);
                    if (hat_want == 0xFFFFFFFFu) ok = false;
                }
            }
            if (ok && need_tp) {
                tp = __ldg(ix.sa + row);                            // EFMIndex::getRowPosition (:2468-2497)
                n_g++;
                // a first-character match stands for the suffix one super-character earlier
                if (sa > 0) tp = tp == 0 ? (uint32_t)(ix.n - 1) : tp - 1;
            }
            if (ok && hat) {
                // EFMIndex::findWholePatternMatches (:1862-1899) / extractSuperCharacter (:2216-2269)
                const uint64_t p = (uint64_t)tp + L - 1;
                ok = false;
                if (p < ix.n) {
                    const uint32_t code = __ldg(ix.text + p);
                    n_g++;
                    ok = __ldg(high + (hat_f - 1) * A + code) == hat_want;
                }
            }
            emit = ok;
        }
        const unsigned em = __ballot_sync(FULL, emit);
        if (em) {
            const uint32_t pl = item / k, sa = item - pl * k;
            if (emit) atomicAdd(a.counts + a.pv.first + pl, 1ull);
            if (locate) {
                uint32_t b = 0;
                const int leader = __ffs(em) - 1;
                if ((int)lane == leader) b = atomicAdd(&rs_cnt, (uint32_t)__popc(em));      // shared-memory atomic
                b = __shfl_sync(FULL, b, leader);
                if (emit) {
                    const uint32_t slot = b + __popc(em & lt);                               // < RS_FLUSH + 256 = RS_CAP
                    rs_pat[slot] = (uint32_t)(a.pv.first + pl);
                    rs_pos[slot] = (uint64_t)tp * k + sa + ix.initial_n;
                }
            }
        }
        // The decision is taken THROUGH the barrier: a thread that read rs_cnt after it could see the next iteration's
        // additions of a faster warp and flush alone. The last read before the barrier sees every addition of this iteration.
        if (locate && __syncthreads_or(rs_cnt >= RS_FLUSH)) flush();
    }
    if (locate) {
        __syncthreads();
        if (rs_cnt) flush();
    }
    const unsigned long long g = warp_sum((unsigned long long)n_g);
    if (lane == 0 && g) atomicAdd(a.stats + ST_GATHER, g);
}

void launch_resolve(const WalkArgs &a, int n_sm, cudaStream_t s) {
    if (a.n_tasks == 0) return;
    uint64_t blocks = a.n_tasks_dev ? ~0ull : ((uint64_t)a.n_tasks + 255) / 256;
    const uint64_t cap = (uint64_t)n_sm * 16;
    if (blocks > cap) blocks = cap;
    if (a.ix.k == 4) resolve_kernel<4><<<(uint32_t)blocks, 256, 0, s>>>(a);
    else resolve_kernel<0><<<(uint32_t)blocks, 256, 0, s>>>(a);
}

// ---------------------------------------------------------------- result pipeline

__global__ void __launch_bounds__(256) scatter_kernel(const uint32_t *__restrict__ res_pat, const uint64_t *__restrict__ res_pos, uint64_t n_res,
                                                       const uint64_t *__restrict__ off, uint32_t *__restrict__ cursor,
                                                       uint64_t *__restrict__ positions) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_res) return;
    const uint32_t p = res_pat[i];
    const uint64_t slot = off[p] + atomicAdd(cursor + p, 1u);
    positions[slot] = res_pos[i];
}

void launch_scatter(const uint32_t *res_pat, const uint64_t *res_pos, uint64_t n_res, const uint64_t *off, uint32_t *cursor,
                    uint64_t *positions, cudaStream_t s) {
    if (n_res == 0) return;
    scatter_kernel<<<(uint32_t)((n_res + 255) / 256), 256, 0, s>>>(res_pat, res_pos, n_res, off, cursor, positions);
}

static constexpr uint32_t SORT_SMALL = 24;     // segments up to this length are sorted by one thread
static constexpr uint32_t SORT_SMEM = 4096;    // ... up to this length by one CTA in shared memory

// adjacent de-duplication of a sorted segment (EFMIndex.cpp:2296-2299); returns the unique length
__device__ __forceinline__ uint32_t unique_serial(uint64_t *p, uint32_t len) {
    uint32_t w = 0;
    for (uint32_t i = 0; i < len; i++)
        if (i == 0 || p[i] != p[w - 1]) p[w++] = p[i];
    return w;
}

__global__ void __launch_bounds__(256) segment_sort_small_kernel(uint64_t *positions, const uint64_t *__restrict__ off, uint64_t n_pat,
                                                                  uint32_t *big_list, uint32_t *ulen, unsigned long long *stats) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pat) return;
    const uint64_t a = off[p], b = off[p + 1];
    const uint64_t len = b - a;
    if (len <= 1) { ulen[p] = (uint32_t)len; return; }
    if (len > SORT_SMALL) {
        const unsigned long long at = atomicAdd(stats + ST_BIGSEG, 1ull);
        big_list[at] = (uint32_t)p;
        return;
    }
    uint64_t *q = positions + a;
    for (uint32_t i = 1; i < (uint32_t)len; i++) {          // insertion sort (std::sort, EFMIndex.cpp:2291)
        const uint64_t v = q[i];
        uint32_t j = i;
        while (j > 0 && q[j - 1] > v) { q[j] = q[j - 1]; j--; }
        q[j] = v;
    }
    const uint32_t u = unique_serial(q, (uint32_t)len);
    ulen[p] = u;
    if (u != len) atomicAdd(stats + ST_DUPS, (unsigned long long)(len - u));
}

void launch_segment_sort(uint64_t *positions, const uint64_t *off, uint64_t n_pat, uint32_t *big_list, uint32_t *ulen,
                         unsigned long long *stats, cudaStream_t s) {
    if (n_pat == 0) return;
    segment_sort_small_kernel<<<(uint32_t)((n_pat + 255) / 256), 256, 0, s>>>(positions, off, n_pat, big_list, ulen, stats);
}

// ascending-only bitonic network (mirror step + half-cleaners): entries at index >= len behave as +infinity, so no
// padding storage is needed. T indexes either shared or global memory.
template <class Arr>
__device__ __forceinline__ void bitonic_sort_block(Arr arr, uint32_t len) {
    uint32_t N = 1;
    while (N < len) N <<= 1;
    for (uint32_t kk = 2; kk <= N; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < N / 2; t += blockDim.x) {
                uint32_t i, l;
                if (j == (kk >> 1)) {
                    const uint32_t blk = t / j, pos = t - blk * j;
                    i = blk * kk + pos;
                    l = blk * kk + kk - 1 - pos;
                } else {
                    i = 2 * j * (t / j) + (t % j);
                    l = i + j;
                }
                if (l < len) {
                    const uint64_t x = arr[i], y = arr[l];
                    if (x > y) { arr[i] = y; arr[l] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256) segment_sort_big_kernel(uint64_t *positions, const uint64_t *__restrict__ off,
                                                                const uint32_t *__restrict__ big_list, uint32_t *ulen,
                                                                unsigned long long *stats) {
    __shared__ uint64_t sm[SORT_SMEM];
    __shared__ uint32_t u_s;
    const uint32_t p = big_list[blockIdx.x];
    const uint64_t a = off[p];
    const uint64_t len64 = off[p + 1] - a;
    const uint32_t len = (uint32_t)len64;
    uint64_t *q = positions + a;
    if (len <= SORT_SMEM) {
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) sm[i] = q[i];
        __syncthreads();
        bitonic_sort_block(sm, len);
        for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) q[i] = sm[i];
    } else {
        __syncthreads();
        bitonic_sort_block(q, len);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u_s = unique_serial(q, len);
        ulen[p] = u_s;
        if (u_s != len) atomicAdd(stats + ST_DUPS, (unsigned long long)(len - u_s));
    }
}

void launch_segment_sort_big(uint64_t *positions, const uint64_t *off, const uint32_t *big_list, uint32_t n_big, uint32_t *ulen,
                             unsigned long long *stats, cudaStream_t s) {
    if (n_big == 0) return;
    segment_sort_big_kernel<<<n_big, 256, 0, s>>>(positions, off, big_list, ulen, stats);
}

__global__ void __launch_bounds__(256) ulen_widen_kernel(const uint32_t *__restrict__ ulen, unsigned long long *__restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ulen[i];
}
void launch_ulen_widen(const uint32_t *ulen, unsigned long long *out, uint64_t n, cudaStream_t s) {
    if (n) ulen_widen_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(ulen, out, n);
}

// one warp per pattern: copy the unique head of its old segment to its new place
__global__ void __launch_bounds__(256) repack_kernel(const uint64_t *__restrict__ old_pos, const uint64_t *__restrict__ old_off,
                                                      const uint64_t *__restrict__ new_off, uint64_t n_pat, uint64_t *__restrict__ new_pos) {
    const uint64_t p = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (p >= n_pat) return;
    const uint64_t src = old_off[p], dst = new_off[p], len = new_off[p + 1] - dst;
    for (uint64_t i = lane; i < len; i += 32) new_pos[dst + i] = old_pos[src + i];
}
void launch_repack(const uint64_t *old_pos, const uint64_t *old_off, const uint64_t *new_off, uint64_t n_pat, uint64_t *new_pos,
                   cudaStream_t s) {
    if (n_pat) repack_kernel<<<(uint32_t)((n_pat * 32 + 255) / 256), 256, 0, s>>>(old_pos, old_off, new_off, n_pat, new_pos);
}

// EFMCollection::search (EFMCollection.cpp:124-148): sequence = first terminator beyond the position
__global__ void __launch_bounds__(256) map_positions_kernel(DevIndex ix, const uint64_t *__restrict__ positions, uint64_t n_pos,
                                                             uint32_t *__restrict__ seq, uint64_t *__restrict__ off) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pos) return;
    const uint64_t pos = positions[i];
    uint32_t lo = 0, hi = ix.n_term;                       // first index with terminators[idx] > pos
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)__ldg(ix.terminators + mid) > pos) hi = mid;
        else lo = mid + 1;
    }
    const bool found = lo < ix.n_term;
    seq[i] = found ? lo : 0xFFFFFFFFu;                     // the reference leaves sequenceIndex = -1
    uint64_t start = 0;
    if (found && lo > 0) start = (uint64_t)__ldg(ix.terminators + lo - 1) + ix.k;
    off[i] = pos - start;
}
void launch_map_positions(const DevIndex &ix, const uint64_t *positions, uint64_t n_pos, uint32_t *seq, uint64_t *off, cudaStream_t s) {
    if (n_pos) map_positions_kernel<<<(uint32_t)((n_pos + 255) / 256), 256, 0, s>>>(ix, positions, n_pos, seq, off);
}

// ---------------------------------------------------------------- extract

// One thread per super-character: text[p] -> natural value -> k symbols (ScrambledSuperAlphabet::getSymbol,
// ScrambledSuperAlphabet.cpp:90-112). The reference walks LF from the next marked row for every snippet (EFMIndex.cpp:2367-2440).
__global__ void __launch_bounds__(256) extract_kernel(DevIndex ix, uint64_t st_from, uint64_t count, uint8_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t nat = __ldg(ix.natural_of_compact + __ldg(ix.text + st_from + i));
    for (int j = (int)ix.k - 1; j >= 0; j--) {
        out[i * ix.k + j] = __ldg(ix.symbols + nat % ix.n_sym);
        nat /= ix.n_sym;
    }
}
void launch_extract(const DevIndex &ix, uint64_t st_from, uint64_t count, uint8_t *out, cudaStream_t s) {
    if (count) extract_kernel<<<(uint32_t)((count + 255) / 256), 256, 0, s>>>(ix, st_from, count, out);
}

// ---------------------------------------------------------------- random-gather microbenchmark

// The roofline this path lives under is not the streaming copy bandwidth but what HBM delivers for independent random
// sector gathers. Every thread issues `per_thread` independent 8-byte loads at hashed addresses of a large buffer
// (the access pattern of the LF / sa / text gathers, minus the dependency between steps).
// `variant` picks the load instruction (what a gather costs in DRAM traffic depends on it, profiles/exp/):
//   0 ld.global.nc.u64 (__ldg)   1 ld.global.u64   2 ld.global.cg.u64 (no L1)   3 ld.global.nc.v8.u32 (one 32-byte load)
//   4 two ld.global.nc.v4.u32 of one sector   5 ld.global.nc.L2::64B.u64   6 ld.global.nc.L1::no_allocate.u64   7 ld.global.nc.L2::128B.u64
//   8 ld.global.nc.L2::64B.v8.u32 (ld_sector: what the search kernels use)
template <int V>
__device__ __forceinline__ unsigned long long gather_load(const uint64_t *p) {
    unsigned long long v = 0;
    if (V == 0) v = __ldg(p);
    else if (V == 1) asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (V == 2) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (V == 3) {
        uint32_t w0, w1, w2, w3, w4, w5, w6, w7;
        const uint64_t *q = reinterpret_cast<const uint64_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)31);
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7) : "l"(q));
        v = (unsigned long long)(w0 ^ w1 ^ w2 ^ w3) | ((unsigned long long)(w4 ^ w5 ^ w6 ^ w7) << 32);
    } else if (V == 4) {
        const uint4 *q = reinterpret_cast<const uint4 *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)31);
        const uint4 a = __ldg(q), b = __ldg(q + 1);
        v = (unsigned long long)(a.x ^ a.y ^ a.z ^ a.w) | ((unsigned long long)(b.x ^ b.y ^ b.z ^ b.w) << 32);
    } else if (V == 5) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (V == 6) asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (V == 7) asm volatile("ld.global.nc.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (V == 8) {
        const Sector s = ld_sector(reinterpret_cast<const uint4 *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)31));
        v = (unsigned long long)(s.lo.x ^ s.lo.y ^ s.lo.z ^ s.lo.w) | ((unsigned long long)(s.hi.x ^ s.hi.y ^ s.hi.z ^ s.hi.w) << 32);
    }
    return v;
}

template <int V>
__global__ void __launch_bounds__(256) gather_bench_kernel(const uint64_t *__restrict__ buf, uint64_t n_words, uint32_t per_thread,
                                                            unsigned long long *sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    uint64_t x = t * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    const uint64_t mask = n_words - 1;                             // n_words is a power of two
#pragma unroll 8
    for (uint32_t i = 0; i < per_thread; i++) {
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;    // splitmix-style scramble
        acc += gather_load<V>(buf + (x & mask));
    }
    if (acc == 0x123456789ull) *sink = acc;                        // keep the loads alive
}
void launch_gather_bench(const uint64_t *buf, uint64_t n_words, uint64_t n_threads, uint32_t per_thread, unsigned long long *sink,
                         int variant, cudaStream_t s) {
    const uint32_t blocks = (uint32_t)((n_threads + 255) / 256);
    switch (variant) {
        case 1: gather_bench_kernel<1><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        case 2: gather_bench_kernel<2><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        case 3: gather_bench_kernel<3><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        case 4: gather_bench_kernel<4><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        case 5: gather_bench_kernel<5><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        case 6: gather_bench_kernel<6><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        case 7: gather_bench_kernel<7><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        case 8: gather_bench_kernel<8><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
        default: gather_bench_kernel<0><<<blocks, 256, 0, s>>>(buf, n_words, per_thread, sink); break;
    }
}

// ---------------------------------------------------------------- test hooks

__global__ void debug_lf_kernel(DevIndex ix, const uint64_t *rows, uint64_t n, uint64_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t row = rows[i];
    out[i] = row < ix.n ? (uint64_t)lf_from_rec(ix, ld_rec(ix, (uint32_t)row)) : ~0ull;
}
__global__ void debug_rank_kernel(DevIndex ix, const uint32_t *codes, const uint64_t *rows, uint64_t n, uint64_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = codes[i];
    const uint64_t row = rows[i];
    if (c >= ix.A || row >= ix.n) { out[i] = ~0ull; return; }
    const Sector s = load_sector(ix, c, (uint32_t)row >> ix.blk_shift);
    out[i] = sector_rank(ix, s, (uint32_t)row & ((1u << ix.blk_shift) - 1));
}
__global__ void debug_bwt_kernel(DevIndex ix, uint64_t row0, uint64_t n, uint32_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = row0 + i < ix.n ? rec_code(ix.rec[row0 + i]) : 0xFFFFFFFFu;
}
__global__ void debug_marks_kernel(DevIndex ix, uint64_t row0, uint64_t n, uint64_t *out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = ~0ull;
    if (row0 + i < ix.n) {
        const uint64_t r = ix.rec[row0 + i];
        if (rec_marked(r)) v = rec_low(r);
    }
    out[i] = v;
}
void launch_debug_lf(const DevIndex &ix, const uint64_t *rows, uint64_t n, uint64_t *out, cudaStream_t s) {
    if (n) debug_lf_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(ix, rows, n, out);
}
void launch_debug_rank(const DevIndex &ix, const uint32_t *codes, const uint64_t *rows, uint64_t n, uint64_t *out, cudaStream_t s) {
    if (n) debug_rank_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(ix, codes, rows, n, out);
}
void launch_debug_bwt(const DevIndex &ix, uint64_t row0, uint64_t n, uint32_t *out, cudaStream_t s) {
    if (n) debug_bwt_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(ix, row0, n, out);
}
void launch_debug_marks(const DevIndex &ix, uint64_t row0, uint64_t n, uint64_t *out, cudaStream_t s) {
    if (n) debug_marks_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(ix, row0, n, out);
}

}  // namespace e2fm
