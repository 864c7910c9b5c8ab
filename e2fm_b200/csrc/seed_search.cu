// seed_search.cu — the search of fixed-length batches over 2-bit packed patterns: one THREAD per pattern, all k shifts.
//
//   pack_patterns_kernel   ASCII -> 2 bits per base (A C G T = 0 1 2 3, base i at bits 2i of the pattern's little-endian bit
//                          string, 16 bytes per 64 bases) + a per-pattern flag: plain / holds an index symbol outside ACGT /
//                          holds a symbol that is not in the index's alphabet (EFMIndex.cpp:1702-1717: no match at all)
//   seed_search_kernel     EFMIndex::computeSuperPatterns (:2055-2207) from the packed registers (funnel shift + a shared-memory
//                          table packed k-mer -> compact code); the interval after the LAST j super-characters of every shift
//                          from the seed table (seed.cuh: start interval :1728-1748 + j-1 backwardSteps :1966 in one gather);
//                          backwardSteps over rank sectors while more than one row is left; and, once ONE row is left, the rest
//                          of EFMIndex::exactMatch(interval) (:1901-1963), backwardStepIfMatch (:2011-2041),
//                          findWholePatternMatches (:1862-1899) and getRowPosition (:2468-2497) answered from the dense tables:
//                          a backward step from a single row r with character c succeeds iff text[sa[r] - 1] == c and leads to
//                          the row of text position sa[r] - 1, so the remaining super-characters, the '?'-prefixed first one
//                          and the '?'-suffixed last one are compared with ONE contiguous run of text[] around sa[r].
//                          A pattern that ends with more than one row in some shift, or meets a saturated seed counter, is
//                          handed to the generic path (search.cu) untouched — its answers come from there.
//   materialize_kernel     the patterns of that list as an ASCII sub-batch for the generic path
//
// Per pattern of the 3.1 Gbp collection (k = 4, m = 50): 4 seed sectors + sa[] + ~2 adjacent text[] sectors for the shift that
// matches, against ~24 rank sectors + 2-3 gathers in resolve_kernel for the chain of backward steps.
#include "search.cuh"

#include <algorithm>

#include "../../include/e2fm_b200.h"
#include "seed.cuh"

namespace e2fm {

static constexpr unsigned FULL = 0xFFFFFFFFu;

// ---------------------------------------------------------------- ASCII -> packed

__global__ void __launch_bounds__(256) pack_patterns_kernel(DevIndex ix, const uint8_t *__restrict__ bytes, uint64_t first, uint32_t count,
                                                             uint32_t m, uint32_t nw, unsigned long long *__restrict__ packed,
                                                             uint8_t *__restrict__ pflag) {
    extern __shared__ __align__(16) uint8_t pk_smem[];
    __shared__ uint8_t lut[256];
    // 0..3 = A C G T, 4 = another symbol of the index's alphabet, 5 = not a symbol of this index
    {
        const uint32_t b = threadIdx.x;
        const uint8_t si = ix.sym_index[b];
        uint8_t v = si == 0xFFu ? 5 : 4;
        if (si != 0xFFu) {
            if (b == 'A') v = 0; else if (b == 'C') v = 1; else if (b == 'G') v = 2; else if (b == 'T') v = 3;
        }
        lut[b] = v;
    }
    const uint64_t p0 = first + (uint64_t)blockIdx.x * 256;
    const uint32_t here = min(256u, count - blockIdx.x * 256u);
    // stage this CTA's bytes with aligned 4-byte loads
    const uint8_t *src = bytes + p0 * m;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t total = here * m + mis;
    const uint32_t *src4 = reinterpret_cast<const uint32_t *>(src - mis);
    uint32_t *dst4 = reinterpret_cast<uint32_t *>(pk_smem);
    for (uint32_t i = threadIdx.x; i < (total + 3) / 4; i += blockDim.x) dst4[i] = __ldg(src4 + i);
    __syncthreads();
    if (threadIdx.x >= here) return;
    const uint8_t *mine = pk_smem + mis + threadIdx.x * m;
    const uint64_t p = p0 + threadIdx.x;
    uint32_t worst = 0, j = 0;
    for (uint32_t w = 0; w < 2 * nw; w++) {
        unsigned long long cur = 0;
        const uint32_t end = min(m, (w + 1) * 32u);
        for (uint32_t sh = 0; j < end; j++, sh += 2) {
            const uint32_t c = lut[mine[j]];
            worst = max(worst, c);
            cur |= (unsigned long long)(c & 3u) << sh;
        }
        packed[p * 2 * nw + w] = cur;
    }
    pflag[p] = worst <= 3 ? 0 : (worst == 4 ? 1 : 2);
}

void launch_pack_patterns(const DevIndex &ix, const uint8_t *bytes, uint64_t first, uint32_t count, uint32_t m, uint32_t nw,
                          uint4 *packed, uint8_t *pflag, cudaStream_t s) {
    if (count == 0) return;
    const size_t smem = (size_t)256 * m + 8;
    pack_patterns_kernel<<<(count + 255) / 256, 256, smem, s>>>(ix, bytes, first, count, m, nw,
                                                               reinterpret_cast<unsigned long long *>(packed), pflag);
}

// ---------------------------------------------------------------- tight 2-bit packing -> 16-byte words

// Patterns that travelled over PCIe as ceil(m / 4) bytes each (the first bytes of the 16-byte form: base i at bits 2i) are widened
// to the 16 bytes per 64 bases the search kernels load with one instruction. A thread per pattern; consecutive threads read
// consecutive byte runs, so the warp's loads cover a contiguous span.
__global__ void __launch_bounds__(256) widen_packed_kernel(const uint8_t *__restrict__ tight, uint64_t first, uint32_t count, uint32_t m,
                                                            uint32_t nw, unsigned long long *__restrict__ packed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t p = first + i;
    const uint32_t tb = (m + 3) / 4;
    const uint8_t *src = tight + p * tb;
    for (uint32_t w = 0; w < 2 * nw; w++) {
        unsigned long long cur = 0;
        const uint32_t b0 = w * 8, b1 = min(tb, b0 + 8);
        for (uint32_t b = b0; b < b1; b++) cur |= (unsigned long long)__ldg(src + b) << (8 * (b - b0));
        const uint32_t lo = w * 32;                       // bases [lo, lo + 32) live in this word; those from m on are zero
        if (m < lo + 32) cur &= m > lo ? (~0ull >> (64 - 2 * (m - lo))) : 0ull;
        packed[p * 2 * nw + w] = cur;
    }
}

void launch_widen_packed(const uint8_t *tight, uint64_t first, uint32_t count, uint32_t m, uint32_t nw, uint4 *packed, cudaStream_t s) {
    if (count == 0) return;
    widen_packed_kernel<<<(count + 255) / 256, 256, 0, s>>>(tight, first, count, m, nw, reinterpret_cast<unsigned long long *>(packed));
}

// ---------------------------------------------------------------- the search: seed lookup, then verification

// which super-characters of shift sa are searched: l = their end (exclusive), lo = the first fixed one, or l <= 0 when the shift
// yields nothing (EFMIndex.cpp:2132 / :1745)
__host__ __device__ __forceinline__ void shift_geometry(uint32_t m, uint32_t k, uint32_t sa, uint32_t &L, bool &last_var, int &l, int &lo) {
    L = (m + sa + k - 1) / k;
    last_var = (m + sa) % k != 0;
    l = -1;
    if (!last_var) { if (!(L == 1 && sa > 0)) l = (int)L; }
    else if (L > 1 && (L > 2 || sa == 0)) l = (int)L - 1;
    lo = sa > 0 ? 1 : 0;
}

// A pattern's bit string in registers: NW x 128 bits. bits(off, mask): up to 32 bits at bit offset off.
template <int NW>
struct PatRegs {
    uint32_t w[4 * NW];
    __device__ __forceinline__ void load(const uint4 *src) {
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const uint4 v = __ldg(src + i);
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
    // word i of the bit string (0 beyond its end), as a tree of selects: a dynamically indexed register array would live in local memory
    __device__ __forceinline__ uint32_t word(uint32_t i) const {
        uint32_t r = 0;
#pragma unroll
        for (int g = 0; g < NW; g++) {
            const uint32_t a = (i & 1u) ? w[4 * g + 1] : w[4 * g], b = (i & 1u) ? w[4 * g + 3] : w[4 * g + 2];
            const uint32_t v = (i & 2u) ? b : a;
            r = (i >> 2) == (uint32_t)g ? v : r;
        }
        return r;
    }
    // WN words of the bit string starting at bit offset `off`, which may be negative (bits before the string read as zero): the
    // caller then picks its fields at compile-time offsets of `out` — one round of selects instead of one per field
    template <int WN>
    __device__ __forceinline__ void window(int off, uint32_t (&out)[WN]) const {
        const int wi = off >> 5;                       // floor
        const uint32_t sh = (uint32_t)off & 31u;
        uint32_t prev = word((uint32_t)wi);
#pragma unroll
        for (int i = 0; i < WN; i++) {
            const uint32_t next = word((uint32_t)(wi + i + 1));
            out[i] = __funnelshift_r(prev, next, sh);
            prev = next;
        }
    }
    __device__ __forceinline__ uint32_t bits(uint32_t off, uint32_t mask) const {
        const uint32_t wi = off >> 5;
        return __funnelshift_r(word(wi), word(wi + 1), off & 31u) & mask;
    }
    // natural value (base n_sym, first base most significant) of the f bases at bit offset off: what kmer_low / kmer_high hold
    __device__ __forceinline__ uint32_t natural(const DevIndex &ix, uint32_t off, uint32_t f) const {
        const uint32_t b = bits(off, f >= 16 ? 0xFFFFFFFFu : ((1u << (2 * f)) - 1u));
        uint32_t v = 0;
        bool bad = false;
        for (uint32_t i = 0; i < f; i++) {
            const uint32_t s = ix.sym4[(b >> (2 * i)) & 3u];
            bad |= s == 0xFFu;
            v = v * ix.n_sym + s;
        }
        return bad ? 0xFFFFFFFFu : v;
    }
};

static constexpr int SEED_U = 2;                                 // tiles per iteration: their loads are in flight together
static constexpr uint32_t SL_DIRECT = 4;                         // seed intervals longer than this are written straight to the task list
static constexpr uint32_t SL_FLUSH = 256;                        // staged row tasks a warp hands over per reservation
static constexpr uint32_t SL_WCAP = SL_FLUSH + 32 * SL_DIRECT;   // per-warp stage: a tile adds at most SL_DIRECT rows per lane

// the interval of the seed c_0 .. c_{j-1} by the backward steps the table replaces (EFMIndex.cpp:1728-1748, :1966-2001): the path
// of a lookup that met a saturated counter (a seed that occurs 15 times or more, or sits behind one in its sector)
template <int KT>
__device__ __forceinline__ void seed_interval_slow(const DevIndex &ix, const uint16_t *ctab, unsigned long long window, uint32_t j,
                                                   uint32_t &s, uint32_t &e) {
    constexpr uint32_t KMASK = (1u << (2 * KT)) - 1u;
    const uint32_t blk_mask = (1u << ix.blk_shift) - 1;
    uint32_t cc = ctab[(uint32_t)(window >> (2 * KT * (j - 1))) & KMASK];
    s = (uint32_t)__ldg(ix.C + cc);
    e = (uint32_t)__ldg(ix.C + cc + 1);
    for (int q = (int)j - 2; q >= 0 && e > s; q--) {
        cc = ctab[(uint32_t)(window >> (2 * KT * q)) & KMASK];
        const uint32_t re = e - 1;
        const uint32_t ns = s == 0 ? (uint32_t)__ldg(ix.C + cc) : rank_c(ix, cc, s - 1);
        e = sector_rank(ix, load_sector(ix, cc, re >> ix.blk_shift), re & blk_mask);
        s = ns;
    }
}

// K_a: one LANE per (pattern, shift). EFMIndex::computeSuperPatterns (:2055-2207) for the last j fixed super-characters of the shift
// from the packed pattern, ONE sector gather of the seed table (start interval :1728-1748 + j-1 backwardSteps :1966), and the rows
// of the interval become row tasks {row, item}. The k lanes of a pattern sit next to each other in a warp.
// The kernel is bound by instruction issue as much as by the gathers (ncu, profiles/): a lane loads only the three words of the
// pattern that hold its seed (where they are depends on the shift alone), the row tasks are staged per WARP (no CTA barrier in the
// loop, one reservation per SL_FLUSH tasks), and the loads of SEED_U tiles are in flight together.
template <int KT, int NW, int JT>                      // JT = compile-time seed length (0: read it from the index)
__global__ void __launch_bounds__(256, 5) seed_lookup_kernel(SeedArgs a) {
    extern __shared__ uint16_t ctab[];                 // [4^k] packed k-mer -> compact code
    __shared__ uint2 st_task[8][SL_WCAP];
    const DevIndex &ix = a.ix;
    constexpr uint32_t k = KT, PPW = 32 / KT, PPC = 8 * PPW;
    constexpr uint32_t KMASK = (1u << (2 * KT)) - 1u;
    constexpr int U = SEED_U;
    for (uint32_t i = threadIdx.x; i < (1u << (2 * KT)); i += blockDim.x) ctab[i] = ix.packed_code[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const uint32_t pl = lane / k, sa = lane - pl * k;
    const uint32_t m = a.m, j = JT ? (uint32_t)JT : ix.seed_chars, A = ix.A;
    uint32_t L;
    bool last_var;
    int l, lo;
    shift_geometry(m, k, sa, L, last_var, l, lo);
    const uint32_t first_bit = l > 0 ? 2u * (((uint32_t)l - j) * k - sa) : 0u;   // bit offset of the seed's first super-character
    const uint32_t w0 = first_bit >> 5, sh = first_bit & 31u;                     // the seed sits in words w0 .. w0 + 2 of the pattern
    const bool live = pl < PPW && l > 0;
    const unsigned group = (pl < PPW ? ((1u << k) - 1u) << (pl * k) : 0u);
    const uint32_t *words = reinterpret_cast<const uint32_t *>(a.packed);
    uint2 *stage = st_task[warp];
    uint32_t wcnt = 0;                                 // staged tasks of this warp (uniform)
    unsigned long long n_sect = 0;
    auto flush = [&]() {                               // whole warp
        unsigned long long b0 = 0;
        if (lane == 0) b0 = atomicAdd(a.task_cursor, (unsigned long long)wcnt);
        b0 = __shfl_sync(FULL, b0, 0);
        __syncwarp();
        for (uint32_t i = lane; i < wcnt; i += 32)
            if (b0 + i < a.task_cap) a.tasks[b0 + i] = stage[i];
        __syncwarp();
        wcnt = 0;
    };
    const uint32_t tiles = (a.count + PPC - 1) / PPC;
    // the pattern words of the NEXT group of tiles are loaded while the seed gathers of the current one are in flight: a thread then
    // has its gathers outstanding nearly all the time instead of half of it
    uint32_t nx0[U], nx1[U], nx2[U], nflag[U];
    auto fetch_words = [&](uint32_t g) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t pq = (g + u) * PPC + warp * PPW + pl;
            nx0[u] = nx1[u] = nx2[u] = 0;
            nflag[u] = 0;
            if (pl < PPW && g + u < tiles && pq < a.count) {
                const uint64_t p = a.first + pq;
                if (live) {
                    const uint32_t *src = words + p * (4 * NW) + w0;
                    nx0[u] = __ldg(src);
                    if (w0 + 1 < 4 * NW) nx1[u] = __ldg(src + 1);
                    if (w0 + 2 < 4 * NW) nx2[u] = __ldg(src + 2);
                }
                if (a.pflag) nflag[u] = a.pflag[p];
                if (sa == 0) a.counts[p] = 0;          // seed_verify_kernel (and the generic path for list members) add to it
            }
        }
    };
    if (blockIdx.x * U < tiles) fetch_words(blockIdx.x * U);
    for (uint32_t g = blockIdx.x * U; g < tiles; g += gridDim.x * U) {
        uint32_t p_local[U], flag[U], x0[U], x1[U], x2[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; u++) {                  // the seed windows of U tiles
            p_local[u] = (g + u) * PPC + warp * PPW + pl;
            valid[u] = pl < PPW && g + u < tiles && p_local[u] < a.count;
            flag[u] = nflag[u];
            x0[u] = nx0[u]; x1[u] = nx1[u]; x2[u] = nx2[u];
        }
        uint4 sec[U];
        unsigned long long window[U];
        uint32_t slot[U];
        bool look[U];
#pragma unroll
        for (int u = 0; u < U; u++) {                  // their seed gathers
            window[u] = (unsigned long long)__funnelshift_r(x0[u], x1[u], sh) | ((unsigned long long)__funnelshift_r(x1[u], x2[u], sh) << 32);
            slot[u] = 0;
            look[u] = false;
            if (valid[u] && live && flag[u] == 0) {
                uint32_t hi = 0, cl = 0;
                bool absent = false;
#pragma unroll
                for (uint32_t q = 0; q < j; q++) {
                    const uint32_t code = ctab[(uint32_t)(window[u] >> (2 * KT * q)) & KMASK];
                    absent |= code == 0xFFFFu;
                    if (q + 1 < j) hi = hi * A + code; else cl = code;
                }
                if (!absent) {
                    size_t entry = seed_sector(ix, hi, cl, slot[u]);
                    if (a.debug & 1u) entry &= 63;
                    sec[u] = seed_load(ix, entry);
                    look[u] = true;
                    n_sect++;
                }
            }
            flag[u] = (valid[u] && flag[u] == 1) ? 1u : 0u;   // from here on: 1 = the pattern goes to the generic path
        }
        if (g + gridDim.x * U < tiles) fetch_words(g + gridDim.x * U);
#pragma unroll
        for (int u = 0; u < U; u++) {                  // rows of the intervals -> staged row tasks
            uint32_t cnt = 0, s = 0, big = 0;
            if (look[u]) {
                const SeedHit h = seed_decode(sec[u], slot[u]);
                if (h.exact) { s = h.s; cnt = h.cnt; }
                else {
                    uint32_t e;
                    seed_interval_slow<KT>(ix, ctab, window[u], j, s, e);
                    cnt = e > s ? e - s : 0u;
                }
                if (cnt > SL_DIRECT) { big = cnt; cnt = 0; }
            }
            // a pattern with an index symbol outside ACGT is searched by the generic path, all its shifts
            const unsigned irr = __ballot_sync(FULL, flag[u] != 0);
            if (irr & group) {
                cnt = 0;
                big = 0;
                if (flag[u] && sa == 0) {
                    const unsigned long long at = atomicAdd(a.stats + ST_COMPLEX, 1ull);
                    a.complex_list[at] = (uint32_t)(a.first + p_local[u]);
                }
            }
            const uint32_t item = p_local[u] * k + sa;
            // longer intervals (repeats) are written by the whole warp, straight to the task list
            unsigned bigm = __ballot_sync(FULL, big != 0);
            while (bigm) {
                const int src = __ffs(bigm) - 1;
                bigm &= bigm - 1;
                const uint32_t r0 = __shfl_sync(FULL, s, src), ln = __shfl_sync(FULL, big, src), im = __shfl_sync(FULL, item, src);
                unsigned long long b0 = 0;
                if (lane == 0) b0 = atomicAdd(a.task_cursor, (unsigned long long)ln);
                b0 = __shfl_sync(FULL, b0, 0);
                for (uint32_t q = lane; q < ln; q += 32)
                    if (b0 + q < a.task_cap) a.tasks[b0 + q] = make_uint2(r0 + q, im);
            }
            const unsigned any = __ballot_sync(FULL, cnt != 0);
            if (any) {
                const unsigned more = __ballot_sync(FULL, cnt > 1);
                uint32_t excl, wtot;
                if (!more) {                           // the usual case: no lane has more than one row
                    excl = __popc(any & lt);
                    wtot = __popc(any);
                } else {
                    uint32_t incl = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t v = __shfl_up_sync(FULL, incl, d);
                        if ((int)lane >= d) incl += v;
                    }
                    excl = incl - cnt;
                    wtot = __shfl_sync(FULL, incl, 31);
                }
                for (uint32_t q = 0; q < cnt; q++) stage[wcnt + excl + q] = make_uint2(s + q, item);
                wcnt += wtot;
                if (wcnt >= SL_FLUSH) flush();
            }
        }
    }
    if (wcnt) flush();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_sect += __shfl_xor_sync(FULL, n_sect, d);
    if (lane == 0 && n_sect) atomicAdd(a.stats + ST_SECT_BS, n_sect);
}

// K_a for short patterns: one LANE per (pattern, lookup). A shift with one fixed super-character fewer than a seed is looked up
// once for every k-mer that completes its '?'-suffixed last super-character (EFMIndex::findWholePatternMatches, :1862-1899, checks
// that character row by row AFTER the search; rows of a suffix-array interval are sorted by what follows, so the rows that pass
// are the union of the intervals of the completed seeds). The lookups of a pattern are laid out behind one another (a small table
// in shared memory says which shift and which completion lookup t is), so that lanes stay busy whatever the shift: for m = 12,
// k = 4 a pattern has 1 + 64 + 16 + 4 of them, all in a seed table of a few MB that lives in the L2.
// The rows of an interval are pre-filtered with the '?'-prefixed first super-character where there is one
// (EFMIndex::backwardStepIfMatch, :2011-2041: the BWT symbol of a row is the character before its suffix, so bwt16[row] is tested
// against the pattern's first bases); seed_verify_rec_kernel repeats that test from the record, so the filter only saves tasks.
static constexpr uint32_t SE_DIRECT = 14;                        // every exact counter is staged
static constexpr uint32_t SE_FLUSH = 128, SE_WCAP = SE_FLUSH + 32 * SE_DIRECT;
static constexpr uint32_t SE_MAX_LOOKUPS = 8192;

template <int KT, int NW>
__global__ void __launch_bounds__(256) seed_lookup_enum_kernel(SeedArgs a) {
    extern __shared__ uint16_t ctab[];                 // [4^k] packed k-mer -> compact code, then tmap[lookups]
    __shared__ uint2 st_task[8][SE_WCAP];
    const DevIndex &ix = a.ix;
    constexpr uint32_t k = KT;
    constexpr uint32_t KMASK = (1u << (2 * KT)) - 1u;
    uint32_t *tmap = reinterpret_cast<uint32_t *>(ctab + (1u << (2 * KT)));   // lookup t of a pattern: shift << 16 | completion
    const uint32_t m = a.m, j = ix.seed_chars, A = ix.A, T = a.lookups;
    for (uint32_t i = threadIdx.x; i < (1u << (2 * KT)); i += blockDim.x) ctab[i] = ix.packed_code[i];
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (uint32_t sa = 0; sa < k; sa++) {
            uint32_t L;
            bool last_var;
            int l, lo;
            shift_geometry(m, k, sa, L, last_var, l, lo);
            if (l <= 0) continue;
            const uint32_t nfix = (uint32_t)(l - lo);
            const uint32_t cand = nfix >= j ? 1u : 1u << (2 * (k - (m + sa) % k));
            for (uint32_t c = 0; c < cand && t < T; c++) tmap[t++] = (sa << 16) | c;
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    uint2 *stage = st_task[warp];
    uint32_t wcnt = 0;
    unsigned long long n_sect = 0;
    auto flush = [&]() {                               // whole warp
        unsigned long long b0 = 0;
        if (lane == 0) b0 = atomicAdd(a.task_cursor, (unsigned long long)wcnt);
        b0 = __shfl_sync(FULL, b0, 0);
        __syncwarp();
        for (uint32_t i = lane; i < wcnt; i += 32)
            if (b0 + i < a.task_cap) a.tasks[b0 + i] = stage[i];
        __syncwarp();
        wcnt = 0;
    };
    const unsigned long long n_items = (unsigned long long)a.count * T;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    // whole warps run the loop (the ballots below need every lane)
    for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x + warp * 32u; base < n_items; base += stride) {
        const unsigned long long it = base + lane;
        uint32_t cnt = 0, s = 0, big = 0, item = 0;
        if (it < n_items) {
            const uint32_t p_local = (uint32_t)(it / T), t = (uint32_t)(it - (unsigned long long)p_local * T);
            const uint64_t p = a.first + p_local;
            const uint32_t flag = a.pflag ? a.pflag[p] : 0;
            if (t == 0) {
                a.counts[p] = 0;                       // seed_verify_rec_kernel (and the generic path for list members) add to it
                if (flag == 1) {                       // an index symbol outside ACGT: the generic path searches this pattern
                    const unsigned long long at = atomicAdd(a.stats + ST_COMPLEX, 1ull);
                    a.complex_list[at] = (uint32_t)p;
                }
            }
            if (flag == 0) {
                const uint32_t tm_ = tmap[t], sa = tm_ >> 16, c = tm_ & 0xFFFFu;
                uint32_t L;
                bool last_var;
                int l, lo;
                shift_geometry(m, k, sa, L, last_var, l, lo);
                const bool en = (uint32_t)(l - lo) < j;                              // the hat's completion c is the seed's last character
                const uint32_t nf = en ? j - 1 : j;                                 // fixed super-characters of the seed: l - nf .. l - 1
                PatRegs<NW> pat;
                pat.load(a.packed + p * NW);
                unsigned long long window = 0;
                for (uint32_t q = 0; q < nf; q++)
                    window |= (unsigned long long)pat.bits(2u * (((uint32_t)l - nf + q) * k - sa), KMASK) << (2 * KT * q);
                if (en) {
                    const uint32_t fh = (m + sa) % k;                               // known (first) bases of the last super-character
                    const uint32_t known = pat.bits(2u * ((uint32_t)l * k - sa), (1u << (2 * fh)) - 1u);
                    window |= (unsigned long long)(known | (c << (2 * fh))) << (2 * KT * nf);
                }
                uint32_t hi = 0, cl = 0;
                bool absent = false;
                for (uint32_t q = 0; q < j; q++) {
                    const uint32_t code = ctab[(uint32_t)(window >> (2 * KT * q)) & KMASK];
                    absent |= code == 0xFFFFu;
                    if (q + 1 < j) hi = hi * A + code; else cl = code;
                }
                item = p_local * k + sa;
                if (!absent) {
                    uint32_t slot;
                    const size_t entry = seed_sector(ix, hi, cl, slot);
                    const SeedHit h = seed_decode(seed_load(ix, entry), slot);
                    n_sect++;
                    if (h.exact) { s = h.s; cnt = h.cnt; }
                    else {
                        uint32_t e;
                        seed_interval_slow<KT>(ix, ctab, window, j, s, e);
                        cnt = e > s ? e - s : 0u;
                    }
                    if (cnt > SE_DIRECT) { big = cnt; cnt = 0; }
                    else if (cnt && sa > 0 && ix.bwt16) {
                        // '?'-prefixed first super-character: keep the rows whose BWT symbol ends with the pattern's first bases.
                        // The survivors of a short interval are re-packed as (first surviving row, bitmap of the others).
                        const uint32_t f = k - sa;
                        const uint32_t want = pat.natural(ix, 0, f);
                        uint32_t keep = 0;
                        if (want != 0xFFFFFFFFu)
                            for (uint32_t q = 0; q < cnt; q++)
                                if (__ldg(ix.kmer_low + (f - 1) * A + (uint32_t)__ldg(ix.bwt16 + s + q)) == want) keep |= 1u << q;
                        n_sect++;
                        big = 0;
                        cnt = __popc(keep);
                        cnt |= keep << 8;                                           // the bitmap rides in the high bits to the staging below
                    }
                }
            }
        }
        // long intervals (repeats) are written by the whole warp, straight to the task list (unfiltered: the verification decides)
        unsigned bigm = __ballot_sync(FULL, big != 0);
        while (bigm) {
            const int src = __ffs(bigm) - 1;
            bigm &= bigm - 1;
            const uint32_t r0 = __shfl_sync(FULL, s, src), ln = __shfl_sync(FULL, big, src), im = __shfl_sync(FULL, item, src);
            unsigned long long b0 = 0;
            if (lane == 0) b0 = atomicAdd(a.task_cursor, (unsigned long long)ln);
            b0 = __shfl_sync(FULL, b0, 0);
            for (uint32_t q = lane; q < ln; q += 32)
                if (b0 + q < a.task_cap) a.tasks[b0 + q] = make_uint2(r0 + q, im);
        }
        const uint32_t keep = cnt >> 8;                // 0: all rows s .. s + cnt - 1
        cnt &= 0xFFu;
        const unsigned any = __ballot_sync(FULL, cnt != 0);
        if (any) {
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, incl, d);
                if ((int)lane >= d) incl += v;
            }
            const uint32_t excl = incl - cnt, wtot = __shfl_sync(FULL, incl, 31);
            if (keep) {
                uint32_t q = 0;
                for (uint32_t bits = keep; bits; bits &= bits - 1) stage[wcnt + excl + q++] = make_uint2(s + (uint32_t)__ffs(bits) - 1u, item);
            } else {
                for (uint32_t q = 0; q < cnt; q++) stage[wcnt + excl + q] = make_uint2(s + q, item);
            }
            wcnt += wtot;
            if (wcnt >= SE_FLUSH) flush();
        }
        (void)lt;
    }
    if (wcnt) flush();
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_sect += __shfl_xor_sync(FULL, n_sect, d);
    if (lane == 0 && n_sect) atomicAdd(a.stats + ST_SECT_BS, n_sect);
}

// text[] is read as whole aligned sectors (one 32-byte load each): with narrower loads a sector is requested from L2 once per load,
// because thousands of gathers in flight per SM thrash the L1
struct Sec8 { uint32_t w[8]; };
__device__ __forceinline__ Sec8 ld_sector256(const void *p) {
    const Sector s = ld_sector(reinterpret_cast<const uint4 *>(p));
    Sec8 r;
    r.w[0] = s.lo.x; r.w[1] = s.lo.y; r.w[2] = s.lo.z; r.w[3] = s.lo.w;
    r.w[4] = s.hi.x; r.w[5] = s.hi.y; r.w[6] = s.hi.z; r.w[7] = s.hi.w;
    return r;
}

// K_b: one thread per row task {row, item}, SEED_UV tasks at a time. The row's suffix starts at text position sa[row] with the j seed
// super-characters; what is left of EFMIndex::exactMatch(interval) (:1901-1963) for this row — the fixed super-characters before the
// seed, the '?'-prefixed first one (backwardStepIfMatch, :2011-2041) and the '?'-suffixed last one (findWholePatternMatches,
// :1862-1899) — is compared with the run of text[] around that position, and the occurrence reported at the position getRowPosition
// (:2468-2497) would walk to. A backward step from a single row r with character c succeeds iff text[sa[r] - 1] == c, and then leads
// to the row of sa[r] - 1.
// The run is read as the two aligned sectors that end with text[sa[r] - 1] (one 32-byte load each, issued for all the thread's tasks
// before any is used) and laid into a per-thread column of shared memory, from where the characters are picked; the few characters
// outside that window (long patterns, the hat character behind the seed) are single loads.
static constexpr int SEED_UV = 2;
static constexpr uint32_t SV_FLUSH = 1024, SV_CAP = SV_FLUSH + 256 * SEED_UV;

template <int KT, int NW>
__global__ void __launch_bounds__(256) seed_verify_kernel(SeedArgs a) {
    extern __shared__ uint16_t ctab[];
    __shared__ uint32_t rs_pat[SV_CAP];
    __shared__ uint64_t rs_pos[SV_CAP];
    __shared__ uint32_t win[16 * 256];                 // [word][thread]: the 32 text codes that end with text[sa[row] - 1]'s sector
    __shared__ uint32_t rs_cnt;
    __shared__ unsigned long long rs_base;
    const DevIndex &ix = a.ix;
    constexpr uint32_t k = KT;
    constexpr uint32_t KMASK = (1u << (2 * KT)) - 1u;
    constexpr int U = SEED_UV;
    for (uint32_t i = threadIdx.x; i < (1u << (2 * KT)); i += blockDim.x) ctab[i] = ix.packed_code[i];
    if (threadIdx.x == 0) rs_cnt = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    const uint32_t m = a.m, j = ix.seed_chars, A = ix.A, n = (uint32_t)ix.n;
    const bool locate = a.mode == E2FM_MODE_LOCATE;
    const unsigned long long nt = *a.task_cursor;
    const uint32_t n_tasks = (uint32_t)(nt > a.task_cap ? 0ull : nt);     // an overflowing chunk is searched again by the host
    unsigned long long n_sect = 0;
    auto flush = [&]() {
        const uint32_t c = rs_cnt;
        if (threadIdx.x == 0) rs_base = atomicAdd(a.stats + ST_RESULTS, (unsigned long long)c);
        __syncthreads();
        const unsigned long long b0 = rs_base;
        uint32_t lost = 0;
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
            if (b0 + i < a.res_cap) { a.res_pat[b0 + i] = rs_pat[i]; a.res_pos[b0 + i] = rs_pos[i]; }
            else lost++;
        }
        if (lost) atomicAdd(a.stats + ST_RES_OVF, (unsigned long long)lost);
        __syncthreads();
        if (threadIdx.x == 0) rs_cnt = 0;
        __syncthreads();
    };
    const uint16_t *text = ix.text;
    uint32_t *col = win + threadIdx.x;                  // word w of this thread's window: col[w * 256]
    const uint32_t stride = gridDim.x * blockDim.x * U;
    for (uint32_t base = blockIdx.x * blockDim.x * U; base < n_tasks; base += stride) {
        uint2 tk[U];
        bool ok[U];
        uint32_t pos[U];
        PatRegs<NW> pat[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t idx = base + (uint32_t)u * blockDim.x + threadIdx.x;
            ok[u] = idx < n_tasks;
            tk[u] = ok[u] ? __ldg(a.tasks + idx) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {                  // sa[] gathers and pattern loads of the thread's tasks in flight together
            pos[u] = 0;
            if (ok[u]) {
                pos[u] = ld_gather_u32(ix.sa + tk[u].x);
                pat[u].load(a.packed + (a.first + tk[u].y / k) * NW);
            }
        }
        uint32_t sa[U], d[U], Ls[U], wbase[U], hat[U];
        bool lastv[U], in_win[U], two[U];
        Sec8 secA[U], secB[U];
#pragma unroll
        for (int u = 0; u < U; u++) {                  // the sectors of text[] of all tasks
            sa[u] = tk[u].y % k;
            int l, lo;
            shift_geometry(m, k, sa[u], Ls[u], lastv[u], l, lo);
            d[u] = (uint32_t)(l - (int)j);              // super-characters 0 .. l-j-1 lie at text positions pos - d .. pos - 1
            in_win[u] = ok[u] && pos[u] >= d[u] && pos[u] >= 32u;   // (else: the run wraps around the text, or sits at its very start)
            const uint32_t sb = in_win[u] ? ((pos[u] - 1) >> 4) : 1u;
            wbase[u] = (sb - 1) << 4;
            two[u] = in_win[u] && pos[u] - d[u] < (sb << 4);
            hat[u] = 0;
            if (in_win[u]) {
                secB[u] = ld_sector256(text + ((size_t)sb << 4));
                if (two[u]) secA[u] = ld_sector256(text + ((size_t)wbase[u]));
                const unsigned long long ph = (unsigned long long)pos[u] - d[u] + Ls[u] - 1;   // the '?'-suffixed last super-character
                if (lastv[u] && ph < ix.n && (ph >> 4) != sb) hat[u] = __ldg(text + ph);
            }
        }
        uint32_t tp[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            tp[u] = pos[u] >= d[u] ? pos[u] - d[u] : pos[u] + (n - d[u]);               // text position of super-character 0
            if (!ok[u]) continue;
            const uint32_t lo = sa[u] > 0 ? 1u : 0u;
            const uint32_t bit0 = 2u * (d[u] * k - sa[u]);                              // bit offset of super-character d (the seed's first)
            if (in_win[u]) {
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    if (two[u]) col[w * 256] = secA[u].w[w];
                    col[(8 + w) * 256] = secB[u].w[w];
                }
            }
            // text code at position q: from the window when it is there
            auto code_at = [&](uint32_t q) -> uint32_t {
                const uint32_t h = q - wbase[u];
                if (in_win[u] && h < 32u && (h >= 16u || two[u])) return (col[(h >> 1) * 256] >> ((h & 1u) * 16u)) & 0xFFFFu;
                return (uint32_t)__ldg(text + q);
            };
            bool good = true;
            // text[pos - r] against super-character d - r, r = 1 .. d - lo (the fixed ones, right to left)
            for (uint32_t r = 1; r + lo <= d[u] && good; r++) {
                const uint32_t code = ctab[pat[u].bits(bit0 - 2u * r * k, KMASK)];
                const uint32_t at = pos[u] >= r ? pos[u] - r : pos[u] + (n - r);        // (pos - r) mod n
                good = code == code_at(at);
            }
            if (good && sa[u] > 0) {                                                    // '?'-prefixed first super-character
                const uint32_t f = k - sa[u];
                const uint32_t want = pat[u].natural(ix, 0, f);
                good = want != 0xFFFFFFFFu && __ldg(ix.kmer_low + (f - 1) * A + code_at(tp[u])) == want;
            }
            if (good && lastv[u]) {                                                     // '?'-suffixed last super-character
                const unsigned long long ph = (unsigned long long)tp[u] + Ls[u] - 1;
                if (ph >= ix.n) good = false;
                else {
                    const uint32_t from = (Ls[u] - 1) * k - sa[u], f = m - from;
                    const uint32_t want = pat[u].natural(ix, 2 * from, f);
                    const uint32_t code = (in_win[u] && (ph >> 4) != ((pos[u] - 1) >> 4)) ? hat[u] : code_at((uint32_t)ph);
                    good = want != 0xFFFFFFFFu && __ldg(ix.kmer_high + (f - 1) * A + code) == want;
                }
            }
            ok[u] = good;
            // sa[] + the distinct 32-byte sectors of text[] this task needs: the run and the hat character
            n_sect += 1;
            if (d[u]) n_sect += pos[u] >= d[u] ? ((pos[u] - 1) >> 4) - ((pos[u] - d[u]) >> 4) + 1 : 2;
            if (lastv[u] && ((tp[u] + Ls[u] - 1) >> 4) != ((pos[u] == 0 ? 0 : pos[u] - 1) >> 4)) n_sect++;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned em = __ballot_sync(FULL, ok[u]);
            if (em) {
                const uint64_t p = a.first + tk[u].y / k;
                if (ok[u]) atomicAdd(a.counts + p, 1ull);                                 // countOccurrences (:2273)
                if (locate) {
                    uint32_t b = 0;
                    const int leader = __ffs(em) - 1;
                    if ((int)lane == leader) b = atomicAdd(&rs_cnt, (uint32_t)__popc(em));
                    b = __shfl_sync(FULL, b, leader);
                    if (ok[u]) {
                        const uint32_t slot = b + __popc(em & lt);
                        rs_pat[slot] = (uint32_t)p;
                        rs_pos[slot] = (uint64_t)tp[u] * k + sa[u] + ix.initial_n;        // EFMIndex.cpp:2312-2319
                    }
                }
            }
        }
        if (locate && __syncthreads_or(rs_cnt >= SV_FLUSH)) flush();
    }
    if (locate) {
        __syncthreads();
        if (rs_cnt) flush();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_sect += __shfl_xor_sync(FULL, n_sect, d);
    if (lane == 0 && n_sect) atomicAdd(a.stats + ST_GATHER, n_sect);
}

// K_b over the verification records (DevIndex::vrec): the same row task answered with ONE gather — the record holds sa[row], the
// VREC_LEFT super-characters before that text position (nearest first, so the r-th comparison reads a fixed register) and the one
// behind the seed. Only patterns with more than VREC_LEFT super-characters left of the seed read text[] for the rest.
template <int KT, int NW>
__global__ void __launch_bounds__(256, 4) seed_verify_rec_kernel(SeedArgs a) {
    extern __shared__ uint16_t ctab[];
    __shared__ uint32_t rs_pat[SV_CAP];
    __shared__ uint64_t rs_pos[SV_CAP];
    __shared__ uint32_t rs_cnt;
    __shared__ unsigned long long rs_base;
    const DevIndex &ix = a.ix;
    constexpr uint32_t k = KT;
    constexpr uint32_t KMASK = (1u << (2 * KT)) - 1u;
    constexpr int U = SEED_UV;
    for (uint32_t i = threadIdx.x; i < (1u << (2 * KT)); i += blockDim.x) ctab[i] = ix.packed_code[i];
    if (threadIdx.x == 0) rs_cnt = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    const uint32_t m = a.m, j = ix.seed_chars, A = ix.A, n = (uint32_t)ix.n;
    const bool locate = a.mode == E2FM_MODE_LOCATE;
    const unsigned long long nt = *a.task_cursor;
    const uint32_t n_tasks = (uint32_t)(nt > a.task_cap ? 0ull : nt);     // an overflowing chunk is searched again by the host
    unsigned long long n_sect = 0;
    auto flush = [&]() {
        const uint32_t c = rs_cnt;
        if (threadIdx.x == 0) rs_base = atomicAdd(a.stats + ST_RESULTS, (unsigned long long)c);
        __syncthreads();
        const unsigned long long b0 = rs_base;
        uint32_t lost = 0;
        for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) {
            if (b0 + i < a.res_cap) { a.res_pat[b0 + i] = rs_pat[i]; a.res_pos[b0 + i] = rs_pos[i]; }
            else lost++;
        }
        if (lost) atomicAdd(a.stats + ST_RES_OVF, (unsigned long long)lost);
        __syncthreads();
        if (threadIdx.x == 0) rs_cnt = 0;
        __syncthreads();
    };
    const uint32_t stride = gridDim.x * blockDim.x * U;
    for (uint32_t base = blockIdx.x * blockDim.x * U; base < n_tasks; base += stride) {
        uint2 tk[U];
        bool ok[U];
        Sector rec[U];
        PatRegs<NW> pat[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t idx = base + (uint32_t)u * blockDim.x + threadIdx.x;
            ok[u] = idx < n_tasks;
            tk[u] = ok[u] ? __ldg(a.tasks + idx) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {                  // the records and the patterns of the thread's tasks in flight together
            if (ok[u]) {
                rec[u] = ld_sector(ix.vrec + 2 * (size_t)tk[u].x);
                pat[u].load(a.packed + (a.first + tk[u].y / k) * NW);
            }
        }
        uint32_t tp[U], sa[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            sa[u] = tk[u].y % k;
            tp[u] = 0;
            if (!ok[u]) continue;
            uint32_t L;
            bool last_var;
            int l, lo_i;
            shift_geometry(m, k, sa[u], L, last_var, l, lo_i);
            const uint32_t lo = (uint32_t)lo_i;
            // a shift with one fixed super-character fewer than a seed was looked up with every completion of its '?'-suffixed last
            // one (seed_lookup_enum_kernel): the seed then ends with that character, nothing is left to check behind it
            const bool hat_in_seed = last_var && (uint32_t)(l - lo_i) + 1u == j;
            const uint32_t d = hat_in_seed ? lo : (uint32_t)(l - (int)j);     // super-characters 0 .. d-1 lie at text positions pos - d .. pos - 1
            const uint32_t pos = rec[u].lo.x;
            const uint32_t left[5] = {rec[u].lo.y, rec[u].lo.z, rec[u].lo.w, rec[u].hi.x, rec[u].hi.y};
            const uint32_t bit0 = 2u * (d * k - sa[u]);                        // bit offset of super-character d (the seed's first)
            tp[u] = pos >= d ? pos - d : pos + (n - d);
            bool good = true;
            uint32_t first_code = 0xFFFFFFFFu;                                 // text code at tp, when the record holds it
            // text[pos - r] against super-character d - r, r = 1 .. d - lo (the fixed ones, right to left). The VREC_LEFT super-
            // characters left of the seed are cut out of the pattern once, so that the r-th sits at a compile-time offset.
            constexpr int WN = (2 * KT * (int)VREC_LEFT + 31) / 32;
            uint32_t win[WN + 1];
            {
                uint32_t wtmp[WN];
                pat[u].template window<WN>((int)bit0 - 32 * WN, wtmp);
#pragma unroll
                for (int i = 0; i < WN; i++) win[i] = wtmp[i];
                win[WN] = 0;
            }
#pragma unroll
            for (uint32_t r = 1; r <= VREC_LEFT; r++) {
                const uint32_t tv = (left[(r - 1) >> 1] >> (((r - 1) & 1u) * 16u)) & 0xFFFFu;
                constexpr int dummy = 0;
                (void)dummy;
                const int at = 32 * WN - 2 * KT * (int)r;                       // bit offset of super-character d - r in the window
                const uint32_t kmer = __funnelshift_r(win[at >> 5], win[(at >> 5) + 1], (uint32_t)(at & 31)) & KMASK;
                if (r + lo <= d) good &= ctab[kmer] == tv;
                if (r == d) first_code = tv;
            }
            for (uint32_t r = VREC_LEFT + 1; r + lo <= d && good; r++) {       // long patterns: the rest from text[]
                const uint32_t at = pos >= r ? pos - r : pos + (n - r);
                good = ctab[pat[u].bits(bit0 - 2u * r * k, KMASK)] == (uint32_t)__ldg(ix.text + at);
                n_sect++;
            }
            if (good && sa[u] > 0) {                                           // '?'-prefixed first super-character
                const uint32_t f = k - sa[u];
                const uint32_t want = pat[u].natural(ix, 0, f);
                const uint32_t code = d <= VREC_LEFT ? first_code : (uint32_t)__ldg(ix.text + tp[u]);
                good = want != 0xFFFFFFFFu && __ldg(ix.kmer_low + (f - 1) * A + code) == want;
            }
            if (good && last_var && !hat_in_seed) {                            // '?'-suffixed last super-character: text[pos + j]
                uint32_t code = j == 2 ? (rec[u].hi.z & 0xFFFFu) : (j == 3 ? (rec[u].hi.z >> 16) : (rec[u].hi.w & 0xFFFFu));
                if (j > 4) code = (unsigned long long)pos + j < ix.n ? (uint32_t)__ldg(ix.text + pos + j) : 0xFFFFu;
                if (code == 0xFFFFu) good = false;                             // beyond the end of the text
                else {
                    const uint32_t from = (L - 1) * k - sa[u], f = m - from;
                    const uint32_t want = pat[u].natural(ix, 2 * from, f);
                    good = want != 0xFFFFFFFFu && __ldg(ix.kmer_high + (f - 1) * A + code) == want;
                }
            }
            ok[u] = good;
            n_sect++;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned em = __ballot_sync(FULL, ok[u]);
            if (em) {
                const uint64_t p = a.first + tk[u].y / k;
                if (ok[u]) atomicAdd(a.counts + p, 1ull);                                 // countOccurrences (:2273)
                if (locate) {
                    uint32_t b = 0;
                    const int leader = __ffs(em) - 1;
                    if ((int)lane == leader) b = atomicAdd(&rs_cnt, (uint32_t)__popc(em));
                    b = __shfl_sync(FULL, b, leader);
                    if (ok[u]) {
                        const uint32_t slot = b + __popc(em & lt);
                        rs_pat[slot] = (uint32_t)p;
                        rs_pos[slot] = (uint64_t)tp[u] * k + sa[u] + ix.initial_n;        // EFMIndex.cpp:2312-2319
                    }
                }
            }
        }
        if (locate && __syncthreads_or(rs_cnt >= SV_FLUSH)) flush();
    }
    if (locate) {
        __syncthreads();
        if (rs_cnt) flush();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_sect += __shfl_xor_sync(FULL, n_sect, d);
    if (lane == 0 && n_sect) atomicAdd(a.stats + ST_GATHER, n_sect);
}

// Which seed length serves a batch of m-base patterns, and which shifts need their '?'-suffixed last super-character enumerated.
// A shift with nfix fixed super-characters is looked up as it is when nfix >= j (the hat is then checked against the row's record);
// when nfix == j - 1 and it ends with a partial super-character, every k-mer that completes that one is looked up (4^(k - fh)
// candidates for fh known bases), which needs the verification records. Among the seed lengths that serve every shift the
// cheapest wins: lookups (a table beyond the L2 costs four times one inside it) plus the rows they are expected to return.
bool seed_plan(const DevIndex &ix, uint32_t m, uint32_t tables_mask, SeedPlan &out) {
    out = SeedPlan{0, false, 0};
    if (!ix.packed_code || !ix.sa || !ix.text || ix.k < 2 || ix.k > 7 || m == 0 || m > 128) return false;
    const uint32_t k = ix.k;
    double best = 0;
    for (uint32_t j = 2; j <= 8; j++) {
        if (!(tables_mask & (1u << j)) || 2 * k * j > 64) continue;   // seed_lookup_kernel reads a seed as one 64-bit window of the pattern
        bool ok = true, enumerate = false;
        uint64_t cand = 0;
        for (uint32_t sa = 0; sa < k && ok; sa++) {
            uint32_t L;
            bool last_var;
            int l, lo;
            shift_geometry(m, k, sa, L, last_var, l, lo);
            if (l <= 0) continue;                                  // the shift yields nothing (quirk 1)
            const uint32_t nfix = (uint32_t)(l - lo);
            if (nfix >= j) cand += 1;
            else if (last_var && nfix + 1 == j && ix.vrec) {
                const uint32_t fh = (m + sa) % k;
                cand += 1ull << (2 * (k - fh));
                enumerate = true;
            } else ok = false;
        }
        if (!ok || cand == 0 || cand > 8192) continue;
        double seeds = 1, entries = (double)ix.seed_gph * 16.0;
        for (uint32_t i = 0; i < j; i++) seeds *= ix.A;
        for (uint32_t i = 0; i + 1 < j; i++) entries *= ix.A;
        const double rows = (double)ix.n / seeds;                  // expected rows per looked-up seed
        const double cost = (double)cand * ((entries > 64e6 ? 4.0 : 1.0) + 1.5 * rows);
        if (out.j == 0 || cost < best) { best = cost; out = SeedPlan{j, enumerate, (uint32_t)cand}; }
    }
    return out.j != 0;
}

template <class Kern>
static uint32_t resident_ctas(Kern kern, size_t smem, int n_sm) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 2;
    return (uint32_t)per_sm * (uint32_t)n_sm;
}

template <int KT, int NW>
static void launch_seed_nw(const SeedArgs &a, int n_sm, bool verify, cudaStream_t s) {
    constexpr uint32_t PPC = 8 * (32 / KT);
    const size_t smem = (size_t)2 << (2 * KT);
    // persistent grids: as many CTAs as the SMs hold (a multiple of the SM count)
    static const uint32_t lookup_max = resident_ctas(seed_lookup_kernel<KT, NW, 0>, smem, n_sm);
    static const uint32_t verify_max = resident_ctas(seed_verify_kernel<KT, NW>, smem, n_sm);
    static const uint32_t verify_rec_max = resident_ctas(seed_verify_rec_kernel<KT, NW>, smem, n_sm);
    if (!verify && a.lookups) {
        const size_t smem_e = smem + (size_t)a.lookups * 4;
        static const bool opted = cudaFuncSetAttribute(seed_lookup_enum_kernel<KT, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024) == cudaSuccess;
        (void)opted;
        static const uint32_t enum_max = resident_ctas(seed_lookup_enum_kernel<KT, NW>, smem + 4096, n_sm);
        const unsigned long long want = ((unsigned long long)a.count * a.lookups + 255) / 256;
        seed_lookup_enum_kernel<KT, NW><<<(uint32_t)std::max<unsigned long long>(1, std::min<unsigned long long>(want, enum_max)), 256, smem_e, s>>>(a);
    } else if (!verify) {
        const uint32_t groups = ((a.count + PPC - 1) / PPC + SEED_U - 1) / SEED_U;
        const uint32_t blocks = std::max(1u, std::min(groups, lookup_max));
        switch (a.ix.seed_chars) {                    // the usual seed lengths get their own instance (unrolled code loop)
            case 2: seed_lookup_kernel<KT, NW, 2><<<blocks, 256, smem, s>>>(a); break;
            case 3: seed_lookup_kernel<KT, NW, 3><<<blocks, 256, smem, s>>>(a); break;
            case 4: seed_lookup_kernel<KT, NW, 4><<<blocks, 256, smem, s>>>(a); break;
            default: seed_lookup_kernel<KT, NW, 0><<<blocks, 256, smem, s>>>(a); break;
        }
    } else if (a.ix.vrec) {
        seed_verify_rec_kernel<KT, NW><<<verify_rec_max, 256, smem, s>>>(a);
    } else {
        seed_verify_kernel<KT, NW><<<verify_max, 256, smem, s>>>(a);
    }
}

template <int KT>
static void launch_seed_kt(const SeedArgs &a, uint32_t nw, int n_sm, bool verify, cudaStream_t s) {
    if (nw == 1) launch_seed_nw<KT, 1>(a, n_sm, verify, s);
    else launch_seed_nw<KT, 2>(a, n_sm, verify, s);
}

static void launch_seed_any(const SeedArgs &a, uint32_t nw, int n_sm, bool verify, cudaStream_t s) {
    if (a.count == 0) return;
    switch (a.ix.k) {
        case 2: launch_seed_kt<2>(a, nw, n_sm, verify, s); break;
        case 3: launch_seed_kt<3>(a, nw, n_sm, verify, s); break;
        case 4: launch_seed_kt<4>(a, nw, n_sm, verify, s); break;
        case 5: launch_seed_kt<5>(a, nw, n_sm, verify, s); break;
        case 6: launch_seed_kt<6>(a, nw, n_sm, verify, s); break;
        case 7: launch_seed_kt<7>(a, nw, n_sm, verify, s); break;
        default: break;
    }
}
// seed lookup (row tasks), then verification (counts, located occurrences)
void launch_seed_lookup(const SeedArgs &a, uint32_t nw, int n_sm, cudaStream_t s) { launch_seed_any(a, nw, n_sm, false, s); }
void launch_seed_verify(const SeedArgs &a, uint32_t nw, int n_sm, cudaStream_t s) { launch_seed_any(a, nw, n_sm, true, s); }

// ---------------------------------------------------------------- sub-batch for the generic path

// out[i*m .. (i+1)*m) = ASCII of pattern sel[i]: copied from the ASCII batch when there is one, else decoded from the packed form
__global__ void __launch_bounds__(256) materialize_kernel(const uint32_t *__restrict__ sel, uint32_t n_sel, uint32_t m, uint32_t nw,
                                                           const uint8_t *__restrict__ ascii, const unsigned long long *__restrict__ packed,
                                                           uint8_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sel) return;
    const uint64_t p = sel[i];
    if (ascii) {
        for (uint32_t q = 0; q < m; q++) out[(uint64_t)i * m + q] = ascii[p * m + q];
    } else {
        for (uint32_t q = 0; q < m; q++) {
            const unsigned long long w = packed[p * 2 * nw + (q >> 5)];
            out[(uint64_t)i * m + q] = (uint8_t)"ACGT"[(w >> (2 * (q & 31u))) & 3u];
        }
    }
}
void launch_materialize(const uint32_t *sel, uint32_t n_sel, uint32_t m, uint32_t nw, const uint8_t *ascii, const uint4 *packed,
                        uint8_t *out, cudaStream_t s) {
    if (n_sel) materialize_kernel<<<(n_sel + 255) / 256, 256, 0, s>>>(sel, n_sel, m, nw, ascii,
                                                                    reinterpret_cast<const unsigned long long *>(packed), out);
}

// after the generic path searched the sub-batch with local pattern numbers: counts and located occurrences back under
// the batch's own pattern numbers
__global__ void __launch_bounds__(256) sub_merge_kernel(const uint32_t *__restrict__ sel, uint32_t n_sel,
                                                         const unsigned long long *__restrict__ sub_counts,
                                                         unsigned long long *__restrict__ counts, uint32_t *__restrict__ res_pat,
                                                         uint64_t r0, uint64_t r1) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_sel) counts[sel[i]] = sub_counts[i];
    if (r0 + i < r1) res_pat[r0 + i] = sel[res_pat[r0 + i]];
}
void launch_sub_merge(const uint32_t *sel, uint32_t n_sel, const unsigned long long *sub_counts, unsigned long long *counts,
                      uint32_t *res_pat, uint64_t r0, uint64_t r1, cudaStream_t s) {
    const uint64_t work = std::max<uint64_t>(n_sel, r1 - r0);
    if (work) sub_merge_kernel<<<(uint32_t)((work + 255) / 256), 256, 0, s>>>(sel, n_sel, sub_counts, counts, res_pat, r0, r1);
}

// u64 -> u32 (the compact result: counts and offsets of a batch fit 32 bits; the host checks the totals)
__global__ void __launch_bounds__(256) narrow_kernel(const unsigned long long *__restrict__ in, uint32_t *__restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}
void launch_narrow(const unsigned long long *in, uint32_t *out, uint64_t n, cudaStream_t s) {
    if (n) narrow_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(in, out, n);
}

// test hook: the seed table against the rank structure — interval of `count` seeds given as j codes each
__global__ void debug_seed_kernel(DevIndex ix, const uint32_t *__restrict__ codes, uint64_t count, uint32_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t j = ix.seed_chars;
    uint32_t hi = 0, cl = 0;
    for (uint32_t q = 0; q < j; q++) {
        const uint32_t c = codes[i * j + q];
        if (q + 1 < j) hi = hi * ix.A + c; else cl = c;
    }
    uint32_t slot;
    const size_t sector = seed_sector(ix, hi, cl, slot);
    const SeedHit h = seed_decode(seed_load(ix, sector), slot);
    out[3 * i] = h.s;
    out[3 * i + 1] = h.cnt;
    out[3 * i + 2] = h.exact ? 1u : 0u;
}
void launch_debug_seed(const DevIndex &ix, const uint32_t *codes, uint64_t count, uint32_t *out, cudaStream_t s) {
    if (count) debug_seed_kernel<<<(uint32_t)((count + 255) / 256), 256, 0, s>>>(ix, codes, count, out);
}

}  // namespace e2fm
