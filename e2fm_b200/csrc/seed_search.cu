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

// ---------------------------------------------------------------- the search

// bits [off, off + width) of the pattern's bit string (width <= 32)
template <int NW>
__device__ __forceinline__ uint32_t pk_field(const unsigned long long (&pk)[2 * NW], uint32_t off, uint32_t mask) {
    const uint32_t wi = off >> 6, sh = off & 63u;
    unsigned long long a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < 2 * NW; i++) {
        if (wi == (uint32_t)i) a = pk[i];
        if (wi + 1 == (uint32_t)i) b = pk[i];
    }
    unsigned long long v = a >> sh;
    if (sh) v |= b << (64 - sh);
    return (uint32_t)v & mask;
}

// natural value (base n_sym, first base most significant) of the f bases at bit offset off: what kmer_low / kmer_high hold
template <int NW>
__device__ __forceinline__ uint32_t pk_natural(const DevIndex &ix, const unsigned long long (&pk)[2 * NW], uint32_t off, uint32_t f) {
    const uint32_t bits = pk_field<NW>(pk, off, f >= 16 ? 0xFFFFFFFFu : ((1u << (2 * f)) - 1u));
    uint32_t v = 0;
    bool bad = false;
    for (uint32_t i = 0; i < f; i++) {
        const uint32_t s = ix.sym4[(bits >> (2 * i)) & 3u];
        bad |= s == 0xFFu;
        v = v * ix.n_sym + s;
    }
    return bad ? 0xFFFFFFFFu : v;
}

template <int KT, int NW>
__global__ void __launch_bounds__(256) seed_search_kernel(SeedArgs a) {
    extern __shared__ uint16_t ctab[];                 // [4^k] packed k-mer -> compact code
    __shared__ uint32_t warp_tot[8];
    __shared__ unsigned long long blk_base;
    const DevIndex &ix = a.ix;
    constexpr uint32_t k = KT;
    constexpr uint32_t KMASK = (1u << (2 * KT)) - 1u;
    for (uint32_t i = threadIdx.x; i < (1u << (2 * KT)); i += blockDim.x) ctab[i] = ix.packed_code[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < a.count;
    const uint64_t p = a.first + t;
    const uint32_t m = a.m, j = ix.seed_chars, A = ix.A;
    const uint32_t blk_shift = ix.blk_shift, blk_mask = (1u << blk_shift) - 1;
    const long long n = (long long)ix.n;

    unsigned long long pk[2 * NW];
#pragma unroll
    for (int i = 0; i < 2 * NW; i++) pk[i] = 0;
    uint32_t flag = 2;
    if (valid) {
        const uint4 *src = a.packed + p * NW;
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const uint4 v = __ldg(src + i);
            pk[2 * i] = (unsigned long long)v.x | ((unsigned long long)v.y << 32);
            pk[2 * i + 1] = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
        }
        flag = a.pflag ? a.pflag[p] : 0;
    }
    // the packed form holds nothing beyond base m-1
    uint32_t cnt_total = 0, nres = 0, n_sect = 0, n_rank = 0;
    unsigned long long respos[KT];
    bool complex = flag == 1;

    if (flag == 0) {
        // ---- EFMIndex::computeSuperPatterns (:2055-2207) + the seed of every shift: k independent sector gathers in flight
        Sector sec[KT];
        uint32_t slot[KT];
        int nexti[KT];        // next super-character to process, -2: the shift is dead
#pragma unroll
        for (int sa = 0; sa < KT; sa++) {
            const uint32_t L = (m + sa + k - 1) / k;
            const bool last_var = (m + sa) % k != 0;
            int l = -1;
            if (!last_var) { if (!(L == 1 && sa > 0)) l = (int)L; }
            else if (L > 1 && (L > 2 || sa == 0)) l = (int)L - 1;
            nexti[sa] = -2;
            slot[sa] = 0;
            if (l > 0) {                               // the host only takes this path when every such shift has >= j fixed characters
                uint32_t hi = 0, cl = 0;
                bool absent = false;
                for (uint32_t q = 0; q < j; q++) {
                    const uint32_t c = (uint32_t)l - j + q;
                    const uint32_t code = ctab[pk_field<NW>(pk, 2 * (c * k - sa), KMASK)];
                    absent |= code == 0xFFFFu;
                    if (q + 1 < j) hi = hi * A + code; else cl = code;
                }
                if (!absent) {
                    sec[sa] = seed_load(ix, seed_sector(ix, hi, cl, slot[sa]));
                    nexti[sa] = l - (int)j - 1;
                    n_sect++;
                }
            }
        }
#pragma unroll
        for (int sa = 0; sa < KT; sa++) {
            if (nexti[sa] == -2 || complex) continue;
            const SeedHit h = seed_decode(sec[sa], slot[sa]);
            if (!h.exact) { complex = true; continue; }
            if (h.cnt == 0) continue;
            const uint32_t L = (m + sa + k - 1) / k;
            const bool last_var = (m + sa) % k != 0;
            const int l = last_var ? (int)L - 1 : (int)L;
            const int lo = sa > 0 ? 1 : 0;
            uint32_t s = h.s, e = h.s + h.cnt;
            int i = nexti[sa];
            // ---- EFMIndex::backwardStep (:1966-2001) over rank sectors while several rows are left
            while (e - s > 1 && i >= lo) {
                const uint32_t cc = ctab[pk_field<NW>(pk, 2 * ((uint32_t)i * k - sa), KMASK)];
                if (cc == 0xFFFFu) { e = s; break; }
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
            }
            if (e <= s) continue;
            if (e - s > 1) { complex = true; continue; }     // several rows at the end: expanded and resolved by the generic path
            // ---- one row left: its suffix starts at text position pos; super-character c of the shift lies at pos - (i + 1 - c)
            const long long pos = (long long)__ldg(ix.sa + s);
            n_sect++;
            auto at = [&](long long x) -> uint32_t {
                if (x < 0) x += n; else if (x >= n) x -= n;
                return (uint32_t)__ldg(ix.text + x);
            };
            bool ok = true;
            for (int c = i; c >= lo; c--) {
                const uint32_t code = ctab[pk_field<NW>(pk, 2 * ((uint32_t)c * k - sa), KMASK)];
                ok &= code == at(pos - (long long)(i + 1 - c));
            }
            long long tp = pos - (long long)(i + 1);          // text position of super-character 0 of the shift
            if (tp < 0) tp += n;
            if (sa > 0) {
                // EFMIndex::backwardStepIfMatch (:2011-2041): the k - sa fixed symbols of the '?'-prefixed first super-character
                const uint32_t f = k - sa;
                const uint32_t want = pk_natural<NW>(ix, pk, 0, f);
                const uint32_t code = (uint32_t)__ldg(ix.text + tp);
                ok &= want != 0xFFFFFFFFu && __ldg(ix.kmer_low + (f - 1) * A + code) == want;
            }
            if (last_var) {
                // EFMIndex::findWholePatternMatches (:1862-1899): the fixed symbols of the '?'-suffixed last super-character
                const long long ph = tp + (long long)L - 1;
                if (ph >= n) ok = false;
                else {
                    const uint32_t from = (L - 1) * k - sa, f = m - from;
                    const uint32_t want = pk_natural<NW>(ix, pk, 2 * from, f);
                    const uint32_t code = (uint32_t)__ldg(ix.text + ph);
                    ok &= want != 0xFFFFFFFFu && __ldg(ix.kmer_high + (f - 1) * A + code) == want;
                }
            }
            {   // distinct 32-byte sectors of text[] this touched: the run [tp or tp+lo, pos-1] and the hat character
                const long long a0 = (sa > 0 ? tp : pos - (long long)(i + 1 - lo)), a1 = pos - 1;
                if (a1 >= a0 && a0 >= 0) n_sect += (uint32_t)((a1 >> 4) - (a0 >> 4) + 1); else n_sect += 2;
                if (last_var && ((tp + (long long)L - 1) >> 4) != ((pos - 1) >> 4)) n_sect++;
            }
            if (ok) {
                cnt_total++;
                respos[nres++] = (unsigned long long)tp * k + sa + ix.initial_n;      // EFMIndex.cpp:2312-2319
            }
        }
    }
    if (complex) { cnt_total = 0; nres = 0; }
    if (valid) {
        a.counts[p] = cnt_total;                               // countOccurrences (:2273); the generic path adds to it for list members
        if (complex) {
            const unsigned long long at = atomicAdd(a.stats + ST_COMPLEX, 1ull);
            a.complex_list[at] = (uint32_t)p;
        }
    }
    // ---- located occurrences: one reservation per CTA
    if (a.mode == E2FM_MODE_LOCATE) {
        uint32_t incl = nres;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, d);
            if ((int)lane >= d) incl += v;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint32_t v = warp_tot[w];
            if ((uint32_t)w < warp) before += v;
            total += v;
        }
        if (threadIdx.x == 0 && total) blk_base = atomicAdd(a.stats + ST_RESULTS, (unsigned long long)total);
        __syncthreads();
        if (nres) {
            unsigned long long slot0 = blk_base + before + incl - nres;
            uint32_t lost = 0;
#pragma unroll
            for (int q = 0; q < KT; q++) {
                if ((uint32_t)q < nres) {
                    if (slot0 + q < a.res_cap) { a.res_pat[slot0 + q] = (uint32_t)p; a.res_pos[slot0 + q] = respos[q]; }
                    else lost++;
                }
            }
            if (lost) atomicAdd(a.stats + ST_RES_OVF, (unsigned long long)lost);
        }
    }
    unsigned long long q = n_sect, r = n_rank;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { q += __shfl_xor_sync(FULL, q, d); r += __shfl_xor_sync(FULL, r, d); }
    if (lane == 0) {
        if (q) atomicAdd(a.stats + ST_SECT_BS, q);
        if (r) atomicAdd(a.stats + ST_RANK, r);
    }
}

bool seed_search_supported(const DevIndex &ix, uint32_t m) {
    if (!ix.seed_tab || !ix.packed_code || !ix.sa || !ix.text || ix.k < 2 || ix.k > 7 || m == 0 || m > 128) return false;
    const uint32_t k = ix.k, j = ix.seed_chars;
    for (uint32_t sa = 0; sa < k; sa++) {
        const uint32_t L = (m + sa + k - 1) / k;
        const bool last_var = (m + sa) % k != 0;
        int l = -1;
        if (!last_var) { if (!(L == 1 && sa > 0)) l = (int)L; }
        else if (L > 1 && (L > 2 || sa == 0)) l = (int)L - 1;
        if (l <= 0) continue;                                  // the shift yields nothing (quirk 1)
        const int lo = sa > 0 ? 1 : 0;
        if (l - lo < (int)j) return false;                     // fewer fixed super-characters than a seed
    }
    return true;
}

template <int KT>
static void launch_seed_kt(const SeedArgs &a, uint32_t nw, cudaStream_t s) {
    const uint32_t blocks = (a.count + 255) / 256;
    const size_t smem = (size_t)2 << (2 * KT);
    if (nw == 1) seed_search_kernel<KT, 1><<<blocks, 256, smem, s>>>(a);
    else seed_search_kernel<KT, 2><<<blocks, 256, smem, s>>>(a);
}

void launch_seed_search(const SeedArgs &a, uint32_t nw, cudaStream_t s) {
    if (a.count == 0) return;
    switch (a.ix.k) {
        case 2: launch_seed_kt<2>(a, nw, s); break;
        case 3: launch_seed_kt<3>(a, nw, s); break;
        case 4: launch_seed_kt<4>(a, nw, s); break;
        case 5: launch_seed_kt<5>(a, nw, s); break;
        case 6: launch_seed_kt<6>(a, nw, s); break;
        case 7: launch_seed_kt<7>(a, nw, s); break;
        default: break;
    }
}

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
