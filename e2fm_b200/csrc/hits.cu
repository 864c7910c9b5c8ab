// hits.cu — result pipeline of ONE pipeline stage of e2fm_search_hits (api.cu): what EFMCollection::search returns for the
// patterns of the stage — per-pattern occurrence offsets, Occurrence::sequenceIndex and Occurrence::patternPosition
// (FMCollection.h:23-44, EFMCollection.cpp:124-148) — written in 32 bits at the batch-wide place of the stage, with no host round
// trip: every count the host would need (located occurrences of the stage, big segments, the running total of the batch) is read
// on the device from the stage's own statistics block.
//
//   hits_block_sums / hits_scan_apply   exclusive scan of the stage's counts -> its CSR; scatter cursors cleared on the way
//   hits_scatter_kernel     (pattern, position) pairs of the stage -> per-pattern segments
//   hits_sortfin_kernel     per pattern: sort (EFMIndex::locateOccurrences' std::sort, EFMIndex.cpp:2291), positions -> (sequence,
//                           offset in the sequence), 32-bit offsets and counts at the stage's place in the batch-wide arrays
//   hits_big_tail_kernel    segments too long for one thread, a CTA each; the last CTA adds the stage to the batch's running total and
//                           writes {first occurrence, occurrences, trouble flags} into page-locked host memory
#include "search.cuh"

#include <algorithm>

#include "../../include/e2fm_b200.h"

namespace e2fm {

__global__ void __launch_bounds__(256) hits_scatter_kernel(const uint32_t *__restrict__ res_pat, const uint64_t *__restrict__ res_pos,
                                                            const unsigned long long *__restrict__ stats, unsigned long long res_cap,
                                                            uint64_t first, const uint64_t *__restrict__ off, uint32_t *__restrict__ cursor,
                                                            uint64_t *__restrict__ positions) {
    const unsigned long long have = stats[ST_RESULTS];
    const unsigned long long n_res = have < res_cap ? have : res_cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_res; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t pl = (uint32_t)(res_pat[i] - first);
        const uint64_t slot = off[pl] + atomicAdd(cursor + pl, 1u);
        positions[slot] = res_pos[i];
    }
}

void launch_hits_scatter(const uint32_t *res_pat, const uint64_t *res_pos, const unsigned long long *stats, uint64_t res_cap, uint64_t first,
                         const uint64_t *off, uint32_t *cursor, uint64_t *positions, int n_sm, cudaStream_t s) {
    const uint64_t want = (res_cap + 255) / 256;
    hits_scatter_kernel<<<(uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)n_sm * 8)), 256, 0, s>>>(
        res_pat, res_pos, stats, res_cap, first, off, cursor, positions);
}

// (sequence, offset in the sequence) of a text position (EFMCollection.cpp:124-148: sequence = the first terminator beyond it)
__device__ __forceinline__ void map_position(const DevIndex &ix, uint64_t pos, uint32_t &seq, uint32_t &off) {
    uint32_t lo = 0, hi = ix.n_term;                              // first index with terminators[idx] > pos
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint64_t)__ldg(ix.terminators + mid) > pos) hi = mid;
        else lo = mid + 1;
    }
    const bool found = lo < ix.n_term;
    uint64_t start = 0;
    if (found && lo > 0) start = (uint64_t)__ldg(ix.terminators + lo - 1) + ix.k;
    seq = found ? lo : 0xFFFFFFFFu;                               // the reference leaves sequenceIndex = -1
    off = (uint32_t)(pos - start);
}

// exclusive scan of the stage's counts, second half: the prefix of a block comes straight from the block sums (a few hundred: no
// separate launch for their scan), the per-pattern scatter cursors are cleared on the way
static constexpr int HS_T = 256, HS_E = 8, HS_B = HS_T * HS_E;

__global__ void __launch_bounds__(HS_T) hits_block_sums_kernel(const unsigned long long *__restrict__ counts, uint32_t cnt, unsigned long long *__restrict__ sums) {
    __shared__ unsigned long long ws[HS_T / 32];
    const uint32_t base = blockIdx.x * HS_B;
    unsigned long long v = 0;
    for (int j = 0; j < HS_E; j++) {
        const uint32_t i = base + j * HS_T + threadIdx.x;
        if (i < cnt) v += counts[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < HS_T / 32; w++) t += ws[w];
        sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(HS_T) hits_scan_apply_kernel(const unsigned long long *__restrict__ counts, uint32_t cnt, const unsigned long long *__restrict__ sums,
                                                                uint32_t nb, uint64_t *__restrict__ off, uint32_t *__restrict__ cursor) {
    __shared__ unsigned long long ws[HS_T / 32], red[2][HS_T / 32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long pre = 0, all = 0;
    for (uint32_t i = threadIdx.x; i < nb; i += HS_T) {
        const unsigned long long v = sums[i];
        all += v;
        if (i < blockIdx.x) pre += v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { pre += __shfl_xor_sync(0xFFFFFFFFu, pre, d); all += __shfl_xor_sync(0xFFFFFFFFu, all, d); }
    if (lane == 0) { red[0][warp] = pre; red[1][warp] = all; }
    __syncthreads();
    unsigned long long block_prefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < HS_T / 32; w++) { block_prefix += red[0][w]; total += red[1][w]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) off[cnt] = total;
    const uint32_t base = blockIdx.x * HS_B + threadIdx.x * HS_E;
    unsigned long long v[HS_E], tot = 0;
#pragma unroll
    for (int j = 0; j < HS_E; j++) {
        v[j] = base + j < cnt ? counts[base + j] : 0;
        tot += v[j];
    }
    unsigned long long inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if ((int)lane >= d) inc += t;
    }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    unsigned long long run = block_prefix + inc - tot;
    for (uint32_t w = 0; w < warp; w++) run += ws[w];
#pragma unroll
    for (int j = 0; j < HS_E; j++) {
        if (base + j < cnt) { off[base + j] = run; cursor[base + j] = 0; }
        run += v[j];
    }
}

void launch_hits_scan(const unsigned long long *counts, uint32_t cnt, uint64_t *off, uint32_t *cursor, unsigned long long *scratch, cudaStream_t s) {
    if (cnt == 0) { cudaMemsetAsync(off, 0, 8, s); return; }
    const uint32_t nb = (cnt + HS_B - 1) / HS_B;
    hits_block_sums_kernel<<<nb, HS_T, 0, s>>>(counts, cnt, scratch);
    hits_scan_apply_kernel<<<nb, HS_T, 0, s>>>(counts, cnt, scratch, nb, off, cursor);
}

// sorted segment -> unique head, by the whole CTA (positions of one pattern are distinct on the seed path; kept for the contract)
__device__ __forceinline__ uint32_t unique_head(uint64_t *q, uint32_t len) {
    uint32_t w = 0;
    for (uint32_t i = 0; i < len; i++)
        if (i == 0 || q[i] != q[w - 1]) q[w++] = q[i];
    return w;
}

static constexpr uint32_t HITS_SORT_SMALL = 24;     // segments up to this length are sorted by one thread

// one thread per pattern of the stage: count and offset in 32 bits at the batch-wide place; its occurrences sorted (std::sort,
// EFMIndex.cpp:2291; a segment of up to HITS_SORT_SMALL entries by insertion, longer ones are listed for hits_big_tail_kernel),
// checked for adjacent duplicates (:2296-2299) and written as (sequence, offset) at *run_total + off[i]
__global__ void __launch_bounds__(256) hits_sortfin_kernel(DevIndex ix, HitsOut o) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= o.cnt) return;
    const unsigned long long c = o.counts[o.first + i];
    o.cnt32[o.first + i] = c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c;
    if (!o.off) return;
    const unsigned long long base = *o.run_total, total = o.off[o.cnt];
    const uint64_t a = o.off[i];
    o.off32[o.first + i] = (uint32_t)(base + a);
    if (total > o.res_cap || base + total > o.out_cap) return;        // reported by the tail; nothing is written past a store
    const uint64_t len = o.off[i + 1] - a;
    if (len == 0) return;
    uint64_t *q = o.positions + a;
    if (len > HITS_SORT_SMALL) {
        const unsigned long long at = atomicAdd(o.stats + ST_BIGSEG, 1ull);
        o.big_list[at] = i;
        return;
    }
    for (uint32_t x = 1; x < (uint32_t)len; x++) {
        const uint64_t v = q[x];
        uint32_t y = x;
        while (y > 0 && q[y - 1] > v) { q[y] = q[y - 1]; y--; }
        q[y] = v;
    }
    uint32_t dups = 0;
    for (uint32_t x = 0; x < (uint32_t)len; x++) {
        const uint64_t pos = q[x];
        if (x && pos == q[x - 1]) dups++;
        uint32_t sq, so;
        map_position(ix, pos, sq, so);
        o.seq32[base + a + x] = sq;
        o.soff32[base + a + x] = so;
    }
    if (dups) atomicAdd(o.stats + ST_DUPS, (unsigned long long)dups);
}

// the long segments, one CTA each (bitonic network in global memory), then the stage's report: the last CTA to finish adds the
// stage to the batch's running total and writes {first occurrence, occurrences, trouble flags} into page-locked host memory
__global__ void __launch_bounds__(256) hits_big_tail_kernel(DevIndex ix, HitsOut o) {
    __shared__ bool last;
    const unsigned long long base = *o.run_total;
    const unsigned long long total = o.off ? o.off[o.cnt] : 0ull;
    const bool fits = o.off && total <= o.res_cap && base + total <= o.out_cap;
    const unsigned long long n_big = fits ? o.stats[ST_BIGSEG] : 0ull;
    for (unsigned long long b = blockIdx.x; b < n_big; b += gridDim.x) {
        const uint32_t p = o.big_list[b];
        const uint64_t a = o.off[p];
        const uint32_t len = (uint32_t)(o.off[p + 1] - a);
        uint64_t *q = o.positions + a;
        uint32_t N = 1;                                 // ascending-only bitonic network: entries at index >= len behave as +infinity
        while (N < len) N <<= 1;
        for (uint32_t kk = 2; kk <= N; kk <<= 1) {
            for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
                for (uint32_t t = threadIdx.x; t < N / 2; t += blockDim.x) {
                    uint32_t x, l;
                    if (j == (kk >> 1)) {
                        const uint32_t blk = t / j, pos = t - blk * j;
                        x = blk * kk + pos;
                        l = blk * kk + kk - 1 - pos;
                    } else {
                        x = 2 * j * (t / j) + (t % j);
                        l = x + j;
                    }
                    if (l < len) {
                        const uint64_t u = q[x], v = q[l];
                        if (u > v) { q[x] = v; q[l] = u; }
                    }
                }
                __syncthreads();
            }
        }
        uint32_t dups = 0;
        for (uint32_t x = threadIdx.x; x < len; x += blockDim.x) {
            const uint64_t pos = q[x];
            if (x && pos == q[x - 1]) dups++;
            uint32_t sq, so;
            map_position(ix, pos, sq, so);
            o.seq32[base + a + x] = sq;
            o.soff32[base + a + x] = so;
        }
        if (dups) atomicAdd(o.stats + ST_DUPS, (unsigned long long)dups);
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(o.stats + ST_QUEUE, 1ull) + 1 == gridDim.x;     // (the seed path does not use the queue word)
    __syncthreads();
    if (!last || threadIdx.x != 0) return;
    __threadfence();
    unsigned long long flags = 0;
    if (*o.task_count > o.task_cap) flags |= HITS_TASK_OVF;
    if (o.stats[ST_RES_OVF] || (o.off && o.stats[ST_RESULTS] > o.res_cap) || total > o.res_cap) flags |= HITS_RES_OVF;
    if (*(volatile unsigned long long *)(o.stats + ST_DUPS)) flags |= HITS_DUPS;
    if (o.stats[ST_COMPLEX]) flags |= HITS_COMPLEX;
    if (base + total > o.out_cap) flags |= HITS_OUT_OVF;
    *o.run_total = base + total;
    if (o.off32_end) *o.off32_end = (uint32_t)(base + total);
    o.info->base = base;
    o.info->total = total;
    o.info->tasks = *o.task_count;
    o.info->flags = flags;
    __threadfence_system();
}

void launch_hits_finish(const DevIndex &ix, const HitsOut &o, int n_sm, cudaStream_t s) {
    if (o.cnt) hits_sortfin_kernel<<<(o.cnt + 255) / 256, 256, 0, s>>>(ix, o);
    hits_big_tail_kernel<<<(uint32_t)std::max(1, n_sm / 4), 256, 0, s>>>(ix, o);
}

}  // namespace e2fm
