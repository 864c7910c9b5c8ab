// hits.cu — result pipeline of ONE pipeline stage of e2fm_search_hits (api.cu): what EFMCollection::search returns for the
// patterns of the stage — per-pattern occurrence offsets, Occurrence::sequenceIndex and Occurrence::patternPosition
// (FMCollection.h:23-44, EFMCollection.cpp:124-148) — written in 32 bits at the batch-wide place of the stage, with no host round
// trip: every count the host would need (located occurrences of the stage, big segments, the running total of the batch) is read
// on the device from the stage's own statistics block.
//
//   hits_scatter_kernel     (pattern, position) pairs of the stage -> per-pattern segments (CSR by the scan of the counts)
//   hits_sort_big_kernel    segments too long for one thread (EFMIndex::locateOccurrences' std::sort, EFMIndex.cpp:2291), CTA each
//   hits_finalize_kernel    offsets + counts in 32 bits at the stage's place; every position -> (sequence, offset in the sequence)
//   hits_tail_kernel        running total of the batch; the stage's {first occurrence, occurrences, trouble flags} for the host
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

// sorted segment -> unique head, by the whole CTA (positions of one pattern are distinct on the seed path; kept for the contract)
__device__ __forceinline__ uint32_t unique_head(uint64_t *q, uint32_t len) {
    uint32_t w = 0;
    for (uint32_t i = 0; i < len; i++)
        if (i == 0 || q[i] != q[w - 1]) q[w++] = q[i];
    return w;
}

static constexpr uint32_t HITS_SORT_SMALL = 24;     // segments up to this length are sorted by one thread

// one thread per pattern: insertion sort of its positions (std::sort, EFMIndex.cpp:2291) + adjacent de-duplication (:2296-2299);
// longer segments are listed for hits_sort_big_kernel. Nothing is touched when the stage's result store overflowed.
__global__ void __launch_bounds__(256) hits_sort_small_kernel(uint64_t *positions, const uint64_t *__restrict__ off, uint32_t cnt,
                                                               unsigned long long res_cap, uint32_t *big_list, uint32_t *ulen,
                                                               unsigned long long *stats) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cnt || off[cnt] > res_cap) return;
    const uint64_t a = off[p];
    const uint64_t len = off[p + 1] - a;
    if (len <= 1) { ulen[p] = (uint32_t)len; return; }
    if (len > HITS_SORT_SMALL) {
        const unsigned long long at = atomicAdd(stats + ST_BIGSEG, 1ull);
        big_list[at] = p;
        return;
    }
    uint64_t *q = positions + a;
    for (uint32_t i = 1; i < (uint32_t)len; i++) {
        const uint64_t v = q[i];
        uint32_t j = i;
        while (j > 0 && q[j - 1] > v) { q[j] = q[j - 1]; j--; }
        q[j] = v;
    }
    const uint32_t u = unique_head(q, (uint32_t)len);
    ulen[p] = u;
    if (u != len) atomicAdd(stats + ST_DUPS, (unsigned long long)(len - u));
}

__global__ void __launch_bounds__(256) hits_sort_big_kernel(uint64_t *positions, const uint64_t *__restrict__ off, uint32_t cnt,
                                                             unsigned long long res_cap, const uint32_t *__restrict__ big_list,
                                                             uint32_t *ulen, unsigned long long *stats) {
    if (off[cnt] > res_cap) return;
    const unsigned long long n_big = stats[ST_BIGSEG];
    for (unsigned long long b = blockIdx.x; b < n_big; b += gridDim.x) {
        const uint32_t p = big_list[b];
        const uint64_t a = off[p];
        const uint32_t len = (uint32_t)(off[p + 1] - a);
        uint64_t *q = positions + a;
        // ascending-only bitonic network in global memory: entries at index >= len behave as +infinity
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
                        const uint64_t x = q[i], y = q[l];
                        if (x > y) { q[i] = y; q[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
        if (threadIdx.x == 0) {
            const uint32_t u = unique_head(q, len);
            ulen[p] = u;
            if (u != len) atomicAdd(stats + ST_DUPS, (unsigned long long)(len - u));
        }
        __syncthreads();
    }
}

void launch_hits_sort(uint64_t *positions, const uint64_t *off, uint32_t cnt, uint64_t res_cap, uint32_t *big_list, uint32_t *ulen,
                      unsigned long long *stats, int n_sm, cudaStream_t s) {
    if (cnt == 0) return;
    hits_sort_small_kernel<<<(cnt + 255) / 256, 256, 0, s>>>(positions, off, cnt, res_cap, big_list, ulen, stats);
    hits_sort_big_kernel<<<(uint32_t)n_sm, 256, 0, s>>>(positions, off, cnt, res_cap, big_list, ulen, stats);
}

// offsets and counts of the stage's patterns, and (sequence, offset) of its occurrences (EFMCollection.cpp:124-148: sequence = the
// first terminator beyond the position), at the batch-wide places: pattern first + i, occurrence *run_total + q
__global__ void __launch_bounds__(256) hits_finalize_kernel(DevIndex ix, const unsigned long long *__restrict__ counts, uint64_t first, uint32_t cnt,
                                                             const uint64_t *__restrict__ off, const uint64_t *__restrict__ positions,
                                                             const unsigned long long *__restrict__ run_total, unsigned long long out_cap,
                                                             unsigned long long res_cap, uint32_t *__restrict__ cnt32, uint32_t *__restrict__ off32,
                                                             uint32_t *__restrict__ seq32, uint32_t *__restrict__ soff32) {
    const unsigned long long base = *run_total;
    const unsigned long long total = off ? off[cnt] : 0ull;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride) {
        const unsigned long long c = counts[first + i];
        cnt32[first + i] = c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c;
        if (off) off32[first + i] = (uint32_t)(base + off[i]);
    }
    if (!off || total > res_cap || base + total > out_cap) return;    // the tail kernel reports the overflow
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const uint64_t pos = positions[q];
        uint32_t lo = 0, hi = ix.n_term;                              // first index with terminators[idx] > pos
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((uint64_t)__ldg(ix.terminators + mid) > pos) hi = mid;
            else lo = mid + 1;
        }
        const bool found = lo < ix.n_term;
        uint64_t start = 0;
        if (found && lo > 0) start = (uint64_t)__ldg(ix.terminators + lo - 1) + ix.k;
        seq32[base + q] = found ? lo : 0xFFFFFFFFu;                   // the reference leaves sequenceIndex = -1
        soff32[base + q] = (uint32_t)(pos - start);
    }
}

void launch_hits_finalize(const DevIndex &ix, const unsigned long long *counts, uint64_t first, uint32_t cnt, const uint64_t *off,
                          const uint64_t *positions, const unsigned long long *run_total, uint64_t out_cap, uint64_t res_cap, uint32_t *cnt32,
                          uint32_t *off32, uint32_t *seq32, uint32_t *soff32, int n_sm, cudaStream_t s) {
    if (cnt == 0) return;
    const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>((cnt + 255) / 256, (uint32_t)n_sm * 8));
    hits_finalize_kernel<<<blocks, 256, 0, s>>>(ix, counts, first, cnt, off, positions, run_total, out_cap, res_cap, cnt32, off32, seq32, soff32);
}

// one thread: the stage's report for the host (written straight into page-locked host memory) and the batch's running total
__global__ void hits_tail_kernel(const uint64_t *__restrict__ off, uint32_t cnt, const unsigned long long *__restrict__ stats,
                                 const unsigned long long *__restrict__ task_count, unsigned long long task_cap, unsigned long long res_cap,
                                 unsigned long long out_cap, unsigned long long *run_total, uint32_t *off32_end, HitsStageInfo *info) {
    const unsigned long long base = *run_total;
    const unsigned long long total = off ? off[cnt] : 0ull;
    unsigned long long flags = 0;
    if (*task_count > task_cap) flags |= HITS_TASK_OVF;
    if (stats[ST_RES_OVF] || stats[ST_RESULTS] > res_cap) flags |= HITS_RES_OVF;
    if (stats[ST_DUPS]) flags |= HITS_DUPS;
    if (stats[ST_COMPLEX]) flags |= HITS_COMPLEX;
    if (base + total > out_cap) flags |= HITS_OUT_OVF;
    *run_total = base + total;
    if (off32_end) *off32_end = (uint32_t)(base + total);
    info->base = base;
    info->total = total;
    info->flags = flags;
    info->tasks = *task_count;
    __threadfence_system();
}

void launch_hits_tail(const uint64_t *off, uint32_t cnt, const unsigned long long *stats, const unsigned long long *task_count, uint64_t task_cap,
                      uint64_t res_cap, uint64_t out_cap, unsigned long long *run_total, uint32_t *off32_end, HitsStageInfo *info, cudaStream_t s) {
    hits_tail_kernel<<<1, 1, 0, s>>>(off, cnt, stats, task_count, task_cap, res_cap, out_cap, run_total, off32_end, info);
}

}  // namespace e2fm
