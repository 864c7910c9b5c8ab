// search.cuh — device side of the batched search: backward search (count), interval expansion,
// row walks (locate / wildcard / hat verification), and the result pipeline (scan, scatter, segmented
// sort, de-duplication, sequence mapping).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "device_index.cuh"

namespace e2fm {

// A chunk of the pattern batch as the kernels see it.
struct PatternView {
    const uint8_t *bytes;       // device: all pattern bytes of the batch
    const uint64_t *off;        // device: byte offsets [n+1] of the whole batch, or nullptr for fixed-length patterns
    uint32_t fixed_len;         // used when off == nullptr
    uint64_t first;             // batch index of the first pattern of the chunk
    uint32_t count;             // patterns in the chunk
};

// device-side statistics of one batch (u64 each)
enum {
    ST_RANK = 0,     // getOffsetInformations(CountOnly) equivalents   (Bucket.cpp:656)
    ST_LF = 1,       // getOffsetInformations(RetrieveCharacter) equivalents
    ST_MARK = 2,     // getMarkedTextPosition equivalents              (Bucket.cpp:823)
    ST_TASKS = 3,    // (unused: the row-task count of a chunk lives in its own device word, see SearchArgs::task_cursor)
    ST_RESULTS = 4,  // located occurrences appended so far
    ST_BIGSEG = 5,   // result segments longer than the per-thread sort limit
    ST_DUPS = 6,     // duplicate positions removed (EFMIndex.cpp:2296-2299)
    ST_QUEUE = 7,    // work-queue cursor of the running kernel
    ST_ANY_LONG = 8, // set by the backward search when the chunk has a long wildcard interval (firstcheck_kernel has work);
                     // kept next to ST_QUEUE so that one 16-byte memset resets both
    ST_SECT_BS = 13, // distinct 32-byte rank sectors gathered by the backward-search kernel
    ST_GATHER = 9,   // dependent 8-byte gathers issued by the walk kernel
    ST_RES_OVF = 10, // occurrences that did not fit the result store (the host then redoes the batch with exact sizes)
    ST_SCAN_ROWS = 11, // rows streamed by firstcheck_kernel (2 bytes each)
    ST_COMPLEX = 12, // patterns seed_search_kernel handed to the generic path
    ST_N = 14
};

struct SearchArgs {
    DevIndex ix;
    PatternView pv;
    int mode;                        // E2FM_MODE_*
    const uint16_t *codes;           // [batch bytes] k-mer code starting at every pattern byte (encode_patterns_kernel)
    uint2 *iv;                       // [count*k] {sp, len} of the interval each (pattern, shift) ends with (windowed path)
    // fast path (dense tables): long intervals whose first super-character is '?'-prefixed are flagged in iv (bit 31 of the length)
    // and tested row by row by firstcheck_kernel, which appends the matching rows to `tasks` after the expanded ones
    uint32_t split_long;
    uint32_t *firstw;                // [count*k] fixed low-order symbols a row's BWT k-mer must equal ('?'-prefixed first character)
    uint32_t *hatw;                  // [count*k] (f << 28 | fixed high-order symbols) of the '?'-suffixed last character; ~0 = none
    uint2 *tasks;
    unsigned long long *task_cursor; // device-resident number of row tasks of the chunk (expand_kernel sets it, firstcheck adds)
    unsigned long long task_cap;
    unsigned long long *counts;      // [batch n] matches per pattern
    unsigned long long *stats;       // [ST_N]
};

struct WalkArgs {
    DevIndex ix;
    PatternView pv;
    int mode;
    const uint2 *tasks;              // {row, item} with item = localPattern*k + shift
    const uint32_t *firstw, *hatw;   // per-item checks written by the backward search (resolve_kernel only)
    uint32_t n_tasks;                // number of tasks, or the capacity of `tasks` when n_tasks_dev is set
    const unsigned long long *n_tasks_dev;   // fast path: device-resident task count written by expand_kernel (nullptr: use n_tasks)
    unsigned long long *counts;
    uint32_t *res_pat;               // batch pattern index of each located occurrence
    uint64_t *res_pos;               // its global text position
    unsigned long long res_cap;      // capacity of the two arrays above
    unsigned long long *stats;
};

// seed_search.cu: fixed-length batches over 2-bit packed patterns, one thread per pattern
struct SeedArgs {
    DevIndex ix;
    const uint4 *packed;             // [batch n * nw] packed patterns (16 bytes per 64 bases)
    const uint8_t *pflag;            // [batch n] 0 plain, 1 index symbol outside ACGT, 2 foreign symbol; nullptr: all plain
    uint32_t m;                      // pattern length
    uint64_t first;                  // first pattern of the chunk
    uint32_t count;
    int mode;
    unsigned long long *counts;      // [batch n]
    uint32_t *res_pat;
    uint64_t *res_pos;
    unsigned long long res_cap;
    uint32_t *complex_list;          // [batch n] patterns for the generic path, ST_COMPLEX of them
    uint2 *tasks;                    // row tasks {row, chunk-local item} of the chunk: seed_lookup_kernel -> seed_verify_kernel
    unsigned long long task_cap;
    unsigned long long *task_cursor; // device-resident number of row tasks of the chunk
    unsigned long long *stats;
    uint32_t debug;                  // measurement aid (E2FM_SEED_DEBUG): 1 = every lookup reads one of 64 entries (no DRAM gathers; wrong answers)
    uint32_t lookups;                // SeedPlan::lookups when some shift is enumerated (seed_lookup_enum_kernel), else 0
};
// the seed length that serves a batch of m-base patterns (tables_mask: bit j set = a table of seed length j is built), whether some
// shift needs the completions of its last super-character enumerated, and the lookups per pattern; false: the generic path
struct SeedPlan { uint32_t j; bool enumerate; uint32_t lookups; };
bool seed_plan(const DevIndex &ix, uint32_t m, uint32_t tables_mask, SeedPlan &out);
void launch_pack_patterns(const DevIndex &ix, const uint8_t *bytes, uint64_t first, uint32_t count, uint32_t m, uint32_t nw,
                          uint4 *packed, uint8_t *pflag, cudaStream_t s);
void launch_widen_packed(const uint8_t *tight, uint64_t first, uint32_t count, uint32_t m, uint32_t nw, uint4 *packed, cudaStream_t s);
void launch_seed_lookup(const SeedArgs &a, uint32_t nw, int n_sm, cudaStream_t s);
void launch_seed_verify(const SeedArgs &a, uint32_t nw, int n_sm, cudaStream_t s);
void launch_materialize(const uint32_t *sel, uint32_t n_sel, uint32_t m, uint32_t nw, const uint8_t *ascii, const uint4 *packed,
                        uint8_t *out, cudaStream_t s);
void launch_sub_merge(const uint32_t *sel, uint32_t n_sel, const unsigned long long *sub_counts, unsigned long long *counts,
                      uint32_t *res_pat, uint64_t r0, uint64_t r1, cudaStream_t s);
void launch_narrow(const unsigned long long *in, uint32_t *out, uint64_t n, cudaStream_t s);
void launch_debug_seed(const DevIndex &ix, const uint32_t *codes, uint64_t count, uint32_t *out, cudaStream_t s);

// codes[j] for the pattern bytes [j_begin, j_end) of a batch of n_bytes bytes (pv describes all its patterns; j_end is a pattern end)
void launch_encode_patterns(const DevIndex &ix, const PatternView &pv, uint64_t j_begin, uint64_t j_end, uint64_t n_bytes, uint16_t *codes,
                            int n_sm, cudaStream_t s);
void launch_backward_search(const SearchArgs &a, int n_sm, cudaStream_t s);
// fast path: test every row of the wildcard intervals K3 listed (streams bwt16[]), survivors become flagged row tasks
void launch_firstcheck(const SearchArgs &a, int n_sm, cudaStream_t s);
// exclusive scan of iv[i].y (interval lengths) into off[0..n_items] (off[n_items] = total)
void launch_task_scan(const uint2 *iv, uint64_t *off, uint32_t n_items, uint64_t *scratch, cudaStream_t s);
// fast path: scan the interval lengths and write the row tasks in one go (block sums, then scan + expand);
// *n_tasks_out receives the chunk's task count, nothing is written when it exceeds cap
void launch_scan_expand(const uint2 *iv, uint32_t n_items, uint64_t *scratch, uint2 *tasks, uint64_t cap, unsigned long long *n_tasks_out,
                        cudaStream_t s);
// n_tasks_out != nullptr: fast path — window [0, w1) is the capacity, *n_tasks_out receives the chunk's task count, nothing is
// written when it exceeds the capacity
void launch_expand(const uint2 *iv, const uint64_t *off, uint32_t n_items, uint64_t w0, uint64_t w1, uint2 *tasks,
                   unsigned long long *n_tasks_out, cudaStream_t s);
void launch_walk(const WalkArgs &a, int n_sm, cudaStream_t s);
// the same row tasks answered from the dense sa[] / text[] tables (requires a.ix.sa != nullptr)
void launch_resolve(const WalkArgs &a, int n_sm, cudaStream_t s);

// exclusive scan of counts[0..n) into off[0..n] (off[n] = total); scratch holds >= scan_scratch_elems(n) u64
size_t scan_scratch_elems(uint64_t n);
void launch_scan_u64(const unsigned long long *in, uint64_t *off, uint64_t n, uint64_t *scratch, cudaStream_t s);

// result pipeline
void launch_scatter(const uint32_t *res_pat, const uint64_t *res_pos, uint64_t n_res, const uint64_t *off, uint32_t *cursor,
                    uint64_t *positions, cudaStream_t s);
void launch_segment_sort(uint64_t *positions, const uint64_t *off, uint64_t n_pat, uint32_t *big_list, uint32_t *ulen,
                         unsigned long long *stats, cudaStream_t s);
void launch_segment_sort_big(uint64_t *positions, const uint64_t *off, const uint32_t *big_list, uint32_t n_big, uint32_t *ulen,
                             unsigned long long *stats, cudaStream_t s);
// after de-duplication changed some segment lengths: ulen[p] unique entries at the head of each old segment
void launch_repack(const uint64_t *old_pos, const uint64_t *old_off, const uint64_t *new_off, uint64_t n_pat, uint64_t *new_pos,
                   cudaStream_t s);
void launch_ulen_widen(const uint32_t *ulen, unsigned long long *out, uint64_t n, cudaStream_t s);
void launch_map_positions(const DevIndex &ix, const uint64_t *positions, uint64_t n_pos, uint32_t *seq, uint64_t *off, cudaStream_t s);

// hits.cu: result pipeline of one pipeline stage of e2fm_search_hits, without host round trips
struct HitsStageInfo { unsigned long long base, total, flags, tasks; };   // written by the device into page-locked host memory
enum : unsigned long long { HITS_TASK_OVF = 1, HITS_RES_OVF = 2, HITS_DUPS = 4, HITS_COMPLEX = 8, HITS_OUT_OVF = 16 };
void launch_hits_scatter(const uint32_t *res_pat, const uint64_t *res_pos, const unsigned long long *stats, uint64_t res_cap, uint64_t first,
                         const uint64_t *off, uint32_t *cursor, uint64_t *positions, int n_sm, cudaStream_t s);
// exclusive scan of the stage's counts into its CSR (off[cnt] = total), scatter cursors cleared; scratch: cnt / 2048 + 1 u64
void launch_hits_scan(const unsigned long long *counts, uint32_t cnt, uint64_t *off, uint32_t *cursor, unsigned long long *scratch, cudaStream_t s);
struct HitsOut {                                    // where the stage's answers go (batch-wide arrays) and what bounds them
    const unsigned long long *counts;               // [batch n]
    uint64_t first;                                 // first pattern of the stage
    uint32_t cnt;                                   // its patterns
    const uint64_t *off;                            // [cnt + 1] stage-local CSR (nullptr: count mode)
    uint64_t *positions;                            // [off[cnt]] stage-local, scattered
    unsigned long long *run_total;                  // occurrences of the stages before this one
    unsigned long long out_cap, res_cap, task_cap;
    uint32_t *cnt32, *off32, *seq32, *soff32;       // batch-wide 32-bit outputs
    uint32_t *big_list, *off32_end;                 // long segments of the stage; where the batch's total goes (last stage) or nullptr
    unsigned long long *stats;                      // the stage's statistics block
    const unsigned long long *task_count;
    HitsStageInfo *info;                            // page-locked host memory
};
// per-pattern sort + (sequence, offset) mapping + 32-bit offsets and counts, then the long segments and the stage's report
void launch_hits_finish(const DevIndex &ix, const HitsOut &o, int n_sm, cudaStream_t s);

// EFMIndex::extractSnippet (EFMIndex.cpp:2367): the k symbols of super-characters [st_from, st_from + count) from the dense text[] table
void launch_extract(const DevIndex &ix, uint64_t st_from, uint64_t count, uint8_t *out, cudaStream_t s);

// random 8-byte gathers over a buffer of n_words u64: the HBM random-gather roofline measurement
void launch_gather_bench(const uint64_t *buf, uint64_t n_words, uint64_t n_threads, uint32_t per_thread, unsigned long long *sink,
                         int variant, cudaStream_t s);

// test hooks
void launch_debug_lf(const DevIndex &ix, const uint64_t *rows, uint64_t n, uint64_t *out, cudaStream_t s);
void launch_debug_rank(const DevIndex &ix, const uint32_t *codes, const uint64_t *rows, uint64_t n, uint64_t *out, cudaStream_t s);
void launch_debug_bwt(const DevIndex &ix, uint64_t row0, uint64_t n, uint32_t *out, cudaStream_t s);
void launch_debug_marks(const DevIndex &ix, uint64_t row0, uint64_t n, uint64_t *out, cudaStream_t s);

}  // namespace e2fm
