// api.cu — the C ABI of libe2fm_b200.so (include/e2fm_b200.h): index load pipeline, batched search,
// result hand-over and the per-kernel test hooks. Everything below the ABI is CUDA; there is no CPU
// search path in this library.
#include "../../include/e2fm_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "bits.h"
#include "device_index.cuh"
#include "encode.cuh"
#include "host_format.h"
#include "index_build.cuh"
#include "internal.h"
#include "search.cuh"
#include "seed.cuh"

using namespace e2fm;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Stream-ordered device buffer (default memory pool; the pool keeps freed blocks, see e2fm_index_open).
template <class T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    DBuf() = default;
    DBuf(size_t count, cudaStream_t st) { alloc(count, st); }
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    void alloc(size_t count, cudaStream_t st) {
        release();
        s = st;
        n = count;
        if (count) CK(cudaMallocAsync((void **)&p, count * sizeof(T), st));
    }
    void zero() { if (n) CK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    T *take() { T *q = p; p = nullptr; n = 0; return q; }
    ~DBuf() { release(); }
};

// Events are expensive to create on the launch path: the index owns a pool that grows on demand and is reused by every batch.
struct EventPool {
    std::vector<cudaEvent_t> timed, untimed;
    size_t used_timed = 0, used_untimed = 0;
    cudaEvent_t get(bool timing) {
        std::vector<cudaEvent_t> &v = timing ? timed : untimed;
        size_t &u = timing ? used_timed : used_untimed;
        if (u == v.size()) {
            cudaEvent_t e;
            CK(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
            v.push_back(e);
        }
        return v[u++];
    }
    void reset() { used_timed = used_untimed = 0; }
    void destroy() {
        for (auto e : timed) cudaEventDestroy(e);
        for (auto e : untimed) cudaEventDestroy(e);
        timed.clear();
        untimed.clear();
    }
};

struct EventTimer {
    struct Span { cudaEvent_t a, b; int slot; };
    std::vector<Span> spans;
    cudaStream_t s;
    EventPool *pool;
    EventTimer(cudaStream_t st, EventPool *p) : s(st), pool(p) { spans.reserve(64); }
    bool phases = true;      // false: only the whole-batch and H2D spans (slots 0 and 6) are recorded
    int begin(int slot, cudaStream_t on = nullptr) {
        if (!phases && slot != 0 && slot != 6) return -1;
        Span sp;
        sp.slot = slot;
        sp.a = pool->get(true);
        sp.b = pool->get(true);
        CK(cudaEventRecord(sp.a, on ? on : s));
        spans.push_back(sp);
        return (int)spans.size() - 1;
    }
    void end(int h, cudaStream_t on = nullptr) { if (h >= 0) CK(cudaEventRecord(spans[h].b, on ? on : s)); }
    // after the streams have been synchronised
    void collect(float *out, int n_slots) {
        for (auto &sp : spans) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.slot < n_slots) out[sp.slot] += ms;
        }
    }
};

uint32_t floor_pow2(uint32_t v) {
    uint32_t r = 1;
    while (r * 2 <= v) r *= 2;
    return r;
}

}  // namespace

// A result may outlive its index (Python's garbage collector frees them in any order): both hold this token. While the index is
// alive a result's arrays are freed stream-ordered on the index's stream; afterwards synchronously.
struct Lifetime {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool alive = true;
    int refs = 1;
};

struct e2fm_index {
    Lifetime *life = nullptr;
    int device = 0;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // copy engines of the host-buffer pipeline
    cudaStream_t res_stream = nullptr;                         // result kernels of e2fm_search_hits stages (beside the next stage's search)
    int phase_timers = 2;                                         // 0 never, 1 always, 2 device-resident batches only
    bool pipe_taper = false;                                      // host batches: stage sizes halve towards the end of the batch
    uint64_t pipe_patterns = 1u << 18;                          // patterns per pipeline stage when the batch comes from the host
    HostImage img;              // header fields (the per-row vectors stay small: terminators, C, tables)
    DevIndex dix{};
    std::vector<void *> owned;  // device allocations that live as long as the index
    uint64_t device_bytes = 0;
    uint64_t seed_bytes = 0, vrec_bytes = 0;
    const uint4 *seed_tabs[9] = {nullptr};   // seed table per seed length (2 .. dix.seed_chars); dix.seed_tab is the longest
    uint64_t max_seq_len = 0;   // longest span between terminators (a 32-bit offset in the sequence fits when it is below 2^32)
    float load_ms[8] = {0};
    float batch_ms[8] = {0};
    uint64_t batch_ctr[16] = {0};
    double results_per_pattern = 2.0; // learned occurrence density: sizes the fast path's result store
    double tasks_per_pattern = 2.0;   // learned row-task density: sizes the fast path's task buffer
    double seed_tasks_per_pattern = 2.0;   // the same for the seed search (rows of seed intervals per pattern)
    bool use_dense = false;     // answer row tasks from sa[] / text[] instead of walking (E2FM_OPT_DENSE)
    bool use_seed = false;      // fixed-length batches through the seed table (E2FM_OPT_SEED)
    uint64_t chunk_patterns = 4u << 20;
    uint64_t window_tasks = 64u << 20;
    unsigned long long *d_stats = nullptr;
    unsigned long long *h_stats = nullptr;   // pinned
    EventPool events;
    // work buffers of the search path: grow-only, reused by every batch (allocation calls cost more than the kernels of a
    // 1M-pattern batch)
    struct Slot { void *p = nullptr; size_t bytes = 0; } ws[40];
    HitsStageInfo *h_stage_info = nullptr;   // pinned, written by hits_tail_kernel
    size_t h_stage_cap = 0;
    uint64_t hits_stage_patterns = 1u << 20;   // patterns per (full) pipeline stage of e2fm_search_hits
    bool hits_stage_auto = true;               // no E2FM_OPT_HITS_STAGE given: small batches get smaller stages, long ones a short head
    // the kernels of a pipeline stage as ONE graph launch (slot = 3 * locate + packing): re-captured per stage, the instantiated
    // graph is updated in place (same topology, other arguments)
    cudaGraphExec_t hits_exec[12] = {};        // [0..5] search part, [6..11] result part
    bool hits_graph = true;                    // E2FM_HITS_GRAPH=0: plain launches
};

struct e2fm_result {
    e2fm_index *ix = nullptr;   // valid only while life->alive
    Lifetime *life = nullptr;
    uint64_t n = 0, total = 0;
    bool located = false;
    unsigned long long *d_counts = nullptr;
    uint64_t *d_off = nullptr, *d_pos = nullptr, *d_seqoff = nullptr;
    uint32_t *d_seq = nullptr;
};

namespace {

template <class T>
T *dev_keep(e2fm_index *ix, size_t count) {
    T *p = nullptr;
    if (count == 0) count = 1;
    CK(cudaMalloc((void **)&p, count * sizeof(T)));
    ix->owned.push_back(p);
    ix->device_bytes += count * sizeof(T);
    return p;
}

template <class T>
T *dev_keep_copy(e2fm_index *ix, const T *src, size_t count) {
    T *p = dev_keep<T>(ix, count);
    if (count) CK(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, ix->stream));
    return p;
}

// view of a workspace slot with the DBuf interface; contents are preserved across a growth only when asked
enum { WS_BYTES = 0, WS_OFF, WS_CODES, WS_IV, WS_SPARE, WS_TASK_OFF, WS_SCRATCH, WS_TASKS, WS_FIRSTW, WS_CHUNK_CTR, WS_RES_PAT, WS_RES_POS,
       WS_CURSOR, WS_BIG_LIST, WS_HATW, WS_PACKED, WS_PFLAG, WS_CLIST, WS_SUB, WS_NARROW, WS_HCOUNTS, WS_STAGE_STATS, WS_STAGE_OFF, WS_STAGE_POS,
       WS_RUN, WS_OUT_CNT, WS_OUT_OFF, WS_OUT_SEQ, WS_OUT_SOFF, WS_N };
static_assert(WS_N <= 40, "e2fm_index::ws is too small");
template <class T>
struct WsBuf {
    e2fm_index *ix;
    int slot;
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s;
    WsBuf(e2fm_index *ix_, int slot_, cudaStream_t st) : ix(ix_), slot(slot_), s(st) {
        p = (T *)ix->ws[slot].p;
        n = ix->ws[slot].bytes / sizeof(T);
    }
    WsBuf(e2fm_index *ix_, int slot_, size_t count, cudaStream_t st) : WsBuf(ix_, slot_, st) { alloc(count); }
    void alloc(size_t count, size_t keep_elems = 0) {
        auto &sl = ix->ws[slot];
        if (count * sizeof(T) > sl.bytes) {
            const size_t want = count * sizeof(T) + count * sizeof(T) / 4 + 256;
            void *q = nullptr;
            CK(cudaMalloc(&q, want));
            if (keep_elems && sl.p) CK(cudaMemcpyAsync(q, sl.p, keep_elems * sizeof(T), cudaMemcpyDeviceToDevice, s));
            if (sl.p) {
                CK(cudaStreamSynchronize(s));
                CK(cudaFree(sl.p));          // waits for the device: nothing in flight still uses the old block
            }
            sl.p = q;
            sl.bytes = want;
        }
        p = (T *)sl.p;
        n = sl.bytes / sizeof(T);
    }
    void zero(size_t count) { if (count) CK(cudaMemsetAsync(p, 0, count * sizeof(T), s)); }
};

// E2FM_DEBUG_SYNC=1: synchronise after every launch of the search paths and name the kernel that faulted
void dbg_sync(cudaStream_t s, const char *what) {
    static const bool on = getenv("E2FM_DEBUG_SYNC") && atoi(getenv("E2FM_DEBUG_SYNC")) != 0;
    if (!on) return;
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) throw CudaError(std::string("after ") + what + ": " + cudaGetErrorString(e));
}

void read_stats(e2fm_index *ix) {
    CK(cudaMemcpyAsync(ix->h_stats, ix->d_stats, ST_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
}

void check_build_status(e2fm_index *ix, const uint32_t *d_status, const char *stage) {
    uint32_t st = 0;
    CK(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    CK(cudaGetLastError());
    if (st == 0) return;
    std::string m = std::string("index decode failed in ") + stage + ":";
    if (st & BUILD_ERR_BOUNDS) m += " field outside the file;";
    if (st & BUILD_ERR_LENGTH) m += " decoded text disagrees with the header (wrong encryption key or corrupt index);";
    if (st & BUILD_ERR_ALPHABET) m += " alphabet bitset mismatch;";
    if (st & BUILD_ERR_MARKS) m += " marked-row value out of range;";
    throw std::runtime_error(m);
}

// The one-time device pipeline: raw image -> flattened layout (device_index.cuh).
void build_device_index(e2fm_index *ix, const uint8_t *file, size_t size) {
    HostImage &img = ix->img;
    cudaStream_t s = ix->stream;
    ix->events.reset();
    EventTimer tm(s, &ix->events);

    // ---- raw image + start tables to the device (temporary: freed once every bucket is decoded)
    int h = tm.begin(1);
    const size_t padded = (size + 15) / 16 * 16 + 64;
    DBuf<uint8_t> d_file(padded, s);
    CK(cudaMemsetAsync(d_file.p + (padded - 80), 0, 80, s));
    CK(cudaMemcpyAsync(d_file.p, file, size, cudaMemcpyHostToDevice, s));
    DBuf<uint64_t> d_bstart(img.B, s), d_sbstart(img.S, s);
    CK(cudaMemcpyAsync(d_bstart.p, img.b_start.data(), img.B * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_sbstart.p, img.sb_start.data(), img.S * 8, cudaMemcpyHostToDevice, s));
    tm.end(h);

    DBuf<uint32_t> d_status(1, s);
    d_status.zero();
    DBuf<uint16_t> d_sbdecode((size_t)img.S * img.A, s);
    DBuf<uint32_t> d_sbalpha(img.S, s);
    uint64_t *rec = dev_keep<uint64_t>(ix, img.n);
    const uint64_t mrn = img.mrp ? img.mrn : 0;
    uint2 *mark_tab = dev_keep<uint2>(ix, mrn);
    CK(cudaMemsetAsync(mark_tab, 0xFF, std::max<uint64_t>(mrn, 1) * sizeof(uint2), s));

    BuildParams bp{};
    bp.file = d_file.p;
    bp.file_bytes = size;
    bp.b_start = d_bstart.p;
    bp.sb_start = d_sbstart.p;
    bp.n = img.n;
    bp.B = img.B;
    bp.S = img.S;
    bp.b_size = img.b_size;
    bp.per = img.sb_size / img.b_size;
    bp.mrn = mrn;
    bp.A = img.A;
    bp.mrp = img.mrp;
    bp.rate = img.rate;
    bp.val_bits = img.val_bits;
    memcpy(bp.poly_key, img.poly_key, sizeof(bp.poly_key));
    bp.sb_decode = d_sbdecode.p;
    bp.sb_alpha = d_sbalpha.p;
    bp.rec = rec;
    bp.mark_tab = mark_tab;
    bp.status = d_status.p;
    bp.list_elems = (img.A + 1) / 2 * 2;
    bp.mark_words = (uint32_t)((img.b_size + 31) / 32);

    // ---- K2a: super-bucket alphabets
    h = tm.begin(2);
    launch_sb_parse(bp, s);
    tm.end(h);

    // ---- K1 + K2: key stream, then decrypt + RLE0^-1 + MTF^-1 of every bucket, in batches of buckets
    h = tm.begin(3);
    const uint32_t kbits_max = int_log2((uint64_t)img.A + 1);
    const uint32_t blocks_per_bucket = (uint32_t)((img.b_size * kbits_max + 511) / 512 + 2);
    const uint32_t ks_stride = blocks_per_bucket * 64;
    const uint64_t batch = std::max<uint64_t>(1, std::min<uint64_t>(img.B, (256ull << 20) / ks_stride));
    DBuf<uint8_t> d_ks((size_t)batch * ks_stride + 64, s);
    const int wpc = 4;
    const size_t per_warp = (size_t)bp.list_elems * 4 + (size_t)bp.mark_words * 4;
    const size_t smem = per_warp * wpc;
    if (smem > 200 * 1024) throw std::runtime_error("unsupported index: bucket size / alphabet too large for the decode kernel");
    CK(configure_bucket_decode(smem));
    for (uint64_t b0 = 0; b0 < img.B; b0 += batch) {
        const uint64_t nb = std::min(batch, img.B - b0);
        launch_keystream(img.poly_key, b0, nb, blocks_per_bucket, 0, reinterpret_cast<uint4 *>(d_ks.p), s);
        launch_bucket_decode(bp, b0, nb, d_ks.p, ks_stride, wpc, smem, s);
    }
    tm.end(h);
    check_build_status(ix, d_status.p, "bucket decode");
    d_ks.release();
    d_file.release();
    d_sbdecode.release();

    // ---- rank sectors + LF
    const uint32_t blk = std::min<uint32_t>(16384, std::max<uint32_t>(64, floor_pow2(8 * img.A)));
    uint32_t blk_shift = 0;
    while ((1u << blk_shift) < blk) blk_shift++;
    const uint32_t n_blk = (uint32_t)((img.n + blk - 1) / blk);
    uint4 *occ = dev_keep<uint4>(ix, (size_t)n_blk * img.A * 2);
    uint64_t *d_C = dev_keep_copy<uint64_t>(ix, img.C.data(), img.C.size());
    DBuf<unsigned long long> d_cursor(2, s);
    d_cursor.zero();

    RankBuildParams rp{};
    rp.n = img.n;
    rp.A = img.A;
    rp.blk_shift = blk_shift;
    rp.n_blk = n_blk;
    rp.rec = rec;
    rp.occ = occ;
    rp.ovf = nullptr;
    rp.ovf_cursor = d_cursor.p;
    rp.C = d_C;
    rp.mark_tab = mark_tab;
    rp.mrn = mrn;
    rp.status = d_status.p;
    CK(configure_rank_hist((size_t)img.A * 4));
    CK(configure_rank_fill((size_t)2 << blk_shift));
    h = tm.begin(4);
    launch_rank_hist(rp, s);
    tm.end(h);
    unsigned long long need = 0;
    CK(cudaMemcpyAsync(&need, d_cursor.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (need >= (1ull << 32)) throw std::runtime_error("unsupported index: rank overflow pool exceeds 2^32 entries");
    rp.ovf = dev_keep<uint16_t>(ix, need + 16);
    h = tm.begin(5);
    const uint32_t chunk_blocks = 256;
    const uint32_t n_chunks = (n_blk + chunk_blocks - 1) / chunk_blocks;
    DBuf<uint32_t> d_tot((size_t)n_chunks * img.A, s);
    launch_rank_scan(rp, d_tot.p, n_chunks, chunk_blocks, s);
    tm.end(h);
    h = tm.begin(6);
    launch_rank_fill(rp, s);
    tm.end(h);
    check_build_status(ix, d_status.p, "rank build");

    // ---- dense locate tables (6 bytes per row) when HBM allows: EFMIndex::getRowPosition walked once for every row
    uint32_t *sa = nullptr;
    uint16_t *text = nullptr, *bwt16 = nullptr;
    if (mrn > 0 && !(getenv("E2FM_DENSE") && atoi(getenv("E2FM_DENSE")) == 0)) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        if ((double)img.n * 6.0 < 0.5 * (double)free_b) {
            h = tm.begin(6);
            sa = dev_keep<uint32_t>(ix, img.n);
            text = dev_keep<uint16_t>(ix, img.n + 32);   // seed_verify_kernel reads whole 32-byte sectors
            CK(cudaMemsetAsync(text + img.n, 0, 64, s));
            launch_dense_build(rec, mark_tab, mrn, img.rate, img.n, sa, text, d_status.p, s);
            bwt16 = dev_keep<uint16_t>(ix, img.n + 16);   // firstcheck_kernel reads aligned 16-byte groups
            launch_bwt16_build(rec, img.n, bwt16, s);
            tm.end(h);
            check_build_status(ix, d_status.p, "dense table build");
            ix->use_dense = true;
        }
    }

    // ---- small tables
    uint8_t sym_index[256];
    memset(sym_index, 0xFF, sizeof(sym_index));
    for (uint32_t i = 0; i < img.n_sym; i++) sym_index[img.syms[i]] = (uint8_t)i;
    DevIndex &d = ix->dix;
    d.n = img.n;
    d.initial_n = img.initial_n;
    d.mrn = mrn;
    d.k = img.k;
    d.n_sym = img.n_sym;
    d.A = img.A;
    d.rate = img.rate;
    d.blk_shift = blk_shift;
    d.n_blk = n_blk;
    d.sigma_k = img.sigma_k;
    d.n_term = (uint32_t)img.terminators.size();
    d.rec = rec;
    d.occ = occ;
    d.ovf = rp.ovf;
    d.mark_tab = mark_tab;
    d.sa = sa;
    d.text = text;
    d.bwt16 = bwt16;
    d.pair_tab = nullptr;
    d.C = d_C;
    d.compact_of_natural = dev_keep_copy<uint16_t>(ix, img.compact_of_natural.data(), img.compact_of_natural.size());
    d.natural_of_compact = dev_keep_copy<uint32_t>(ix, img.natural_of_compact.data(), img.natural_of_compact.size());
    d.sym_index = dev_keep_copy<uint8_t>(ix, sym_index, 256);
    d.symbols = dev_keep_copy<uint8_t>(ix, img.syms, img.n_sym);
    {
        // low / high symbol parts of every k-mer (EFMIndex::backwardStepIfMatch and findWholePatternMatches compare k-mer strings
        // position by position, EFMIndex.cpp:2011-2041, :1862-1899; here one table lookup and an integer compare)
        const uint32_t km1 = img.k > 1 ? img.k - 1 : 0;
        std::vector<uint32_t> low((size_t)km1 * img.A + 1), high((size_t)km1 * img.A + 1);
        for (uint32_t f = 1; f <= km1; f++) {
            uint64_t pf = 1, pr = 1;
            for (uint32_t i = 0; i < f; i++) pf *= img.n_sym;
            for (uint32_t i = 0; i < img.k - f; i++) pr *= img.n_sym;
            for (uint32_t c = 0; c < img.A; c++) {
                low[(size_t)(f - 1) * img.A + c] = (uint32_t)(img.natural_of_compact[c] % pf);
                high[(size_t)(f - 1) * img.A + c] = (uint32_t)(img.natural_of_compact[c] / pr);
            }
        }
        d.kmer_low = dev_keep_copy<uint32_t>(ix, low.data(), low.size());
        d.kmer_high = dev_keep_copy<uint32_t>(ix, high.data(), high.size());
        CK(cudaStreamSynchronize(s));   // the vectors above go out of scope
    }
    d.terminators = dev_keep_copy<int64_t>(ix, img.terminators.data(), img.terminators.size());
    // start-interval table over pairs of super-characters (needs the finished DevIndex: it ranks through the sectors)
    if ((uint64_t)img.A * img.A <= (4u << 20)) {
        uint2 *pt = dev_keep<uint2>(ix, (size_t)img.A * img.A);
        launch_pair_build(d, pt, s);
        d.pair_tab = pt;
    }
    // ---- 2-bit packed patterns: packed k-mer (base i at bits 2i) -> compact code
    for (int b4 = 0; b4 < 4; b4++) d.sym4[b4] = sym_index[(uint8_t)"ACGT"[b4]];
    d.packed_code = nullptr;
    if (img.k >= 1 && img.k <= 7) {
        const uint32_t np = 1u << (2 * img.k);
        std::vector<uint16_t> pc(np, 0xFFFF);
        for (uint32_t v = 0; v < np; v++) {
            uint64_t nat = 0;
            bool bad = false;
            for (uint32_t i = 0; i < img.k; i++) {
                const uint8_t si = d.sym4[(v >> (2 * i)) & 3u];
                bad |= si == 0xFF;
                nat = nat * img.n_sym + si;
            }
            if (!bad && nat < img.compact_of_natural.size()) pc[v] = img.compact_of_natural[nat];
        }
        d.packed_code = dev_keep_copy<uint16_t>(ix, pc.data(), pc.size());
        CK(cudaStreamSynchronize(s));
    }
    // ---- seed table (seed.cuh): j = the fewest super-characters with A^j >= n/2 (E2FM_SEED_CHARS overrides, 0 = none)
    d.seed_tab = nullptr;
    d.seed_chars = 0;
    d.seed_gph = (img.A + SEED_GROUP - 1) / SEED_GROUP;
    if (sa && text && img.A >= 2) {
        uint32_t j = 2;
        double seeds = (double)img.A * img.A;
        while (j < 4 && seeds < (double)img.n / 2) { seeds *= img.A; j++; }
        if (const char *ov = getenv("E2FM_SEED_CHARS")) j = (uint32_t)atoi(ov);
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        auto table_bytes = [&](uint32_t jj) { double hi = 1; for (uint32_t i = 0; i + 1 < jj; i++) hi *= img.A; return hi * d.seed_gph * 16.0; };
        auto hi_values = [&](uint32_t jj) { double hi = 1; for (uint32_t i = 0; i + 1 < jj; i++) hi *= img.A; return hi; };
        while (j >= 2 && (table_bytes(j) > 0.25 * (double)free_b || hi_values(j) > 2147483648.0)) j--;
        // the main table (seed length j) and the smaller ones below it (a few MB: short patterns are searched with a shorter seed,
        // seed_plan): the same build for every length
        for (uint32_t jj = 2; jj <= j && j <= 8; jj++) {
            h = tm.begin(6);
            const uint64_t n_sectors = (uint64_t)hi_values(jj) * d.seed_gph;
            uint4 *tab = dev_keep<uint4>(ix, n_sectors);
            SeedBuildParams sp{img.n, img.A, jj, d.seed_gph, sa, text, tab, d_status.p};
            launch_seed_build(sp, n_sectors, s);
            tm.end(h);
            uint32_t st = 0;
            CK(cudaMemcpyAsync(&st, d_status.p, 4, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            CK(cudaGetLastError());
            if (st & BUILD_ERR_SEED) {
                // rows not in seed order (never seen on a reference-built index): search without the tables
                CK(cudaFree(tab));
                ix->owned.pop_back();
                ix->device_bytes -= n_sectors * 16;
                for (auto &t : ix->seed_tabs) t = nullptr;
                d.seed_tab = nullptr;
                d.seed_chars = 0;
                ix->seed_bytes = 0;
                break;
            }
            ix->seed_tabs[jj] = tab;
            d.seed_tab = tab;
            d.seed_chars = jj;
            ix->seed_bytes += n_sectors * 16;
        }
    }
    // ---- verification records for the seed search: 32 bytes per row, when HBM allows (E2FM_VREC=0: never)
    d.vrec = nullptr;
    if (d.seed_tab && sa && text && !(getenv("E2FM_VREC") && atoi(getenv("E2FM_VREC")) == 0)) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        if ((double)img.n * 32.0 < 0.3 * (double)free_b) {
            h = tm.begin(6);
            uint4 *vr = dev_keep<uint4>(ix, (size_t)img.n * 2);
            launch_vrec_build(img.n, sa, text, vr, s);
            tm.end(h);
            d.vrec = vr;
            ix->vrec_bytes = (uint64_t)img.n * 32;
        }
    }
    ix->use_seed = d.seed_tab != nullptr && d.packed_code != nullptr && !(getenv("E2FM_SEED") && atoi(getenv("E2FM_SEED")) == 0);
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    tm.collect(ix->load_ms, 7);
    for (int i = 1; i < 7; i++) ix->load_ms[7] += ix->load_ms[i];
}

struct Batch {
    const uint8_t *d_bytes;
    const uint64_t *d_off;
    uint32_t fixed_len;
    uint64_t n;
    uint64_t total_bytes;
    // host-buffer pipeline (all null for device-resident batches)
    const uint8_t *h_bytes = nullptr;     // patterns on the host: copied chunk by chunk on the H2D stream while earlier chunks search
    const uint64_t *h_off = nullptr;      // their offsets (ragged batches)
    uint64_t *h_counts = nullptr;         // optional: counts are copied back chunk by chunk on the D2H stream
    // 2-bit packed input (e2fm_search_batch_packed*): d_bytes / h_bytes hold 16 * ceil(fixed_len / 64) bytes per pattern
    bool packed_in = false;
    uint64_t unit() const { return packed_in ? 16ull * ((fixed_len + 63) / 64) : (uint64_t)fixed_len; }   // input bytes per fixed-length pattern
};

// what the search paths of one batch share: the counts, the list of located occurrences, the bookkeeping
struct RunState {
    e2fm_index *ix;
    cudaStream_t s;
    int mode;
    EventTimer &tm;
    unsigned long long *counts;           // [batch n], device
    WsBuf<uint32_t> res_pat;
    WsBuf<uint64_t> res_pos;
    uint64_t n_results = 0, total_tasks = 0, launches = 0, n_redone = 0;
    unsigned long long redo_rank = 0, redo_sect = 0;
    bool fast_used = false;
    RunState(e2fm_index *ix_, int mode_, EventTimer &tm_, unsigned long long *counts_)
        : ix(ix_), s(ix_->stream), mode(mode_), tm(tm_), counts(counts_), res_pat(ix_, WS_RES_PAT, ix_->stream), res_pos(ix_, WS_RES_POS, ix_->stream) {}
    void grow_results(uint64_t need) {
        if (res_pos.n >= need && res_pat.n >= need) return;
        const uint64_t cap = std::max<uint64_t>(need, res_pos.n + res_pos.n / 2);
        res_pat.alloc(cap, n_results);
        res_pos.alloc(cap, n_results);
    }
    uint64_t res_cap() const { return std::min(res_pat.n, res_pos.n); }
};

// chunk boundaries + the H2D copies of a host batch, queued up front on the copy stream: the search of chunk c waits for its
// own event only, so copies of later chunks overlap the kernels of earlier ones
struct Staging {
    std::vector<uint64_t> bounds;
    std::vector<cudaEvent_t> h2d_ev;
};

Staging stage_batch(e2fm_index *ix, const Batch &b, uint64_t chunk, EventTimer &tm) {
    cudaStream_t s = ix->stream;
    Staging st;
    st.bounds.assign(1, 0);
    // Device batches: equal chunks. Host batches (option TAPER): stages of DECREASING size (each half the previous, down to an
    // eighth of the first): the end-to-end time of the pipeline is the H2D time of the whole batch plus the search of the LAST
    // stage only, so the last stage should be small, while big early stages keep the per-stage overheads few.
    uint64_t sz = chunk;
    const uint64_t floor_sz = std::max<uint64_t>(1, chunk / 8);
    while (st.bounds.back() < b.n) {
        st.bounds.push_back(std::min<uint64_t>(b.n, st.bounds.back() + sz));
        if (b.h_bytes && ix->pipe_taper) sz = std::max<uint64_t>(floor_sz, sz / 2);
    }
    if (b.h_bytes) {
        auto byte_at = [&](uint64_t i) -> uint64_t { return b.h_off ? b.h_off[i] : i * b.unit(); };
        cudaEvent_t ready = ix->events.get(false);
        CK(cudaEventRecord(ready, s));                       // the device buffers were allocated in order on s
        CK(cudaStreamWaitEvent(ix->h2d_stream, ready, 0));
        const int h = tm.begin(0, ix->h2d_stream);
        for (size_t c = 0; c + 1 < st.bounds.size(); c++) {
            const uint64_t lo = byte_at(st.bounds[c]), hi = byte_at(st.bounds[c + 1]);
            if (hi > lo) CK(cudaMemcpyAsync(const_cast<uint8_t *>(b.d_bytes) + lo, b.h_bytes + lo, hi - lo, cudaMemcpyHostToDevice, ix->h2d_stream));
            cudaEvent_t e = ix->events.get(false);
            st.h2d_ev.push_back(e);
            CK(cudaEventRecord(e, ix->h2d_stream));
        }
        tm.end(h, ix->h2d_stream);
    }
    return st;
}

uint64_t pick_chunk(e2fm_index *ix, const Batch &b) {
    const uint32_t k = ix->dix.k;
    uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(ix->chunk_patterns, (1ull << 31) / k));
    if (b.h_bytes) chunk = std::max<uint64_t>(1, std::min<uint64_t>(chunk, ix->pipe_patterns));
    chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(1024, (uint64_t)((double)ix->window_tasks / ix->tasks_per_pattern)));
    return std::min<uint64_t>(chunk, std::max<uint64_t>(b.n, 1));
}

// The generic path (search.cu): any batch — ragged, any k, dense tables or per-query walks. `allow_fast` = false keeps to the
// windowed path (exact sizes, no overflow redo): the sub-batch of a seed search must not disturb what is already in the result list.
void generic_search(RunState &st, const Batch &b, bool allow_fast) {
    e2fm_index *ix = st.ix;
    cudaStream_t s = st.s;
    EventTimer &tm = st.tm;
    const int mode = st.mode;
    const DevIndex &dix = ix->dix;
    const uint32_t k = dix.k;
    uint64_t &launches = st.launches, &n_results = st.n_results, &total_tasks = st.total_tasks;
    unsigned long long *counts = st.counts;

    const uint64_t chunk = pick_chunk(ix, b);
    const Staging sg = stage_batch(ix, b, chunk, tm);
    const std::vector<uint64_t> &bounds = sg.bounds;
    auto byte_at = [&](uint64_t i) -> uint64_t { return b.h_off ? b.h_off[i] : i * (uint64_t)b.fixed_len; };

    const uint64_t items_cap = chunk * k;
    WsBuf<uint2> iv(ix, WS_IV, items_cap, s);
    WsBuf<uint32_t> firstw(ix, WS_FIRSTW, items_cap, s), hatw(ix, WS_HATW, items_cap, s);
    WsBuf<uint64_t> task_off(ix, WS_TASK_OFF, items_cap + 1, s);
    WsBuf<uint64_t> scratch(ix, WS_SCRATCH, scan_scratch_elems(std::max<uint64_t>(items_cap, b.n)), s);
    WsBuf<uint2> tasks(ix, WS_TASKS, s);
    // The single-row step of the backward search (one 8-byte record instead of a rank sector once the interval is one row)
    // saves instructions while everything is L2-resident, but on an HBM-resident index it adds a second large array to the
    // gather working set and loses (measured: 4.6 Mbp 0.164 vs 0.170 ms; 116 Mbp 0.65 vs 0.56 ms; 3.1 Gbp 5.6 vs 4.9 ms per
    // batch): it is used only while rec[] + occ[] fit the L2. E2FM_K3_OPTS (2 = off, 4 = no pair-table start) overrides.
    uint32_t k3_opts = (uint64_t)dix.n * 12 > (96ull << 20) ? 2u : 0u;
    if (const char *ov = getenv("E2FM_K3_OPTS")) k3_opts = (uint32_t)atoi(ov) & 6u;

    // ---- pre-pass: k-mer code at every pattern byte (device batches: all at once; host batches: per chunk, as it arrives)
    WsBuf<uint16_t> codes(ix, WS_CODES, b.total_bytes + 8, s);
    const PatternView all{b.d_bytes, b.d_off, b.fixed_len, 0, (uint32_t)b.n};
    if (!b.h_bytes) {
        const int h = tm.begin(4);
        launch_encode_patterns(dix, all, 0, b.total_bytes, b.total_bytes, codes.p, ix->n_sm, s);
        launches++;
        tm.end(h);
        dbg_sync(s, "encode_patterns_kernel");
    }

    // counts of a finished chunk go back to the host on the D2H stream while the next chunk searches
    auto stream_counts_back = [&](uint64_t first, uint32_t cnt) {
        if (!b.h_counts) return;
        cudaEvent_t e = ix->events.get(false);
        CK(cudaEventRecord(e, s));
        CK(cudaStreamWaitEvent(ix->d2h_stream, e, 0));
        CK(cudaMemcpyAsync(b.h_counts + first, counts + first, (size_t)cnt * 8, cudaMemcpyDeviceToHost, ix->d2h_stream));
    };

    // ---- FAST PATH (dense tables): every chunk is queued without a host round trip. The backward search hands its
    // intervals straight on as row tasks / wildcard intervals, the task count stays on the device. Row tasks go into a
    // buffer of `cap` entries (density learned from earlier batches); a chunk whose tasks do not fit is redone below.
    const bool fast = allow_fast && ix->use_dense && dix.sa && dix.bwt16;
    st.fast_used = fast;
    const uint64_t n_chunks = bounds.size() - 1;
    const uint64_t cap = std::max<uint64_t>(1024, std::min<uint64_t>(ix->window_tasks, (uint64_t)(ix->tasks_per_pattern * (double)chunk) + 1024));
    WsBuf<unsigned long long> chunk_ctr(ix, WS_CHUNK_CTR, std::max<uint64_t>(n_chunks, 1), s);   // row tasks of every chunk
    chunk_ctr.zero(std::max<uint64_t>(n_chunks, 1));

    std::vector<unsigned long long> h_chunk_tasks(n_chunks, 0);
    std::vector<char> redo(n_chunks, fast ? 0 : 1);
    bool full_redo = false;
    auto encode_chunk = [&](uint64_t first, uint32_t cnt, uint64_t chunk_no) {
        if (!b.h_bytes) return;
        CK(cudaStreamWaitEvent(s, sg.h2d_ev[chunk_no], 0));
        const int h = tm.begin(4);
        launch_encode_patterns(dix, all, byte_at(first), byte_at(first + cnt), b.total_bytes, codes.p, ix->n_sm, s);
        launches++;
        tm.end(h);
    };
    if (fast) {
        tasks.alloc(cap);
        // result store: every row task yields at most one occurrence; sized from the learned density, the kernels flag an
        // overflow instead of writing past the end
        if (mode == E2FM_MODE_LOCATE)
            st.grow_results(std::min<uint64_t>(n_chunks * cap, (uint64_t)(ix->results_per_pattern * (double)b.n) + 4096));
        for (uint64_t chunk_no = 0; chunk_no < n_chunks; chunk_no++) {
            const uint64_t first = bounds[chunk_no];
            const uint32_t cnt = (uint32_t)(bounds[chunk_no + 1] - first);
            PatternView pv{b.d_bytes, b.d_off, b.fixed_len, first, cnt};
            encode_chunk(first, cnt, chunk_no);
            CK(cudaMemsetAsync(ix->d_stats + ST_QUEUE, 0, 16, s));   // work-queue cursor + ST_ANY_LONG
            int h = tm.begin(1);
            SearchArgs sa{dix, pv, mode, codes.p, iv.p, 1u | k3_opts, firstw.p, hatw.p, tasks.p, chunk_ctr.p + chunk_no, cap, counts, ix->d_stats};
            launch_backward_search(sa, ix->n_sm, s);
            launches++;
            tm.end(h);
            dbg_sync(s, "backward_search_kernel (fast path)");
            h = tm.begin(2);
            launch_scan_expand(iv.p, cnt * k, scratch.p, tasks.p, cap, chunk_ctr.p + chunk_no, s);   // sets the chunk's task count
            launches += 2;
            tm.end(h);
            dbg_sync(s, "scan_expand");
            h = tm.begin(7);
            launch_firstcheck(sa, ix->n_sm, s);                                                      // appends the wildcard survivors
            launches++;
            tm.end(h);
            dbg_sync(s, "firstcheck_kernel");
            h = tm.begin(3);
            WalkArgs wa{dix, pv, mode, tasks.p, firstw.p, hatw.p, (uint32_t)cap, chunk_ctr.p + chunk_no, counts, st.res_pat.p, st.res_pos.p, st.res_cap(), ix->d_stats};
            launch_resolve(wa, ix->n_sm, s);
            launches++;
            tm.end(h);
            dbg_sync(s, "resolve_kernel (fast path)");
            stream_counts_back(first, cnt);
        }
        if (n_chunks) CK(cudaMemcpyAsync(h_chunk_tasks.data(), chunk_ctr.p, n_chunks * 8, cudaMemcpyDeviceToHost, s));
        read_stats(ix);
        n_results = ix->h_stats[ST_RESULTS];
        for (uint64_t c = 0; c < n_chunks; c++) {
            total_tasks += h_chunk_tasks[c];
            redo[c] = h_chunk_tasks[c] > cap;
        }
        if (ix->h_stats[ST_RES_OVF]) {
            // the result store was too small: start the whole batch over through the exact-size path
            CK(cudaMemsetAsync(counts, 0, std::max<uint64_t>(b.n, 1) * 8, s));
            CK(cudaMemsetAsync(ix->d_stats, 0, ST_N * sizeof(unsigned long long), s));
            n_results = 0;
            full_redo = true;
            std::fill(redo.begin(), redo.end(), 1);
        }
    }

    // ---- WINDOWED PATH: per-query walks (dense tables off), and the chunks of the fast path that overflowed. One interval
    // per (pattern, shift), scanned and expanded into row-task windows sized from the host.
    for (uint64_t c = 0; c < n_chunks; c++) {
        if (!redo[c]) continue;
        if (fast) st.n_redone++;
        const uint64_t first = bounds[c];
        const uint32_t cnt = (uint32_t)(bounds[c + 1] - first);
        PatternView pv{b.d_bytes, b.d_off, b.fixed_len, first, cnt};
        unsigned long long rank0 = 0, sect0 = 0;
        if (fast) {
            // the fast path already added this chunk's directly counted intervals: start its counts over
            CK(cudaMemsetAsync(counts + first, 0, (size_t)cnt * 8, s));
            read_stats(ix);
            rank0 = ix->h_stats[ST_RANK];
            sect0 = ix->h_stats[ST_SECT_BS];
        } else encode_chunk(first, cnt, c);
        CK(cudaMemsetAsync(ix->d_stats + ST_QUEUE, 0, 8, s));
        int h = tm.begin(1);
        SearchArgs sa{dix, pv, mode, codes.p, iv.p, k3_opts, firstw.p, hatw.p, nullptr, nullptr, 0, counts, ix->d_stats};
        launch_backward_search(sa, ix->n_sm, s);
        launches++;
        tm.end(h);
        h = tm.begin(2);
        launch_task_scan(iv.p, task_off.p, cnt * k, scratch.p, s);
        launches += 3;
        tm.end(h);
        unsigned long long n_tasks = 0;
        CK(cudaMemcpyAsync(&n_tasks, task_off.p + (size_t)cnt * k, 8, cudaMemcpyDeviceToHost, s));
        read_stats(ix);
        if (fast && !full_redo) {   // the backward search ran twice for this chunk: keep its operations counted once
            st.redo_rank += ix->h_stats[ST_RANK] - rank0;
            st.redo_sect += ix->h_stats[ST_SECT_BS] - sect0;
        }
        if (!fast) total_tasks += n_tasks;
        const uint64_t win = std::min<uint64_t>(n_tasks, ix->window_tasks);
        if (tasks.n < win) tasks.alloc(win);
        for (uint64_t w0 = 0; w0 < n_tasks; w0 += win) {
            const uint64_t w1 = std::min<uint64_t>(n_tasks, w0 + win);
            h = tm.begin(2);
            launch_expand(iv.p, task_off.p, cnt * k, w0, w1, tasks.p, nullptr, s);
            launches++;
            tm.end(h);
            if (mode == E2FM_MODE_LOCATE) st.grow_results(n_results + (w1 - w0));   // every row task yields at most one occurrence
            CK(cudaMemsetAsync(ix->d_stats + ST_QUEUE, 0, 8, s));
            h = tm.begin(3);
            WalkArgs wa{dix, pv, mode, tasks.p, firstw.p, hatw.p, (uint32_t)(w1 - w0), nullptr, counts, st.res_pat.p, st.res_pos.p, st.res_cap(), ix->d_stats};
            if (ix->use_dense && dix.sa) launch_resolve(wa, ix->n_sm, s);
            else launch_walk(wa, ix->n_sm, s);
            launches++;
            tm.end(h);
            read_stats(ix);
            n_results = ix->h_stats[ST_RESULTS];
        }
        stream_counts_back(first, cnt);
    }
    if (b.h_counts) CK(cudaStreamSynchronize(ix->d2h_stream));
}

// the seed length for a batch of m-base patterns, from the tables this index holds
bool plan_seed(const e2fm_index *ix, uint32_t m, SeedPlan &plan) {
    uint32_t mask = 0;
    for (uint32_t j = 2; j <= 8; j++) if (ix->seed_tabs[j]) mask |= 1u << j;
    return seed_plan(ix->dix, m, mask, plan);
}

// The seed path (seed_search.cu): fixed-length batches whose every shift holds a whole seed; ASCII input is packed on the
// device first. Patterns it cannot finish (several rows left at the end of a shift, a saturated seed counter, an index symbol
// outside ACGT) are searched by the generic path as a sub-batch.
void seed_search(RunState &st, const Batch &b, const SeedPlan &plan) {
    e2fm_index *ix = st.ix;
    cudaStream_t s = st.s;
    EventTimer &tm = st.tm;
    DevIndex dix = ix->dix;                       // the kernels see the table of this batch's seed length
    dix.seed_tab = ix->seed_tabs[plan.j];
    dix.seed_chars = plan.j;
    const uint32_t m = b.fixed_len, nw = (m + 63) / 64;
    const bool locate = st.mode == E2FM_MODE_LOCATE;

    uint64_t chunk = std::max<uint64_t>(1, ix->chunk_patterns);
    if (b.h_bytes) chunk = std::max<uint64_t>(1, std::min<uint64_t>(chunk, ix->pipe_patterns));
    chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(b.n, 1));
    const Staging sg = stage_batch(ix, b, chunk, tm);
    const uint64_t n_chunks = sg.bounds.size() - 1;

    const uint4 *packed = reinterpret_cast<const uint4 *>(b.d_bytes);
    const uint8_t *pflag = nullptr;
    WsBuf<uint4> packed_ws(ix, WS_PACKED, s);
    WsBuf<uint8_t> pflag_ws(ix, WS_PFLAG, s);
    if (!b.packed_in) {
        packed_ws.alloc(b.n * nw + 1);
        pflag_ws.alloc(b.n + 1);
        packed = packed_ws.p;
        pflag = pflag_ws.p;
    }
    WsBuf<uint32_t> clist(ix, WS_CLIST, std::max<uint64_t>(b.n, 1), s);
    if (locate) st.grow_results((uint64_t)(ix->results_per_pattern * (double)b.n) + 4096);
    // row tasks of a chunk (seed_lookup_kernel -> seed_verify_kernel): capacity from the density of earlier batches; the kernels
    // never write past it and the host searches again with room when a chunk needed more
    WsBuf<uint2> tasks(ix, WS_TASKS, s);
    uint64_t cap = (uint64_t)(ix->seed_tasks_per_pattern * (double)chunk) + 4096;
    WsBuf<unsigned long long> chunk_ctr(ix, WS_CHUNK_CTR, std::max<uint64_t>(n_chunks, 1), s);
    std::vector<unsigned long long> h_chunk_tasks(n_chunks, 0);

    uint64_t n_complex = 0, seed_tasks = 0;
    for (int attempt = 0;; attempt++) {
        tasks.alloc(cap);
        chunk_ctr.zero(std::max<uint64_t>(n_chunks, 1));
        for (uint64_t c = 0; c < n_chunks; c++) {
            const uint64_t first = sg.bounds[c];
            const uint32_t cnt = (uint32_t)(sg.bounds[c + 1] - first);
            if (b.h_bytes && attempt == 0) CK(cudaStreamWaitEvent(s, sg.h2d_ev[c], 0));
            if (!b.packed_in && attempt == 0) {
                const int h = tm.begin(4);
                launch_pack_patterns(dix, b.d_bytes, first, cnt, m, nw, packed_ws.p, pflag_ws.p, s);
                st.launches++;
                tm.end(h);
            }
            SeedArgs sa{dix, packed, pflag, m, first, cnt, st.mode, st.counts, st.res_pat.p, st.res_pos.p, st.res_cap(), clist.p,
                        tasks.p, cap, chunk_ctr.p + c, ix->d_stats, getenv("E2FM_SEED_DEBUG") ? (uint32_t)atoi(getenv("E2FM_SEED_DEBUG")) : 0u,
                        plan.enumerate ? plan.lookups : 0u};
            int h = tm.begin(1);
            launch_seed_lookup(sa, nw, ix->n_sm, s);
            tm.end(h);
            dbg_sync(s, plan.enumerate ? "seed_lookup_enum_kernel" : "seed_lookup_kernel");
            h = tm.begin(3);
            launch_seed_verify(sa, nw, ix->n_sm, s);
            tm.end(h);
            dbg_sync(s, "seed_verify kernel");
            st.launches += 2;
        }
        if (n_chunks) CK(cudaMemcpyAsync(h_chunk_tasks.data(), chunk_ctr.p, n_chunks * 8, cudaMemcpyDeviceToHost, s));
        read_stats(ix);
        uint64_t most = 0;
        seed_tasks = 0;
        for (uint64_t c = 0; c < n_chunks; c++) { most = std::max<uint64_t>(most, h_chunk_tasks[c]); seed_tasks += h_chunk_tasks[c]; }
        const bool task_ovf = most > cap, res_ovf = ix->h_stats[ST_RES_OVF] != 0;
        if (!task_ovf && !res_ovf) break;
        // a store sized from the density of earlier batches was too small: enlarge it and search again
        if (task_ovf) cap = most + most / 8 + 4096;
        if (locate) {
            const uint64_t need = std::max<uint64_t>(ix->h_stats[ST_RESULTS] + ix->h_stats[ST_RES_OVF], task_ovf ? seed_tasks : 0);
            st.n_results = 0;
            st.grow_results(need + need / 8 + 4096);
        }
        CK(cudaMemsetAsync(ix->d_stats, 0, ST_N * sizeof(unsigned long long), s));
        st.n_redone++;
    }
    st.total_tasks += seed_tasks;
    if (b.n) ix->seed_tasks_per_pattern = std::min(1.0e6, std::max(1.5, 1.3 * (double)seed_tasks / (double)b.n));
    st.n_results = ix->h_stats[ST_RESULTS];
    n_complex = ix->h_stats[ST_COMPLEX];
    if (n_complex) {
        // ---- what the seed search left: an ASCII sub-batch with local pattern numbers through the generic path
        WsBuf<uint8_t> sub(ix, WS_SUB, n_complex * m + 16, s);
        launch_materialize(clist.p, (uint32_t)n_complex, m, nw, b.packed_in ? nullptr : b.d_bytes, packed, sub.p, s);
        st.launches++;
        DBuf<unsigned long long> sub_counts(n_complex, s);
        sub_counts.zero();
        Batch sb{sub.p, nullptr, m, n_complex, n_complex * m};
        RunState sub_st(ix, st.mode, tm, sub_counts.p);
        sub_st.n_results = st.n_results;
        generic_search(sub_st, sb, false);
        launch_sub_merge(clist.p, (uint32_t)n_complex, sub_counts.p, st.counts, sub_st.res_pat.p, st.n_results, sub_st.n_results, s);
        st.launches += sub_st.launches + 1;
        st.total_tasks += sub_st.total_tasks;
        st.n_results = sub_st.n_results;
        st.res_pat.p = sub_st.res_pat.p; st.res_pat.n = sub_st.res_pat.n;     // the list may have grown
        st.res_pos.p = sub_st.res_pos.p; st.res_pos.n = sub_st.res_pos.n;
        CK(cudaStreamSynchronize(s));                                          // sub_counts goes out of scope
    }
    ix->batch_ctr[10] = n_complex;
    if (b.h_counts && b.n) {
        CK(cudaMemcpyAsync(b.h_counts, st.counts, b.n * 8, cudaMemcpyDeviceToHost, s));
    }
}

void run_batch(e2fm_index *ix, const Batch &b, int mode, e2fm_result *res, EventTimer &tm) {
    cudaStream_t s = ix->stream;
    const DevIndex &dix = ix->dix;
    res->n = b.n;
    res->located = mode == E2FM_MODE_LOCATE;
    DBuf<unsigned long long> counts(std::max<uint64_t>(b.n, 1), s);
    CK(cudaMemsetAsync(ix->d_stats, 0, ST_N * sizeof(unsigned long long), s));
    // every CUDA call on the launch path costs the host a few microseconds, and a pipeline stage of a host batch holds only
    // ~0.1 ms of kernels: per-phase event pairs are recorded for device batches only (E2FM_OPT_PHASE_TIMERS overrides)
    tm.phases = ix->phase_timers == 1 || (ix->phase_timers == 2 && !b.h_bytes);
    ix->batch_ctr[10] = ix->batch_ctr[12] = ix->batch_ctr[13] = 0;

    RunState st(ix, mode, tm, counts.p);
    SeedPlan plan{0, false, 0};
    const bool seeded = ix->use_seed && ix->use_dense && !b.d_off && !b.h_off && b.n > 0 && plan_seed(ix, b.fixed_len, plan);
    if (b.packed_in && !seeded && b.n > 0) throw std::runtime_error("packed patterns need the seed search: dense tables and seed table built, k <= 7, "
                                                         "pattern length <= 128 with a whole seed in every shift (use the ASCII entry points)");
    if (seeded) seed_search(st, b, plan);  // writes every count itself
    else {
        counts.zero();
        generic_search(st, b, true);
    }
    uint64_t n_results = st.n_results;
    uint64_t &launches = st.launches;
    WsBuf<uint64_t> scratch(ix, WS_SCRATCH, scan_scratch_elems(std::max<uint64_t>(b.n, 1)), s);
    // learn the densities for the next batch's capacities (x1.5 head room)
    if (b.n) {
        if (st.fast_used) ix->tasks_per_pattern = std::min(1.0e6, std::max(2.0, 1.5 * (double)st.total_tasks / (double)b.n));
        if (mode == E2FM_MODE_LOCATE) ix->results_per_pattern = std::min(1.0e6, std::max(seeded ? 1.25 : 2.0, (seeded ? 1.2 : 1.5) * (double)n_results / (double)b.n));
    }

    // ---- result pipeline (locate): CSR offsets, scatter, segmented sort + de-duplication, sequence mapping
    uint64_t dups = 0;
    if (mode == E2FM_MODE_LOCATE) {
        int h = tm.begin(5);
        DBuf<uint64_t> off(b.n + 1, s);
        launch_scan_u64(counts.p, off.p, b.n, scratch.p, s);
        launches += 3;
        DBuf<uint64_t> pos(std::max<uint64_t>(n_results, 1), s);
        WsBuf<uint32_t> cursor(ix, WS_CURSOR, std::max<uint64_t>(b.n, 1), s);
        cursor.zero(std::max<uint64_t>(b.n, 1));
        launch_scatter(st.res_pat.p, st.res_pos.p, n_results, off.p, cursor.p, pos.p, s);
        launches++;
        WsBuf<uint32_t> big_list(ix, WS_BIG_LIST, n_results / 24 + 1, s);
        uint32_t *ulen = cursor.p;   // the cursors are dead after the scatter
        launch_segment_sort(pos.p, off.p, b.n, big_list.p, ulen, ix->d_stats, s);
        launches++;
        tm.end(h);
        read_stats(ix);
        const uint64_t n_big = ix->h_stats[ST_BIGSEG];
        if (n_big) {
            h = tm.begin(5);
            launch_segment_sort_big(pos.p, off.p, big_list.p, (uint32_t)n_big, ulen, ix->d_stats, s);
            launches++;
            tm.end(h);
            read_stats(ix);
        }
        dups = ix->h_stats[ST_DUPS];
        h = tm.begin(5);
        if (dups) {
            // EFMIndex::locateOccurrences erased adjacent duplicates (:2296-2299): rebuild the CSR around the unique heads
            DBuf<unsigned long long> wide(b.n, s);
            launch_ulen_widen(ulen, wide.p, b.n, s);
            DBuf<uint64_t> noff(b.n + 1, s);
            launch_scan_u64(wide.p, noff.p, b.n, scratch.p, s);
            DBuf<uint64_t> npos(std::max<uint64_t>(n_results - dups, 1), s);
            launch_repack(pos.p, off.p, noff.p, b.n, npos.p, s);
            launches += 5;
            std::swap(pos.p, npos.p); std::swap(pos.n, npos.n);
            std::swap(off.p, noff.p); std::swap(off.n, noff.n);
            n_results -= dups;
        }
        DBuf<uint32_t> seq(std::max<uint64_t>(n_results, 1), s);
        DBuf<uint64_t> seqoff(std::max<uint64_t>(n_results, 1), s);
        launch_map_positions(dix, pos.p, n_results, seq.p, seqoff.p, s);
        launches++;
        tm.end(h);
        res->total = n_results;
        res->d_off = off.take();
        res->d_pos = pos.take();
        res->d_seq = seq.take();
        res->d_seqoff = seqoff.take();
    }
    res->d_counts = counts.take();
    // Every path above ends with a statistics read (which synchronises the stream); the search kernels ran before it and the
    // result pipeline, which runs after it in locate mode, only touches the DUPS / BIGSEG words it reads itself.
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    ix->batch_ctr[0] = ix->h_stats[ST_RANK] - st.redo_rank;
    ix->batch_ctr[1] = ix->h_stats[ST_LF];
    ix->batch_ctr[2] = ix->h_stats[ST_MARK];
    ix->batch_ctr[3] = st.total_tasks;
    ix->batch_ctr[4] = ix->h_stats[ST_SECT_BS] - st.redo_sect;
    ix->batch_ctr[5] = launches;
    ix->batch_ctr[6] = n_results;
    ix->batch_ctr[7] = ix->h_stats[ST_GATHER];
    ix->batch_ctr[8] = ix->h_stats[ST_SCAN_ROWS];
    ix->batch_ctr[9] = st.n_redone;
    ix->batch_ctr[11] = seeded ? 1 : 0;
}

// The stages of an e2fm_search_hits batch of n patterns: bounds[c] .. bounds[c + 1] is stage c; *stage_out = patterns of a full stage.
// Full stages of `opt_stage` patterns; automatic mode (no E2FM_OPT_HITS_STAGE given): a batch shorter than four stages is cut into
// quarters (>= 2^17 patterns: a stage costs ~60 us of launches and kernel tails whatever its size), a long one starts with 1/4, 1/4,
// 1/2 of a stage so that the search begins early. The last stage is cut into 1/2, 1/4, 1/4: what the call takes is the H2D time of
// the whole batch plus search and D2H of the LAST stage (measured: profiles/exp/e2e_stages.py).
std::vector<uint64_t> hits_stage_plan(uint64_t n, uint64_t opt_stage, bool auto_stage, uint64_t *stage_out) {
    uint64_t stage = std::max<uint64_t>(1, std::min<uint64_t>(opt_stage, n));
    if (auto_stage && n < 4 * stage) stage = std::min<uint64_t>(n, std::max<uint64_t>(131072, (n / 4 + 65535) / 65536 * 65536));
    std::vector<uint64_t> bounds(1, 0);
    if (auto_stage && n >= 8 * stage) {
        bounds.push_back(stage / 4);
        bounds.push_back(stage / 2);
        bounds.push_back(stage);
    }
    while (bounds.back() < n) {
        const uint64_t left = n - bounds.back();
        if (left > stage) { bounds.push_back(bounds.back() + stage); continue; }
        if (left >= (auto_stage ? 262144u : 65536u)) {
            bounds.push_back(bounds.back() + left / 2);
            bounds.push_back(bounds.back() + left / 4);
        }
        bounds.push_back(n);
    }
    if (stage_out) *stage_out = stage;
    return bounds;
}

// e2fm_search_hits, the pipelined path: host patterns in, host results out, stage by stage. Stage c is copied in on the H2D stream
// while stage c-1 runs on the compute stream (seed lookup, verification, the stage's result pipeline of hits.cu) and the results of
// stage c-2 leave on the D2H stream. Nothing on the compute stream needs the host: counts stay on the device, the device carries the
// running occurrence total, and each stage reports {first occurrence, occurrences, trouble} into page-locked host memory, which is
// all the host needs to queue the stage's D2H copies. Returns false when a store sized from earlier batches was too small or a
// pattern needs the generic path (an index symbol outside ACGT): the caller then runs the plain path, which sizes exactly.
bool hits_pipeline(e2fm_index *ix, const uint8_t *h_in, uint64_t n, uint32_t m, int packing, int mode, e2fm_hits *hits, const SeedPlan &plan) {
    cudaStream_t s = ix->stream;
    DevIndex dix = ix->dix;                       // the kernels see the table of this batch's seed length
    dix.seed_tab = ix->seed_tabs[plan.j];
    dix.seed_chars = plan.j;
    const bool locate = mode == E2FM_MODE_LOCATE;
    const uint32_t nw = (m + 63) / 64;
    Batch b{nullptr, nullptr, m, n, 0};
    const bool packed_in = packing == E2FM_PACKED_WORDS;     // already in the 16-byte form the kernels load
    const bool tight_in = packing == E2FM_PACKED_TIGHT;      // ceil(m / 4) bytes each: widened on the device, stage by stage
    b.packed_in = packed_in;
    const uint64_t unit = tight_in ? e2fm_packed_tight_stride(m) : b.unit();   // input bytes per pattern
    b.total_bytes = n * unit;
    b.h_bytes = h_in;
    memset(ix->batch_ms, 0, sizeof(ix->batch_ms));
    memset(ix->batch_ctr, 0, sizeof(ix->batch_ctr));
    ix->events.reset();
    EventTimer tm(s, &ix->events);
    tm.phases = false;
    const int whole = tm.begin(6);
    uint64_t stage = 1;
    Staging sg;
    sg.bounds = hits_stage_plan(n, ix->hits_stage_patterns, ix->hits_stage_auto, &stage);
    const uint64_t S_max = sg.bounds.size() + 8;              // sizes the per-stage arrays

    WsBuf<uint8_t> d_in(ix, WS_BYTES, b.total_bytes + 16, s);
    b.d_bytes = d_in.p;
    const uint4 *packed = reinterpret_cast<const uint4 *>(d_in.p);
    const uint8_t *pflag = nullptr;
    WsBuf<uint4> packed_ws(ix, WS_PACKED, s);
    WsBuf<uint8_t> pflag_ws(ix, WS_PFLAG, s);
    if (!packed_in) {
        packed_ws.alloc(n * nw + 1);
        packed = packed_ws.p;
        if (!tight_in) {
            pflag_ws.alloc(n + 1);
            pflag = pflag_ws.p;
        }
    }
    WsBuf<unsigned long long> counts(ix, WS_HCOUNTS, n, s);
    WsBuf<uint32_t> clist(ix, WS_CLIST, stage, s);
    const uint64_t task_cap = (uint64_t)(ix->seed_tasks_per_pattern * (double)stage) + 4096;
    WsBuf<uint2> tasks(ix, WS_TASKS, task_cap, s);
    WsBuf<unsigned long long> task_ctr(ix, WS_CHUNK_CTR, S_max, s);
    task_ctr.zero(S_max);
    WsBuf<unsigned long long> stage_stats(ix, WS_STAGE_STATS, S_max * ST_N, s);
    stage_stats.zero(S_max * ST_N);
    WsBuf<unsigned long long> run(ix, WS_RUN, 2, s);
    run.zero(2);
    WsBuf<uint32_t> cnt32(ix, WS_OUT_CNT, n, s);
    const uint64_t res_cap = locate ? (uint64_t)(ix->results_per_pattern * (double)stage) + 4096 : 0;
    const uint64_t out_cap = locate ? std::min<uint64_t>(hits->capacity, (uint64_t)(ix->results_per_pattern * (double)n) + 65536) : 0;
    WsBuf<uint32_t> res_pat(ix, WS_RES_PAT, s), cursor(ix, WS_CURSOR, s), big_list(ix, WS_BIG_LIST, s), off32(ix, WS_OUT_OFF, s),
        seq32(ix, WS_OUT_SEQ, s), soff32(ix, WS_OUT_SOFF, s);
    WsBuf<uint64_t> res_pos(ix, WS_RES_POS, s), st_off(ix, WS_STAGE_OFF, s), st_pos(ix, WS_STAGE_POS, s), scratch(ix, WS_SCRATCH, s);
    if (locate) {
        res_pat.alloc(2 * res_cap);                          // two halves: stage c uses half c & 1
        res_pos.alloc(2 * res_cap);
        st_off.alloc(stage + 1);
        st_pos.alloc(res_cap);
        cursor.alloc(stage);
        big_list.alloc(res_cap / 24 + 1);
        scratch.alloc(scan_scratch_elems(stage));
        off32.alloc(n + 1);
        seq32.alloc(out_cap + 1);
        soff32.alloc(out_cap + 1);
    }
    if (ix->h_stage_cap < S_max) {
        if (ix->h_stage_info) CK(cudaFreeHost(ix->h_stage_info));
        ix->h_stage_info = nullptr;
        ix->h_stage_cap = 0;
        CK(cudaHostAlloc((void **)&ix->h_stage_info, (S_max + 8) * sizeof(HitsStageInfo), cudaHostAllocDefault));
        ix->h_stage_cap = S_max + 8;
    }
    HitsStageInfo *info = ix->h_stage_info;
    memset(info, 0, S_max * sizeof(HitsStageInfo));

    const uint64_t S = sg.bounds.size() - 1;
    {
        cudaEvent_t ready = ix->events.get(false);
        CK(cudaEventRecord(ready, s));                       // the device buffers were allocated in order on s
        CK(cudaStreamWaitEvent(ix->h2d_stream, ready, 0));
        const int hh = tm.begin(0, ix->h2d_stream);
        for (uint64_t c = 0; c < S; c++) {                   // every H2D copy queued up front, one event per stage
            const uint64_t lo = sg.bounds[c] * unit, hi = sg.bounds[c + 1] * unit;
            CK(cudaMemcpyAsync(d_in.p + lo, h_in + lo, hi - lo, cudaMemcpyHostToDevice, ix->h2d_stream));
            cudaEvent_t e = ix->events.get(false);
            sg.h2d_ev.push_back(e);
            CK(cudaEventRecord(e, ix->h2d_stream));
        }
        tm.end(hh, ix->h2d_stream);
    }
    std::vector<cudaEvent_t> done(S);
    // E2FM_HITS_TRACE=1: per-stage timeline (CUDA events on the three streams) on stderr
    // (2: also an event after every kernel group of a stage; E2FM_HITS_SERIAL=1: the search starts when the whole batch has arrived)
    static const int trace_level = getenv("E2FM_HITS_TRACE") ? atoi(getenv("E2FM_HITS_TRACE")) : 0;
    static const bool serial = getenv("E2FM_HITS_SERIAL") && atoi(getenv("E2FM_HITS_SERIAL")) != 0;
    const bool trace = trace_level != 0, trace_k = trace_level >= 2;
    std::vector<cudaEvent_t> tr_begin(trace ? S : 0), tr_search(trace ? S : 0), tr_out(trace ? S : 0);
    std::vector<cudaEvent_t> tr_k(trace_k ? S * 4 : 0);
    // two compute streams: the search kernels of stage c + 1 (s) run beside the result kernels of stage c (rs); the lists of located
    // occurrences between them are double-buffered (E2FM_HITS_OVERLAP=0: one stream)
    static const bool overlap_env = !(getenv("E2FM_HITS_OVERLAP") && atoi(getenv("E2FM_HITS_OVERLAP")) == 0);
    cudaStream_t rs = overlap_env ? ix->res_stream : s;
    auto mark = [&](uint64_t c, int which, cudaStream_t on) {
        if (!trace_k) return;
        tr_k[c * 4 + which] = ix->events.get(true);
        CK(cudaEventRecord(tr_k[c * 4 + which], on));
    };
    if (serial) CK(cudaStreamWaitEvent(s, sg.h2d_ev[S - 1], 0));
    static const bool graph_env = !(getenv("E2FM_HITS_GRAPH") && atoi(getenv("E2FM_HITS_GRAPH")) == 0);
    const bool use_graph = graph_env && ix->hits_graph && !trace_k;
    cudaEvent_t tr_zero = nullptr;
    if (trace) { tr_zero = ix->events.get(true); CK(cudaEventRecord(tr_zero, s)); }
    if (rs != s) {                                             // the buffers were allocated (and zeroed) in order on s
        cudaEvent_t ready = ix->events.get(false);
        CK(cudaEventRecord(ready, s));
        CK(cudaStreamWaitEvent(rs, ready, 0));
    }
    uint64_t launches = 0;
    // The kernels of a part go down as ONE graph launch: while the batch is still arriving the PCIe link is full of H2D traffic, and
    // every command the GPU front end fetches from the host waits behind it (profiles/exp/logs/hits_trace2_*: the gap between two
    // launches grows from 3 to 25 us). The graph is re-captured per stage and the instantiated one updated in place.
    auto graph_part = [&](cudaStream_t on, int slot, auto &&body) {
        if (!use_graph) { body(); return; }
        CK(cudaStreamBeginCapture(on, cudaStreamCaptureModeRelaxed));
        try {
            body();
        } catch (...) {
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(on, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        cudaGraph_t g = nullptr;
        CK(cudaStreamEndCapture(on, &g));
        cudaGraphExec_t &ge = ix->hits_exec[slot];
        if (ge) {
            cudaGraphExecUpdateResultInfo upd;
            if (cudaGraphExecUpdate(ge, g, &upd) != cudaSuccess) {          // another topology: instantiate anew
                cudaGetLastError();
                cudaGraphExecDestroy(ge);
                ge = nullptr;
            }
        }
        if (!ge) {
            const cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
            if (e != cudaSuccess) { ge = nullptr; cudaGraphDestroy(g); CK(e); }
        }
        cudaGraphDestroy(g);
        CK(cudaGraphLaunch(ge, on));
    };
    std::vector<cudaEvent_t> searched(S);
    auto enqueue = [&](uint64_t c) {
        const uint64_t first = sg.bounds[c];
        const uint32_t cnt = (uint32_t)(sg.bounds[c + 1] - first);
        unsigned long long *st = stage_stats.p + c * ST_N;
        uint32_t *rpat = res_pat.p ? res_pat.p + (c & 1) * res_cap : nullptr;
        uint64_t *rpos = res_pos.p ? res_pos.p + (c & 1) * res_cap : nullptr;
        const int slot = (locate ? 3 : 0) + packing;
        // search part: [widen | pack], seed lookup, verification
        CK(cudaStreamWaitEvent(s, sg.h2d_ev[c], 0));
        if (c >= 2 && rs != s) CK(cudaStreamWaitEvent(s, done[c - 2], 0));     // the results of stage c - 2 have left this half of the lists
        if (trace) { tr_begin[c] = ix->events.get(true); CK(cudaEventRecord(tr_begin[c], s)); }
        graph_part(s, slot, [&]() {
            if (tight_in) { launch_widen_packed(b.d_bytes, first, cnt, m, nw, packed_ws.p, s); launches++; }
            else if (!packed_in) { launch_pack_patterns(dix, b.d_bytes, first, cnt, m, nw, packed_ws.p, pflag_ws.p, s); launches++; }
            SeedArgs sa{dix, packed, pflag, m, first, cnt, mode, counts.p, rpat, rpos, res_cap, clist.p, tasks.p, task_cap, task_ctr.p + c, st, 0u, plan.enumerate ? plan.lookups : 0u};
            mark(c, 0, s);
            launch_seed_lookup(sa, nw, ix->n_sm, s);
            mark(c, 1, s);
            launch_seed_verify(sa, nw, ix->n_sm, s);
            launches += 2;
        });
        searched[c] = ix->events.get(trace);
        CK(cudaEventRecord(searched[c], s));
        if (trace) tr_search[c] = searched[c];
        // result part: the stage's CSR, scatter, sort + (sequence, offset), report
        if (rs != s) CK(cudaStreamWaitEvent(rs, searched[c], 0));
        graph_part(rs, 6 + slot, [&]() {
            if (locate) {
                launch_hits_scan(counts.p + first, cnt, st_off.p, cursor.p, reinterpret_cast<unsigned long long *>(scratch.p), rs);
                mark(c, 2, rs);
                launch_hits_scatter(rpat, rpos, st, res_cap, first, st_off.p, cursor.p, st_pos.p, ix->n_sm, rs);
                launches += 3;
            }
            mark(c, 3, rs);
            HitsOut ho{counts.p, first, cnt, locate ? st_off.p : nullptr, st_pos.p, run.p, out_cap, res_cap, task_cap, cnt32.p, off32.p, seq32.p, soff32.p,
                       big_list.p, locate && c + 1 == S ? off32.p + n : nullptr, st, task_ctr.p + c, info + c};
            launch_hits_finish(dix, ho, ix->n_sm, rs);
            launches += 2;
        });
        done[c] = ix->events.get(trace);
        CK(cudaEventRecord(done[c], rs));
    };
    bool trouble = false;
    uint64_t total = 0;
    auto drain = [&](uint64_t c) {
        const uint64_t first = sg.bounds[c];
        const uint64_t cnt = sg.bounds[c + 1] - first;
        if (locate) {
            CK(cudaEventSynchronize(done[c]));                 // the stage's report is in host memory
            if (info[c].flags) { trouble = true; return; }
        } else CK(cudaStreamWaitEvent(ix->d2h_stream, done[c], 0));
        if (hits->counts) CK(cudaMemcpyAsync(hits->counts + first, cnt32.p + first, cnt * 4, cudaMemcpyDeviceToHost, ix->d2h_stream));
        if (locate) {
            const uint64_t base = info[c].base, tot = info[c].total;
            if (hits->occ_offsets)
                CK(cudaMemcpyAsync(hits->occ_offsets + first, off32.p + first, (cnt + (c + 1 == S ? 1 : 0)) * 4, cudaMemcpyDeviceToHost, ix->d2h_stream));
            if (tot && hits->seq_index) CK(cudaMemcpyAsync(hits->seq_index + base, seq32.p + base, tot * 4, cudaMemcpyDeviceToHost, ix->d2h_stream));
            if (tot && hits->seq_offset) CK(cudaMemcpyAsync(hits->seq_offset + base, soff32.p + base, tot * 4, cudaMemcpyDeviceToHost, ix->d2h_stream));
            total = base + tot;
        }
        if (trace) { tr_out[c] = ix->events.get(true); CK(cudaEventRecord(tr_out[c], ix->d2h_stream)); }
    };
    const uint64_t ahead = 2;                                  // stages queued on the compute stream beyond the one being drained
    uint64_t enqueued = 0;
    for (uint64_t c = 0; c < S && !trouble; c++) {
        enqueue(c);
        enqueued = c + 1;
        if (c >= ahead) drain(c - ahead);
    }
    for (uint64_t c = S > ahead ? S - ahead : 0; c < S && !trouble; c++) drain(c);
    if (rs != s && enqueued) CK(cudaStreamWaitEvent(s, done[enqueued - 1], 0));
    tm.end(whole);
    CK(cudaStreamSynchronize(s));
    CK(cudaStreamSynchronize(rs));
    CK(cudaStreamSynchronize(ix->d2h_stream));
    CK(cudaStreamSynchronize(ix->h2d_stream));
    CK(cudaGetLastError());
    if (trace && !trouble) {
        for (uint64_t c = 0; c < S; c++) {
            float t0 = 0, t1 = 0, t2 = 0, t3 = 0;
            cudaEventElapsedTime(&t0, tr_zero, tr_begin[c]);
            cudaEventElapsedTime(&t1, tr_zero, tr_search[c]);
            cudaEventElapsedTime(&t2, tr_zero, done[c]);
            cudaEventElapsedTime(&t3, tr_zero, tr_out[c]);
            fprintf(stderr, "[hits] stage %2llu: %8llu patterns  search starts %7.3f  seed kernels done %7.3f  results done %7.3f  D2H done %7.3f ms\n",
                    (unsigned long long)c, (unsigned long long)(sg.bounds[c + 1] - sg.bounds[c]), t0, t1, t2, t3);
            if (trace_k && locate) {
                float k0 = 0, k1 = 0, k2 = 0, k3 = 0, k4 = 0, k5 = 0;
                cudaEventElapsedTime(&k0, tr_begin[c], tr_k[c * 4 + 0]);        // widen / pack
                cudaEventElapsedTime(&k1, tr_k[c * 4 + 0], tr_k[c * 4 + 1]);    // lookup
                cudaEventElapsedTime(&k2, tr_k[c * 4 + 1], tr_search[c]);       // verify
                cudaEventElapsedTime(&k3, tr_search[c], tr_k[c * 4 + 2]);       // scan (2 kernels)
                cudaEventElapsedTime(&k4, tr_k[c * 4 + 2], tr_k[c * 4 + 3]);    // scatter
                cudaEventElapsedTime(&k5, tr_k[c * 4 + 3], done[c]);            // sort + finish, tail + report
                fprintf(stderr, "[hits]           us: widen %5.1f  lookup %6.1f  verify %6.1f  scan %5.1f  scatter %5.1f  sortfin+tail %5.1f\n",
                        1e3f * k0, 1e3f * k1, 1e3f * k2, 1e3f * k3, 1e3f * k4, 1e3f * k5);
            }
        }
    }
    // what the stages report: trouble flags (count mode reads them only now), densities for the next batch's stores
    uint64_t tasks_sum = 0, worst_tasks = 0;
    for (uint64_t c = 0; c < S; c++) {
        if (info[c].flags) trouble = true;
        ix->batch_ctr[13] |= info[c].flags;
        tasks_sum += info[c].tasks;
        worst_tasks = std::max<uint64_t>(worst_tasks, info[c].tasks);
    }
    if (worst_tasks) ix->seed_tasks_per_pattern = std::min(1.0e6, std::max(1.5, 1.3 * (double)worst_tasks / (double)stage));
    if (trouble) return false;
    if (locate && n) ix->results_per_pattern = std::min(1.0e6, std::max(1.25, 1.2 * (double)total / (double)n));
    hits->total = total;
    std::vector<unsigned long long> hs(S * ST_N);
    CK(cudaMemcpyAsync(hs.data(), stage_stats.p, S * ST_N * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint64_t c = 0; c < S; c++) {
        ix->batch_ctr[4] += hs[c * ST_N + ST_SECT_BS];
        ix->batch_ctr[7] += hs[c * ST_N + ST_GATHER];
    }
    ix->batch_ctr[3] = tasks_sum;
    ix->batch_ctr[5] = launches;
    ix->batch_ctr[6] = total;
    ix->batch_ctr[11] = 1;
    ix->batch_ctr[12] = S;
    tm.collect(ix->batch_ms, 8);
    return true;
}

// a failed batch may have copies queued that still read the caller's host buffers: nothing is in flight when the call returns
void quiesce(e2fm_index *ix) {
    if (ix->h2d_stream) cudaStreamSynchronize(ix->h2d_stream);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    if (ix->res_stream) cudaStreamSynchronize(ix->res_stream);
    if (ix->d2h_stream) cudaStreamSynchronize(ix->d2h_stream);
    cudaGetLastError();
}

int search_common(e2fm_index *ix, const uint8_t *bytes, const uint64_t *offsets, uint32_t fixed_len, uint64_t n, uint64_t total_bytes,
                  bool on_device, int mode, e2fm_result **out, uint64_t *counts_out = nullptr, bool packed_in = false) {
    if (!ix || !out || (mode != E2FM_MODE_COUNT && mode != E2FM_MODE_LOCATE)) return fail(E2FM_ERR_ARG, "bad argument");
    if (n && !bytes && total_bytes) return fail(E2FM_ERR_ARG, "pattern bytes missing");
    if (n >= (1ull << 32)) return fail(E2FM_ERR_ARG, "more than 2^32-1 patterns in one batch");
    // EFMIndex.cpp:2305-2306 throws this for locate; count needs the marked rows as well (hat verification walks to
    // them, :1862-1899 — the reference crashes there), so both modes are refused.
    if (ix->img.mrp == 0)
        return fail(E2FM_ERR_UNSUPPORTED, "The index doesn't contain marked rows: locate operation not supported");
    *out = nullptr;
    e2fm_result *res = new (std::nothrow) e2fm_result();
    if (!res) return fail(E2FM_ERR_OOM, "out of host memory");
    res->ix = ix;
    res->life = ix->life;
    ix->life->refs++;
    try {
        CK(cudaSetDevice(ix->device));
        cudaStream_t s = ix->stream;
        memset(ix->batch_ms, 0, sizeof(ix->batch_ms));
        ix->events.reset();
        EventTimer tm(s, &ix->events);
        const int whole = tm.begin(6);
        WsBuf<uint8_t> d_bytes(ix, WS_BYTES, s);
        WsBuf<uint64_t> d_off(ix, WS_OFF, s);
        Batch b{bytes, offsets, fixed_len, n, total_bytes};
        b.packed_in = packed_in;
        if (!on_device) {
            // the pattern bytes travel chunk by chunk inside run_batch (H2D stream); only the offsets go ahead of them
            d_bytes.alloc(total_bytes + 16);
            b.d_bytes = d_bytes.p;
            b.h_bytes = bytes;
            b.h_off = offsets;
            b.h_counts = counts_out;
            if (offsets) {
                d_off.alloc(n + 1);
                CK(cudaMemcpyAsync(d_off.p, offsets, (n + 1) * 8, cudaMemcpyHostToDevice, s));
                b.d_off = d_off.p;
            }
            if (n == 0 || total_bytes == 0) b.h_bytes = nullptr;
        }
        run_batch(ix, b, mode, res, tm);
        if (on_device && counts_out && n) CK(cudaMemcpyAsync(counts_out, res->d_counts, n * 8, cudaMemcpyDeviceToHost, s));
        tm.end(whole);
        CK(cudaStreamSynchronize(s));
        tm.collect(ix->batch_ms, 8);
    } catch (const CudaError &e) {
        quiesce(ix);
        e2fm_result_free(res);
        return fail(E2FM_ERR_CUDA, e.what());
    } catch (const std::bad_alloc &) {
        quiesce(ix);
        e2fm_result_free(res);
        return fail(E2FM_ERR_OOM, "out of host memory");
    } catch (const std::exception &e) {
        quiesce(ix);
        e2fm_result_free(res);
        return fail(E2FM_ERR_FORMAT, e.what());
    }
    *out = res;
    return E2FM_OK;
}

static int with_device(int device, cudaStream_t *s) {
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(E2FM_ERR_CUDA, "no CUDA device: e2fm_b200 has no CPU search path");
    }
    if (device < 0 || device >= n_dev) return fail(E2FM_ERR_ARG, "no such CUDA device");
    if (cudaSetDevice(device) != cudaSuccess) return fail(E2FM_ERR_CUDA, "cudaSetDevice failed");
    *s = nullptr;
    return E2FM_OK;
}

static void key_words(const uint8_t *k32, uint32_t w[8]) {
    for (int i = 0; i < 8; i++)
        w[i] = (uint32_t)k32[4 * i] | ((uint32_t)k32[4 * i + 1] << 8) | ((uint32_t)k32[4 * i + 2] << 16) | ((uint32_t)k32[4 * i + 3] << 24);
}

template <class F>
static int debug_call(const e2fm_index *ix, F f) {
    if (!ix) return fail(E2FM_ERR_ARG, "bad argument");
    try {
        CK(cudaSetDevice(ix->device));
        f(ix->stream);
        CK(cudaStreamSynchronize(ix->stream));
        CK(cudaGetLastError());
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

}  // namespace

extern "C" {

const char *e2fm_last_error(void) { return g_err.c_str(); }
const char *e2fm_version(void) { return "e2fm_b200 0.1 (sm_100a)"; }

int e2fm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void *e2fm_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        g_err = "cudaHostAlloc failed";
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void e2fm_host_free(void *p) { if (p) cudaFreeHost(p); }

// host-only: what e2fm_index_open does before it touches the device (no CUDA call in here)
int e2fm_parse_header(const uint8_t *file, size_t size, const uint8_t key64[64], e2fm_info *info, uint64_t *c_out, uint64_t c_cap,
                      int64_t *term_out, uint64_t term_cap) {
    if (!file || !key64 || !info) return fail(E2FM_ERR_ARG, "bad argument");
    try {
        HostImage g;
        parse_index_header(file, size, key64, g);
        memset(info, 0, sizeof(*info));
        info->text_length = g.n;
        info->bucket_size = g.b_size;
        info->superbucket_size = g.sb_size;
        info->n_buckets = g.B;
        info->n_superbuckets = g.S;
        info->n_sequences = g.terminators.size();
        info->initial_n = g.initial_n;
        info->k = g.k;
        info->n_symbols = g.n_sym;
        info->compact_alphabet = g.A;
        info->marked_rows_percentage = g.mrp;
        info->marking_rate = g.rate;
        info->quirks = g.quirks;
        for (uint64_t i = 0; i < g.C.size() && i < c_cap && c_out; i++) c_out[i] = g.C[i];
        for (uint64_t i = 0; i < g.terminators.size() && i < term_cap && term_out; i++) term_out[i] = g.terminators[i];
    } catch (const std::bad_alloc &) {
        return fail(E2FM_ERR_OOM, "out of host memory");
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_FORMAT, e.what());
    }
    return E2FM_OK;
}

int e2fm_index_open(const uint8_t *file, size_t size, const uint8_t key64[64], int device, e2fm_index **out) {
    if (!file || !key64 || !out) return fail(E2FM_ERR_ARG, "bad argument");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(E2FM_ERR_CUDA, "no CUDA device: e2fm_b200 has no CPU search path");
    }
    if (device < 0 || device >= n_dev) return fail(E2FM_ERR_ARG, "no such CUDA device");
    e2fm_index *ix = new (std::nothrow) e2fm_index();
    if (!ix) return fail(E2FM_ERR_OOM, "out of host memory");
    ix->device = device;
    int rc = E2FM_OK;
    try {
        const double t0 = now_ms();
        parse_index_header(file, size, key64, ix->img);
        ix->load_ms[0] = (float)(now_ms() - t0);
        {
            const HostImage &g = ix->img;
            const uint64_t text_end = g.n * g.k + g.initial_n;
            uint64_t prev = 0, mx = g.terminators.empty() ? text_end : 0;
            for (int64_t t : g.terminators) { mx = std::max<uint64_t>(mx, (uint64_t)t - prev); prev = (uint64_t)t; }
            ix->max_seq_len = std::max<uint64_t>(mx, text_end > prev ? text_end - prev : 0);
        }
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        ix->n_sm = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
        ix->life = new Lifetime();
        ix->life->device = device;
        ix->life->stream = ix->stream;
        CK(cudaStreamCreateWithFlags(&ix->h2d_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ix->d2h_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ix->res_stream, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = ~0ull;   // keep freed blocks in the pool: batches reuse their work buffers
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        CK(cudaMalloc((void **)&ix->d_stats, ST_N * sizeof(unsigned long long)));
        CK(cudaHostAlloc((void **)&ix->h_stats, ST_N * sizeof(unsigned long long), cudaHostAllocDefault));
        build_device_index(ix, file, size);
    } catch (const CudaError &e) {
        rc = fail(E2FM_ERR_CUDA, e.what());
    } catch (const std::bad_alloc &) {
        rc = fail(E2FM_ERR_OOM, "out of host memory");
    } catch (const std::exception &e) {
        rc = fail(E2FM_ERR_FORMAT, e.what());
    }
    if (rc != E2FM_OK) {
        e2fm_index_close(ix);
        return rc;
    }
    *out = ix;
    return E2FM_OK;
}

void e2fm_index_close(e2fm_index *ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    for (void *p : ix->owned) cudaFree(p);
    for (auto &sl : ix->ws) if (sl.p) cudaFree(sl.p);
    ix->events.destroy();
    if (ix->d_stats) cudaFree(ix->d_stats);
    if (ix->h_stats) cudaFreeHost(ix->h_stats);
    if (ix->h_stage_info) cudaFreeHost(ix->h_stage_info);
    for (auto &ge : ix->hits_exec) if (ge) cudaGraphExecDestroy(ge);
    if (ix->h2d_stream) cudaStreamDestroy(ix->h2d_stream);
    if (ix->d2h_stream) cudaStreamDestroy(ix->d2h_stream);
    if (ix->res_stream) cudaStreamDestroy(ix->res_stream);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    if (ix->life) {
        ix->life->alive = false;                 // results that are still around free their arrays synchronously from now on
        if (--ix->life->refs == 0) delete ix->life;
    }
    cudaGetLastError();
    delete ix;
}

int e2fm_index_info(const e2fm_index *ix, e2fm_info *o) {
    if (!ix || !o) return fail(E2FM_ERR_ARG, "bad argument");
    memset(o, 0, sizeof(*o));
    const HostImage &g = ix->img;
    o->text_length = g.n;
    o->bucket_size = g.b_size;
    o->superbucket_size = g.sb_size;
    o->n_buckets = g.B;
    o->n_superbuckets = g.S;
    o->n_sequences = g.terminators.size();
    o->initial_n = g.initial_n;
    o->device_bytes = ix->device_bytes;
    o->k = g.k;
    o->n_symbols = g.n_sym;
    o->compact_alphabet = g.A;
    o->marked_rows_percentage = g.mrp;
    o->marking_rate = g.rate;
    o->rank_block = 1u << ix->dix.blk_shift;
    o->quirks = g.quirks;
    o->dense = ix->dix.sa ? (ix->use_dense ? 1u : 2u) : 0u;
    o->seed_chars = ix->dix.seed_tab ? ix->dix.seed_chars : 0u;
    o->seed_table_mb = (uint32_t)(ix->seed_bytes >> 20);
    o->vrec_mb = (uint32_t)(ix->vrec_bytes >> 20);
    memcpy(o->load_ms, ix->load_ms, sizeof(o->load_ms));
    return E2FM_OK;
}

int e2fm_index_terminators(const e2fm_index *ix, int64_t *out, uint64_t cap) {
    if (!ix || (!out && cap)) return fail(E2FM_ERR_ARG, "bad argument");
    const uint64_t n = std::min<uint64_t>(cap, ix->img.terminators.size());
    for (uint64_t i = 0; i < n; i++) out[i] = ix->img.terminators[i];
    return E2FM_OK;
}

uint64_t e2fm_index_fasta_headers(const e2fm_index *ix, char *out, uint64_t cap) {
    if (!ix) return 0;
    const std::string &h = ix->img.fasta_headers;
    if (out && cap) memcpy(out, h.data(), (size_t)std::min<uint64_t>(cap, h.size()));
    return h.size();
}

void *e2fm_index_stream(const e2fm_index *ix) { return ix ? (void *)ix->stream : nullptr; }

int e2fm_search_batch(e2fm_index *ix, const uint8_t *bytes, const uint64_t *offsets, uint64_t n, int mode, e2fm_result **out) {
    if (n && !offsets) return fail(E2FM_ERR_ARG, "offsets missing");
    return search_common(ix, bytes, offsets, 0, n, n ? offsets[n] : 0, false, mode, out);
}
int e2fm_search_batch_fixed(e2fm_index *ix, const uint8_t *bytes, uint64_t n, uint32_t length, int mode, e2fm_result **out) {
    return search_common(ix, bytes, nullptr, length, n, n * length, false, mode, out);
}
int e2fm_search_batch_stream(e2fm_index *ix, const uint8_t *bytes, const uint64_t *offsets, uint64_t n, uint32_t length, int mode,
                             uint64_t *counts_out, e2fm_result **out) {
    if (offsets) return search_common(ix, bytes, offsets, 0, n, n ? offsets[n] : 0, false, mode, out, counts_out);
    return search_common(ix, bytes, nullptr, length, n, n * length, false, mode, out, counts_out);
}
int e2fm_search_batch_device(e2fm_index *ix, const uint8_t *d_bytes, const uint64_t *d_offsets, uint64_t n, uint64_t total_bytes, int mode,
                             e2fm_result **out) {
    if (n && !d_offsets) return fail(E2FM_ERR_ARG, "offsets missing");
    return search_common(ix, d_bytes, d_offsets, 0, n, total_bytes, true, mode, out);
}
int e2fm_search_batch_device_fixed(e2fm_index *ix, const uint8_t *d_bytes, uint64_t n, uint32_t length, int mode, e2fm_result **out) {
    return search_common(ix, d_bytes, nullptr, length, n, n * length, true, mode, out);
}

uint32_t e2fm_packed_stride(uint32_t length) { return 16u * ((length + 63u) / 64u); }
uint32_t e2fm_packed_tight_stride(uint32_t length) { return (length + 3u) / 4u; }

int e2fm_pack_patterns_tight(const uint8_t *ascii, uint64_t n, uint32_t length, uint8_t *packed_out, uint64_t *n_unpackable) {
    if ((n && (!ascii || !packed_out)) || length == 0) return fail(E2FM_ERR_ARG, "bad argument");
    static const struct Lut { uint8_t v[256]; Lut() { memset(v, 4, 256); v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3; } } lut;
    const uint32_t tb = e2fm_packed_tight_stride(length);
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *src = ascii + i * length;
        uint8_t *dst = packed_out + i * tb;
        uint32_t worst = 0;
        for (uint32_t b = 0; b < tb; b++) {
            uint32_t cur = 0;
            const uint32_t end = std::min<uint32_t>(length, 4 * b + 4);
            for (uint32_t j = 4 * b, sh = 0; j < end; j++, sh += 2) {
                const uint32_t c = lut.v[src[j]];
                worst |= c;
                cur |= (c & 3u) << sh;
            }
            dst[b] = (uint8_t)cur;
        }
        if (worst & 4u) bad++;
    }
    if (n_unpackable) *n_unpackable = bad;
    return E2FM_OK;
}

int e2fm_pack_patterns(const uint8_t *ascii, uint64_t n, uint32_t length, uint8_t *packed_out, uint64_t *n_unpackable) {
    if ((n && (!ascii || !packed_out)) || length == 0) return fail(E2FM_ERR_ARG, "bad argument");
    const uint32_t stride = e2fm_packed_stride(length);
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *src = ascii + i * length;
        uint64_t *dst = reinterpret_cast<uint64_t *>(packed_out + i * stride);
        bool ok = true;
        for (uint32_t w = 0; w < stride / 8; w++) {
            uint64_t cur = 0;
            const uint32_t end = std::min<uint32_t>(length, (w + 1) * 32);
            for (uint32_t j = w * 32, sh = 0; j < end; j++, sh += 2) {
                uint32_t c;
                switch (src[j]) {
                    case 'A': c = 0; break;
                    case 'C': c = 1; break;
                    case 'G': c = 2; break;
                    case 'T': c = 3; break;
                    default: c = 0; ok = false;
                }
                cur |= (uint64_t)c << sh;
            }
            dst[w] = cur;
        }
        if (!ok) bad++;
    }
    if (n_unpackable) *n_unpackable = bad;
    return E2FM_OK;
}

int e2fm_search_batch_packed(e2fm_index *ix, const uint8_t *packed, uint64_t n, uint32_t length, int mode, uint64_t *counts_out,
                             e2fm_result **out) {
    return search_common(ix, packed, nullptr, length, n, n * e2fm_packed_stride(length), false, mode, out, counts_out, true);
}
int e2fm_search_batch_packed_device(e2fm_index *ix, const uint8_t *d_packed, uint64_t n, uint32_t length, int mode, e2fm_result **out) {
    if (d_packed && (reinterpret_cast<uintptr_t>(d_packed) & 15u)) return fail(E2FM_ERR_ARG, "packed patterns must be 16-byte aligned");
    return search_common(ix, d_packed, nullptr, length, n, n * e2fm_packed_stride(length), true, mode, out, nullptr, true);
}

int e2fm_search_hits(e2fm_index *ix, const uint8_t *patterns, uint64_t n, uint32_t length, int packed, int mode, e2fm_hits *hits) {
    if (!ix || !hits || length == 0 || (mode != E2FM_MODE_COUNT && mode != E2FM_MODE_LOCATE)) return fail(E2FM_ERR_ARG, "bad argument");
    if (n && !patterns) return fail(E2FM_ERR_ARG, "pattern bytes missing");
    if (n >= (1ull << 32)) return fail(E2FM_ERR_ARG, "more than 2^32-1 patterns in one batch");
    if (packed < 0 || packed > E2FM_PACKED_TIGHT) return fail(E2FM_ERR_ARG, "packed must be E2FM_ASCII, E2FM_PACKED_WORDS or E2FM_PACKED_TIGHT");
    const bool locate = mode == E2FM_MODE_LOCATE;
    if (locate && (!hits->occ_offsets || ((!hits->seq_index || !hits->seq_offset) && hits->capacity)))
        return fail(E2FM_ERR_ARG, "locate needs occ_offsets, seq_index and seq_offset");
    if (ix->img.mrp == 0) return fail(E2FM_ERR_UNSUPPORTED, "The index doesn't contain marked rows: locate operation not supported");
    if (locate && ix->max_seq_len >= (1ull << 32)) return fail(E2FM_ERR_ARG, "a sequence is longer than 2^32-1: use e2fm_result_fetch");
    hits->total = 0;
    bool piped = false;
    SeedPlan plan{0, false, 0};
    if (n && ix->use_seed && ix->use_dense && plan_seed(ix, length, plan)) {
        try {
            CK(cudaSetDevice(ix->device));
            piped = hits_pipeline(ix, patterns, n, length, packed, mode, hits, plan);
        } catch (const CudaError &e) {
            quiesce(ix);
            return fail(E2FM_ERR_CUDA, e.what());
        } catch (const std::bad_alloc &) {
            quiesce(ix);
            return fail(E2FM_ERR_OOM, "out of host memory");
        } catch (const std::exception &e) {
            quiesce(ix);
            return fail(E2FM_ERR_FORMAT, e.what());
        }
    }
    if (piped) return E2FM_OK;
    // the plain path: whole batch, exact sizes, then the compact fetch
    e2fm_result *r = nullptr;
    const uint64_t unit = packed ? e2fm_packed_stride(length) : length;
    const uint64_t gave_up = ix->batch_ctr[13];
    std::vector<uint8_t> widened;                            // tight input: the plain path takes the 16-byte form
    if (packed == E2FM_PACKED_TIGHT) {
        try { widened.assign(n * unit, 0); } catch (const std::bad_alloc &) { return fail(E2FM_ERR_OOM, "out of host memory"); }
        const uint32_t tb = e2fm_packed_tight_stride(length);
        const uint32_t tail_bits = 2 * (length & 3u);        // bits of the last byte that belong to the pattern (0: all 8)
        for (uint64_t i = 0; i < n; i++) {
            memcpy(widened.data() + i * unit, patterns + i * tb, tb);
            if (tail_bits) widened[i * unit + tb - 1] &= (uint8_t)((1u << tail_bits) - 1u);
        }
        patterns = widened.data();
    }
    int rc = search_common(ix, patterns, nullptr, length, n, n * unit, false, mode, &r, nullptr, packed != 0);
    if (rc != E2FM_OK) return rc;
    ix->batch_ctr[13] = gave_up;
    hits->total = r->total;
    if (locate && r->total > hits->capacity) {
        e2fm_result_free(r);
        return fail(E2FM_ERR_ARG, "e2fm_hits.capacity is too small for the occurrences of this batch (see e2fm_hits.total)");
    }
    rc = e2fm_result_fetch_compact(r, hits->counts, locate ? hits->occ_offsets : nullptr, locate ? hits->seq_index : nullptr,
                                   locate ? hits->seq_offset : nullptr);
    e2fm_result_free(r);
    return rc;
}

int e2fm_extract(e2fm_index *ix, uint64_t from, uint64_t to, char *out, uint64_t cap, uint64_t *out_len) {
    if (!ix || !out_len || (!out && cap)) return fail(E2FM_ERR_ARG, "bad argument");
    *out_len = 0;
    if (ix->img.mrp == 0)   // EFMIndex.cpp:2326-2327
        return fail(E2FM_ERR_UNSUPPORTED, "The index doesn't contain marked rows: the extraction operation is not supported");
    if (!ix->dix.text) return fail(E2FM_ERR_UNSUPPORTED, "extract needs the dense text table, which was not built for this index");
    const HostImage &g = ix->img;
    // EFMIndex::extractSnippet(from, to) (EFMIndex.cpp:2325-2356), unsigned arithmetic and all
    from -= g.initial_n;
    to -= g.initial_n;
    if (from > to) return E2FM_OK;
    const uint64_t k = g.k;
    uint64_t st_from = from / k, st_to = to / k;
    const uint64_t off_from = from % k, off_to = to % k;
    if (st_to > g.n - 2) st_to = g.n - 2;
    if (st_from > st_to || (st_from == st_to && off_from > off_to)) return E2FM_OK;
    const uint64_t count = st_to - st_from + 1;
    try {
        CK(cudaSetDevice(ix->device));
        cudaStream_t s = ix->stream;
        DBuf<uint8_t> d(count * k, s);
        launch_extract(ix->dix, st_from, count, d.p, s);
        std::vector<char> h(count * k);
        CK(cudaMemcpyAsync(h.data(), d.p, count * k, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        // EFMIndex.cpp:2441-2452. One super-character: substr(stFromOffset, stToOffset + 1) — the second argument is a LENGTH
        // there, so the reference returns up to stToOffset+1 symbols starting at stFromOffset; mirrored.
        uint64_t a, b;   // [a, b) of the decoded symbols
        if (count == 1) { a = off_from; b = std::min<uint64_t>(k, off_from + off_to + 1); }
        else { a = off_from; b = (count - 1) * k + off_to + 1; }
        const uint64_t len = b > a ? b - a : 0;
        *out_len = len;
        if (len > cap) return fail(E2FM_ERR_ARG, "output buffer too small");
        memcpy(out, h.data() + a, len);
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

uint64_t e2fm_result_patterns(const e2fm_result *r) { return r ? r->n : 0; }
uint64_t e2fm_result_total_positions(const e2fm_result *r) { return r ? r->total : 0; }

int e2fm_result_fetch(e2fm_result *r, uint64_t *counts, uint64_t *pos_offsets, uint64_t *positions, uint32_t *seq_index,
                      uint64_t *seq_offset) {
    if (!r) return fail(E2FM_ERR_ARG, "bad argument");
    if (!r->life || !r->life->alive) return fail(E2FM_ERR_ARG, "the index of this result has been closed");
    try {
        CK(cudaSetDevice(r->ix->device));
        cudaStream_t s = r->ix->stream;
        if (counts && r->n) CK(cudaMemcpyAsync(counts, r->d_counts, r->n * 8, cudaMemcpyDeviceToHost, s));
        if (pos_offsets) {
            if (r->located) CK(cudaMemcpyAsync(pos_offsets, r->d_off, (r->n + 1) * 8, cudaMemcpyDeviceToHost, s));
            else memset(pos_offsets, 0, (r->n + 1) * 8);
        }
        if (r->located && r->total) {
            if (positions) CK(cudaMemcpyAsync(positions, r->d_pos, r->total * 8, cudaMemcpyDeviceToHost, s));
            if (seq_index) CK(cudaMemcpyAsync(seq_index, r->d_seq, r->total * 4, cudaMemcpyDeviceToHost, s));
            if (seq_offset) CK(cudaMemcpyAsync(seq_offset, r->d_seqoff, r->total * 8, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaStreamSynchronize(s));
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

int e2fm_result_fetch_compact(e2fm_result *r, uint32_t *counts, uint32_t *pos_offsets, uint32_t *seq_index, uint32_t *seq_offset) {
    if (!r) return fail(E2FM_ERR_ARG, "bad argument");
    if (!r->life || !r->life->alive) return fail(E2FM_ERR_ARG, "the index of this result has been closed");
    if (r->total >= (1ull << 32)) return fail(E2FM_ERR_ARG, "more than 2^32-1 occurrences: use e2fm_result_fetch");
    try {
        e2fm_index *ix = r->ix;
        CK(cudaSetDevice(ix->device));
        cudaStream_t s = ix->stream;
        const uint64_t need = (counts ? r->n : 0) + (pos_offsets && r->located ? r->n + 1 : 0) + (seq_offset && r->located ? r->total : 0);
        WsBuf<uint32_t> tmp(ix, WS_NARROW, need + 4, s);
        uint32_t *q = tmp.p;
        if (counts && r->n) {
            launch_narrow(r->d_counts, q, r->n, s);
            CK(cudaMemcpyAsync(counts, q, r->n * 4, cudaMemcpyDeviceToHost, s));
            q += r->n;
        }
        if (pos_offsets) {
            if (r->located) {
                launch_narrow(reinterpret_cast<const unsigned long long *>(r->d_off), q, r->n + 1, s);
                CK(cudaMemcpyAsync(pos_offsets, q, (r->n + 1) * 4, cudaMemcpyDeviceToHost, s));
                q += r->n + 1;
            } else memset(pos_offsets, 0, (r->n + 1) * 4);
        }
        if (r->located && r->total) {
            if (seq_index) CK(cudaMemcpyAsync(seq_index, r->d_seq, r->total * 4, cudaMemcpyDeviceToHost, s));
            if (seq_offset) {
                if (ix->max_seq_len >= (1ull << 32)) return fail(E2FM_ERR_ARG, "a sequence is longer than 2^32-1: use e2fm_result_fetch");
                launch_narrow(reinterpret_cast<const unsigned long long *>(r->d_seqoff), q, r->total, s);
                CK(cudaMemcpyAsync(seq_offset, q, r->total * 4, cudaMemcpyDeviceToHost, s));
            }
        }
        CK(cudaStreamSynchronize(s));
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

int e2fm_result_device_ptrs(e2fm_result *r, const uint64_t **d_counts, const uint64_t **d_pos_offsets, const uint64_t **d_positions,
                            const uint32_t **d_seq_index, const uint64_t **d_seq_offset) {
    if (!r) return fail(E2FM_ERR_ARG, "bad argument");
    if (d_counts) *d_counts = reinterpret_cast<const uint64_t *>(r->d_counts);
    if (d_pos_offsets) *d_pos_offsets = r->d_off;
    if (d_positions) *d_positions = r->d_pos;
    if (d_seq_index) *d_seq_index = r->d_seq;
    if (d_seq_offset) *d_seq_offset = r->d_seqoff;
    return E2FM_OK;
}

void e2fm_result_free(e2fm_result *r) {
    if (!r) return;
    Lifetime *life = r->life;
    void *arrays[5] = {r->d_counts, r->d_off, r->d_pos, r->d_seq, r->d_seqoff};
    if (life) {
        cudaSetDevice(life->device);
        for (void *p : arrays) {
            if (!p) continue;
            if (life->alive) cudaFreeAsync(p, life->stream);
            else cudaFree(p);                    // the index (and its stream) are gone
        }
        if (--life->refs == 0) delete life;
        cudaGetLastError();
    }
    delete r;
}

int e2fm_internal_device(const e2fm_index *ix) { return ix ? ix->device : 0; }
void e2fm_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }
e2fm_result *e2fm_internal_make_result(e2fm_index *ix, uint64_t n, uint64_t total, bool located, unsigned long long *d_counts,
                                       uint64_t *d_csr, uint64_t *d_pos, uint32_t *d_seq, uint64_t *d_seqoff) {
    e2fm_result *r = new (std::nothrow) e2fm_result();
    if (!r) return nullptr;
    r->ix = ix;
    r->life = ix->life;
    ix->life->refs++;
    r->n = n;
    r->total = total;
    r->located = located;
    r->d_counts = d_counts;
    r->d_off = d_csr;
    r->d_pos = d_pos;
    r->d_seq = d_seq;
    r->d_seqoff = d_seqoff;
    return r;
}

int e2fm_last_batch_ms(const e2fm_index *ix, float out[8]) {
    if (!ix || !out) return fail(E2FM_ERR_ARG, "bad argument");
    memcpy(out, ix->batch_ms, sizeof(ix->batch_ms));
    return E2FM_OK;
}
int e2fm_last_batch_counters(const e2fm_index *ix, uint64_t out[16]) {
    if (!ix || !out) return fail(E2FM_ERR_ARG, "bad argument");
    memcpy(out, ix->batch_ctr, sizeof(ix->batch_ctr));
    return E2FM_OK;
}
int e2fm_set_option(e2fm_index *ix, int option, uint64_t value) {
    if (!ix) return fail(E2FM_ERR_ARG, "bad argument");
    switch (option) {
        case E2FM_OPT_CHUNK: ix->chunk_patterns = value ? value : (4u << 20); return E2FM_OK;
        case E2FM_OPT_WINDOW: ix->window_tasks = value ? value : (64u << 20); return E2FM_OK;
        case E2FM_OPT_DENSITY:   // test hook: forget the learned densities (value = per-mille of one task / result per pattern)
            ix->tasks_per_pattern = ix->results_per_pattern = ix->seed_tasks_per_pattern = (double)value / 1000.0;
            return E2FM_OK;
        case E2FM_OPT_PIPE: ix->pipe_patterns = value ? value : (1u << 18); return E2FM_OK;
        case E2FM_OPT_TAPER: ix->pipe_taper = value != 0; return E2FM_OK;
        case E2FM_OPT_HITS_STAGE:
            ix->hits_stage_patterns = value ? value : (1u << 20);
            ix->hits_stage_auto = value == 0;
            return E2FM_OK;
        case E2FM_OPT_PHASE_TIMERS: ix->phase_timers = (int)std::min<uint64_t>(value, 2); return E2FM_OK;
        case E2FM_OPT_SEED:
            if (value && !(ix->dix.seed_tab && ix->dix.packed_code)) return fail(E2FM_ERR_UNSUPPORTED, "the seed table was not built for this index");
            ix->use_seed = value != 0;
            return E2FM_OK;
        case E2FM_OPT_DENSE:
            if (value && !ix->dix.sa) return fail(E2FM_ERR_UNSUPPORTED, "the dense locate tables were not built for this index");
            ix->use_dense = value != 0;
            return E2FM_OK;
        default: return fail(E2FM_ERR_ARG, "unknown option");
    }
}

// ---------------------------------------------------------------- test hooks

int e2fm_debug_keystream(int device, const uint8_t key64[64], uint64_t nonce, uint32_t max_value, uint64_t start, uint64_t count,
                         uint32_t *out) {
    if (!key64 || (!out && count)) return fail(E2FM_ERR_ARG, "bad argument");
    cudaStream_t s;
    int rc = with_device(device, &s);
    if (rc) return rc;
    if (count == 0) return E2FM_OK;
    try {
        uint32_t kw[8];
        key_words(key64 + 32, kw);
        const uint32_t bits = int_log2((uint64_t)max_value + 1);   // RandomGenerator.cpp:136
        DBuf<uint32_t> d(count, s);
        launch_keystream_debug(kw, nonce, bits, start, count, d.p, s);
        CK(cudaMemcpyAsync(out, d.p, count * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

int e2fm_encode_buckets(int device, const uint8_t key64[64], const uint32_t *bucket_codes, uint64_t n_rows, uint64_t bucket_size,
                        uint64_t first_bucket, uint64_t n_buckets, uint64_t buckets_per_super_bucket, uint64_t total_buckets,
                        const uint32_t *bucket_alpha_size, uint64_t *compressed_length, uint64_t *payload_offset, uint8_t *payload,
                        uint64_t payload_cap, uint64_t *payload_bytes, uint32_t *sequence, float *device_ms) {
    if (!key64 || bucket_size == 0 || buckets_per_super_bucket == 0 || (n_buckets && (!bucket_codes || !bucket_alpha_size || !compressed_length)))
        return fail(E2FM_ERR_ARG, "bad argument");
    if (n_rows > n_buckets * bucket_size || (n_buckets && n_rows <= (n_buckets - 1) * bucket_size))
        return fail(E2FM_ERR_ARG, "n_rows does not fit n_buckets buckets of bucket_size rows (only the last one may be shorter)");
    if (first_bucket + n_buckets > total_buckets) return fail(E2FM_ERR_ARG, "first_bucket + n_buckets exceeds total_buckets");
    if (payload && !payload_offset) return fail(E2FM_ERR_ARG, "payload needs payload_offset");
    if (payload_bytes) *payload_bytes = 0;
    if (device_ms) *device_ms = 0;
    cudaStream_t s;
    int rc = with_device(device, &s);
    if (rc) return rc;
    if (n_buckets == 0) { if (payload_offset) payload_offset[0] = 0; return E2FM_OK; }
    try {
        uint32_t a_max = 0;
        for (uint64_t b = 0; b < n_buckets; b++) a_max = std::max(a_max, bucket_alpha_size[b]);
        if (a_max == 0 || a_max > 65535) return fail(E2FM_ERR_ARG, "bucket alphabet sizes must be in 1..65535");
        EncodeParams p{};
        p.n_rows = n_rows;
        p.b_size = bucket_size;
        p.first_bucket = first_bucket;
        p.n_buckets = n_buckets;
        p.per = buckets_per_super_bucket;
        p.total_buckets = total_buckets;
        p.list_elems = (a_max + 1) / 2 * 2;
        const int wpc = 4;
        const size_t smem = (size_t)wpc * 2 * p.list_elems * sizeof(uint16_t);
        if (smem > 200 * 1024) return fail(E2FM_ERR_UNSUPPORTED, "bucket alphabet too large for the encode kernel");
        CK(configure_bucket_encode(smem));
        uint32_t kw[8];
        key_words(key64 + 32, kw);                              // EncryptionManager: the poly-alphabetic key is bytes 32..63
        DBuf<uint32_t> d_codes(n_rows, s), d_alpha(n_buckets, s), d_status(1, s), d_seq(sequence ? n_rows : 0, s);
        DBuf<uint16_t> d_vals(n_rows + 8, s);
        DBuf<uint64_t> d_clen(n_buckets, s), d_off(n_buckets + 1, s);
        DBuf<uint8_t> d_payload(payload ? payload_cap + 8 : 0, s);
        d_status.zero();
        CK(cudaMemcpyAsync(d_codes.p, bucket_codes, n_rows * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_alpha.p, bucket_alpha_size, n_buckets * 4, cudaMemcpyHostToDevice, s));
        p.codes = d_codes.p;
        p.alpha = d_alpha.p;
        p.vals = d_vals.p;
        p.clen = d_clen.p;
        p.payload_off = d_off.p;
        p.payload = d_payload.p;
        p.payload_cap = payload ? payload_cap / 8 * 8 : 0;
        p.sequence = d_seq.p;
        p.status = d_status.p;
        struct EventPair {                                     // destroyed on every way out
            cudaEvent_t a = nullptr, b = nullptr;
            ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
        } ev;
        CK(cudaEventCreate(&ev.a));
        CK(cudaEventCreate(&ev.b));
        CK(cudaEventRecord(ev.a, s));
        launch_bucket_mtf_rle(p, wpc, smem, s);
        launch_payload_offsets(p, s);
        // key stream + substitution + packing, in batches of buckets (the key stream of a batch stays below 256 MB)
        const uint32_t kbits_max = int_log2((uint64_t)a_max + 1);
        const uint32_t blocks_per_bucket = (uint32_t)((bucket_size * kbits_max + 511) / 512 + 2);
        const uint32_t ks_stride = blocks_per_bucket * 64;
        const uint64_t batch = std::max<uint64_t>(1, std::min<uint64_t>(n_buckets, (256ull << 20) / ks_stride));
        DBuf<uint8_t> d_ks((size_t)batch * ks_stride + 64, s);
        for (uint64_t b0 = 0; b0 < n_buckets; b0 += batch) {
            const uint64_t nb = std::min(batch, n_buckets - b0);
            launch_keystream(kw, first_bucket + b0, nb, blocks_per_bucket, 0, reinterpret_cast<uint4 *>(d_ks.p), s);
            launch_bucket_substitute_pack(p, b0, nb, d_ks.p, ks_stride, s);
        }
        CK(cudaEventRecord(ev.b, s));
        uint32_t status = 0;
        std::vector<uint64_t> off(n_buckets + 1);
        CK(cudaMemcpyAsync(&status, d_status.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(compressed_length, d_clen.p, n_buckets * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(off.data(), d_off.p, (n_buckets + 1) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, ev.a, ev.b);
        if (device_ms) *device_ms = ms;
        if (payload_offset) memcpy(payload_offset, off.data(), (n_buckets + 1) * 8);
        if (payload_bytes) *payload_bytes = off[n_buckets];
        if (status & ENCODE_ERR_ALPHABET) return fail(E2FM_ERR_ARG, "a bucket alphabet size is 0");
        if (status & ENCODE_ERR_CODE) return fail(E2FM_ERR_ARG, "a row holds a code that is not below its bucket's alphabet size");
        if (status & ENCODE_ERR_BOUNDS) return fail(E2FM_ERR_FORMAT, "key stream too short for a bucket (internal)");
        if (status & ENCODE_ERR_CAPACITY) return fail(E2FM_ERR_ARG, "payload_cap is too small (see payload_bytes)");
        if (payload && off[n_buckets]) CK(cudaMemcpyAsync(payload, d_payload.p, off[n_buckets], cudaMemcpyDeviceToHost, s));
        if (sequence) CK(cudaMemcpyAsync(sequence, d_seq.p, n_rows * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    } catch (const std::exception &e) {
        cudaGetLastError();
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

int e2fm_debug_hits_stages(uint64_t n, uint64_t stage_option, uint64_t *bounds_out, uint64_t cap, uint64_t *n_bounds) {
    if (!n_bounds || (cap && !bounds_out)) return fail(E2FM_ERR_ARG, "bad argument");
    const std::vector<uint64_t> b = hits_stage_plan(n, stage_option ? stage_option : (1u << 20), stage_option == 0, nullptr);
    *n_bounds = b.size();
    if (b.size() > cap) return fail(E2FM_ERR_ARG, "bounds_out is too small (see n_bounds)");
    memcpy(bounds_out, b.data(), b.size() * 8);
    return E2FM_OK;
}

int e2fm_debug_gather_bench(int device, uint64_t buffer_bytes, uint64_t n_gathers, double *giga_gathers_per_s) {
    return e2fm_debug_gather_bench_ex(device, buffer_bytes, n_gathers, 0, giga_gathers_per_s);
}

int e2fm_debug_gather_bench_ex(int device, uint64_t buffer_bytes, uint64_t n_gathers, int variant, double *giga_gathers_per_s) {
    if (!giga_gathers_per_s || buffer_bytes < 4096 || n_gathers == 0) return fail(E2FM_ERR_ARG, "bad argument");
    cudaStream_t s;
    int rc = with_device(device, &s);
    if (rc) return rc;
    try {
        uint64_t n_words = 512;
        while (n_words * 2 * 8 <= buffer_bytes) n_words *= 2;      // largest power of two that fits
        uint64_t *buf = nullptr;
        unsigned long long *sink = nullptr;
        CK(cudaMalloc((void **)&buf, n_words * 8));
        CK(cudaMalloc((void **)&sink, 8));
        CK(cudaMemset(buf, 1, n_words * 8));
        const uint32_t per_thread = 32;
        const uint64_t n_threads = (n_gathers + per_thread - 1) / per_thread;
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        launch_gather_bench(buf, n_words, n_threads, per_thread, sink, variant, s);   // warm-up
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(a, s));
            launch_gather_bench(buf, n_words, n_threads, per_thread, sink, variant, s);
            CK(cudaEventRecord(b, s));
            CK(cudaEventSynchronize(b));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, a, b));
            best = std::min(best, ms);
        }
        CK(cudaGetLastError());
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        cudaFree(buf);
        cudaFree(sink);
        *giga_gathers_per_s = (double)(n_threads * per_thread) / (best * 1e-3) / 1e9;
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

int e2fm_debug_salsa20(int device, const uint8_t key32[32], uint64_t nonce, uint64_t first_block, uint64_t n_blocks, uint8_t *out) {
    if (!key32 || (!out && n_blocks)) return fail(E2FM_ERR_ARG, "bad argument");
    cudaStream_t s;
    int rc = with_device(device, &s);
    if (rc) return rc;
    if (n_blocks == 0) return E2FM_OK;
    try {
        uint32_t kw[8];
        key_words(key32, kw);
        DBuf<uint32_t> d(n_blocks * 16, s);
        launch_salsa_debug(kw, nonce, first_block, n_blocks, d.p, s);
        CK(cudaMemcpyAsync(out, d.p, n_blocks * 64, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
    } catch (const std::exception &e) {
        return fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

int e2fm_debug_bwt(const e2fm_index *ix, uint64_t row0, uint64_t count, uint32_t *out) {
    return debug_call(ix, [&](cudaStream_t s) {
        DBuf<uint32_t> d(std::max<uint64_t>(count, 1), s);
        launch_debug_bwt(ix->dix, row0, count, d.p, s);
        if (count) CK(cudaMemcpyAsync(out, d.p, count * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    });
}
int e2fm_debug_marks(const e2fm_index *ix, uint64_t row0, uint64_t count, uint64_t *out) {
    return debug_call(ix, [&](cudaStream_t s) {
        DBuf<uint64_t> d(std::max<uint64_t>(count, 1), s);
        launch_debug_marks(ix->dix, row0, count, d.p, s);
        if (count) CK(cudaMemcpyAsync(out, d.p, count * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    });
}
int e2fm_debug_lf(const e2fm_index *ix, const uint64_t *rows, uint64_t count, uint64_t *out) {
    return debug_call(ix, [&](cudaStream_t s) {
        DBuf<uint64_t> d_in(std::max<uint64_t>(count, 1), s), d(std::max<uint64_t>(count, 1), s);
        if (count) CK(cudaMemcpyAsync(d_in.p, rows, count * 8, cudaMemcpyHostToDevice, s));
        launch_debug_lf(ix->dix, d_in.p, count, d.p, s);
        if (count) CK(cudaMemcpyAsync(out, d.p, count * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    });
}
int e2fm_debug_rank(const e2fm_index *ix, const uint32_t *codes, const uint64_t *rows, uint64_t count, uint64_t *out) {
    return debug_call(ix, [&](cudaStream_t s) {
        DBuf<uint32_t> d_c(std::max<uint64_t>(count, 1), s);
        DBuf<uint64_t> d_in(std::max<uint64_t>(count, 1), s), d(std::max<uint64_t>(count, 1), s);
        if (count) {
            CK(cudaMemcpyAsync(d_c.p, codes, count * 4, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_in.p, rows, count * 8, cudaMemcpyHostToDevice, s));
        }
        launch_debug_rank(ix->dix, d_c.p, d_in.p, count, d.p, s);
        if (count) CK(cudaMemcpyAsync(out, d.p, count * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    });
}

int e2fm_debug_seed(const e2fm_index *ix, const uint32_t *codes, uint64_t count, uint32_t *out) {
    if (!ix || !ix->dix.seed_tab) return fail(E2FM_ERR_UNSUPPORTED, "the seed table was not built for this index");
    return debug_call(ix, [&](cudaStream_t s) {
        const uint32_t j = ix->dix.seed_chars;
        DBuf<uint32_t> d_c(std::max<uint64_t>(count * j, 1), s), d(std::max<uint64_t>(count * 3, 1), s);
        if (count) CK(cudaMemcpyAsync(d_c.p, codes, count * j * 4, cudaMemcpyHostToDevice, s));
        launch_debug_seed(ix->dix, d_c.p, count, d.p, s);
        if (count) CK(cudaMemcpyAsync(out, d.p, count * 3 * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    });
}

}  // extern "C"
