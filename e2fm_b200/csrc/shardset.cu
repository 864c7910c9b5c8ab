// shardset.cu — a collection sharded by sequence over the GPUs of one box, below the C ABI.
//
// One process per GPU. Rank r holds the E2FM index the reference built for the r-th contiguous group of sequences; exact patterns
// never span a terminator and every sequence starts k-aligned (EFMCollection.cpp:218-225), so an occurrence has the same
// (sequence, offset) whether its sequence is indexed alone, in a shard or in the whole collection; adding the shard's text base
// gives the position the unsharded reference index reports, and because shards are contiguous sequence ranges the per-pattern
// concatenation in shard order is already sorted (EFMIndex.cpp:2291). e2fm_shardset_search is ONE exchange step:
//
//   H2D of this rank's 1/N of the batch  ->  ncclAllGather over NVLink (every rank has the whole batch)
//   local search on the shard (seed search / generic path, api.cu)
//   counts:      ncclAllReduce(sum, u64)                     = EFMIndex::countOccurrences over the whole collection
//   occurrences: ncclAllGather of the per-pattern occurrence counts of every shard -> merged CSR offsets (scan kernel);
//                each shard's (position + text base, sequence + first sequence, offset) arrays broadcast once, in one NCCL group;
//                merge kernels place shard r's occurrences of pattern p behind those of shards 0..r-1
//
// The NCCL communicator is owned here (ncclCommInitRank from a unique id the caller distributes). libnccl is resolved with dlopen
// at the first use, so single-GPU users of libe2fm_b200.so do not need it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/e2fm_b200.h"
#include "internal.h"
#include "search.cuh"

using namespace e2fm;

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.error = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
        return api;
    }
    auto sym = [&](const char *n) -> void * {
        void *p = dlsym(api.lib, n);
        if (!p) api.error = std::string("libnccl lacks ") + n;
        return p;
    };
    *(void **)&api.GetUniqueId = sym("ncclGetUniqueId");
    *(void **)&api.CommInitRank = sym("ncclCommInitRank");
    *(void **)&api.CommDestroy = sym("ncclCommDestroy");
    *(void **)&api.AllReduce = sym("ncclAllReduce");
    *(void **)&api.AllGather = sym("ncclAllGather");
    *(void **)&api.Broadcast = sym("ncclBroadcast");
    *(void **)&api.GroupStart = sym("ncclGroupStart");
    *(void **)&api.GroupEnd = sym("ncclGroupEnd");
    *(void **)&api.GetErrorString = sym("ncclGetErrorString");
    return api;
}

struct NcclError : std::runtime_error {
    explicit NcclError(const std::string &m) : std::runtime_error(m) {}
};
struct CudaFail : std::runtime_error {
    explicit CudaFail(const std::string &m) : std::runtime_error(m) {}
};

#define NC(call)                                                                                                   \
    do {                                                                                                           \
        ncclResult_t r_ = (call);                                                                                  \
        if (r_ != ncclSuccess) throw NcclError(std::string(#call) + ": " + nccl().GetErrorString(r_));             \
    } while (0)
#define CU(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) throw CudaFail(std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

// grow-only device buffer (plain cudaMalloc: NCCL buffers must not move while a collective is queued; freed at destroy)
struct Grow {
    void *p = nullptr;
    size_t bytes = 0;
    void *ensure(size_t need, cudaStream_t s) {
        if (need > bytes) {
            if (p) { CU(cudaStreamSynchronize(s)); CU(cudaFree(p)); p = nullptr; }
            const size_t want = need + need / 8 + 4096;
            CU(cudaMalloc(&p, want));
            bytes = want;
        }
        return p;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace

struct e2fm_shardset {
    e2fm_index *ix = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    uint64_t first_seq = 0, text_base = 0;
    cudaStream_t stream = nullptr;
    Grow batch, nocc_all, nocc_sum, roff, before, scratch, stage_pos, stage_seq, stage_off, totals;
    cudaEvent_t ev[6] = {nullptr};
    float ms[8] = {0};
    unsigned long long *h_totals = nullptr;   // pinned, world entries
};

// ---------------------------------------------------------------- kernels of the merge

namespace {

// occurrences of every pattern in this shard, and this shard's occurrences moved into the collection's coordinates
__global__ void __launch_bounds__(256) shard_local_kernel(const uint64_t *__restrict__ off, uint64_t n, uint32_t *__restrict__ nocc,
                                                           uint64_t *__restrict__ pos, uint32_t *__restrict__ seq, uint64_t total,
                                                           uint64_t text_base, uint32_t first_seq) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nocc[i] = (uint32_t)(off[i + 1] - off[i]);
    if (i < total) {
        pos[i] += text_base;
        const uint32_t s = seq[i];
        if (s != 0xFFFFFFFFu) seq[i] = s + first_seq;     // sequenceIndex -1 (no terminator beyond the position, EFMCollection.cpp:124-133) stays
    }
}

__global__ void __launch_bounds__(256) shard_sum_kernel(const uint32_t *__restrict__ nocc_all, uint64_t n, uint32_t world,
                                                         unsigned long long *__restrict__ sum) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long t = 0;
    for (uint32_t r = 0; r < world; r++) t += nocc_all[(uint64_t)r * n + i];
    sum[i] = t;
}

// shard r's occurrences of pattern p go behind those of the shards before it: merged[moff[p] + before[p] + i] = stage_r[roff[p] + i]
__global__ void __launch_bounds__(256) shard_merge_kernel(const uint32_t *__restrict__ nocc_r, const uint64_t *__restrict__ roff,
                                                           const uint64_t *__restrict__ moff, uint32_t *__restrict__ before, uint64_t n,
                                                           const uint64_t *__restrict__ s_pos, const uint32_t *__restrict__ s_seq,
                                                           const uint64_t *__restrict__ s_off, uint64_t *__restrict__ m_pos,
                                                           uint32_t *__restrict__ m_seq, uint64_t *__restrict__ m_off) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t c = nocc_r[p];
    if (c == 0) return;
    const uint32_t b = before[p];
    const uint64_t src = roff[p], dst = moff[p] + b;
    for (uint32_t i = 0; i < c; i++) {
        m_pos[dst + i] = s_pos[src + i];
        m_seq[dst + i] = s_seq[src + i];
        m_off[dst + i] = s_off[src + i];
    }
    before[p] = b + c;
}

__global__ void __launch_bounds__(256) shard_total_kernel(const uint64_t *__restrict__ off_end, unsigned long long *__restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = *off_end;
}

thread_local std::string g_ss_err;
int ss_fail(int code, const std::string &m) {
    e2fm_internal_set_error(m.c_str());
    return code;
}

}  // namespace

extern "C" {

int e2fm_shardset_unique_id(uint8_t id_out[128]) {
    if (!id_out) return ss_fail(E2FM_ERR_ARG, "bad argument");
    NcclApi &api = nccl();
    if (!api.error.empty() || !api.lib) return ss_fail(E2FM_ERR_NCCL, api.error);
    ncclUniqueId id;
    const ncclResult_t r = api.GetUniqueId(&id);
    if (r != ncclSuccess) return ss_fail(E2FM_ERR_NCCL, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id_out, &id, 128);
    return E2FM_OK;
}

int e2fm_shardset_create(e2fm_index *ix, const uint8_t id[128], int rank, int world, uint64_t first_sequence, uint64_t text_base,
                         e2fm_shardset **out) {
    if (!ix || !id || !out || world < 1 || rank < 0 || rank >= world) return ss_fail(E2FM_ERR_ARG, "bad argument");
    *out = nullptr;
    NcclApi &api = nccl();
    if (!api.error.empty() || !api.lib) return ss_fail(E2FM_ERR_NCCL, api.error);
    e2fm_shardset *ss = new (std::nothrow) e2fm_shardset();
    if (!ss) return ss_fail(E2FM_ERR_OOM, "out of host memory");
    try {
        ss->ix = ix;
        ss->rank = rank;
        ss->world = world;
        ss->first_seq = first_sequence;
        ss->text_base = text_base;
        ss->device = e2fm_internal_device(ix);
        ss->stream = (cudaStream_t)e2fm_index_stream(ix);
        CU(cudaSetDevice(ss->device));
        ncclUniqueId nid;
        memcpy(&nid, id, 128);
        NC(api.CommInitRank(&ss->comm, world, nid, rank));
        for (auto &e : ss->ev) CU(cudaEventCreate(&e));
        CU(cudaHostAlloc((void **)&ss->h_totals, sizeof(unsigned long long) * (size_t)(world + 1), cudaHostAllocDefault));
    } catch (const NcclError &e) {
        e2fm_shardset_destroy(ss);
        return ss_fail(E2FM_ERR_NCCL, e.what());
    } catch (const std::exception &e) {
        e2fm_shardset_destroy(ss);
        return ss_fail(E2FM_ERR_CUDA, e.what());
    }
    *out = ss;
    return E2FM_OK;
}

void e2fm_shardset_destroy(e2fm_shardset *ss) {
    if (!ss) return;
    cudaSetDevice(ss->device);
    if (ss->stream) cudaStreamSynchronize(ss->stream);
    if (ss->comm) nccl().CommDestroy(ss->comm);
    for (Grow *g : {&ss->batch, &ss->nocc_all, &ss->nocc_sum, &ss->roff, &ss->before, &ss->scratch, &ss->stage_pos, &ss->stage_seq,
                    &ss->stage_off, &ss->totals})
        g->release();
    for (auto &e : ss->ev) if (e) cudaEventDestroy(e);
    if (ss->h_totals) cudaFreeHost(ss->h_totals);
    cudaGetLastError();
    delete ss;
}

int e2fm_shardset_last_ms(const e2fm_shardset *ss, float out[8]) {
    if (!ss || !out) return ss_fail(E2FM_ERR_ARG, "bad argument");
    memcpy(out, ss->ms, sizeof(ss->ms));
    return E2FM_OK;
}

uint64_t e2fm_shardset_slice(uint64_t n_total, int world, int rank, uint64_t *count) {
    const uint64_t per = world > 0 ? (n_total + (uint64_t)world - 1) / (uint64_t)world : n_total;
    const uint64_t lo = std::min<uint64_t>(n_total, per * (uint64_t)rank);
    if (count) *count = std::min<uint64_t>(n_total, lo + per) - lo;
    return lo;
}

int e2fm_shardset_search(e2fm_shardset *ss, const uint8_t *slice, uint64_t n_total, uint32_t length, int packed, int mode,
                         e2fm_result **out) {
    if (!ss || !out || length == 0 || (mode != E2FM_MODE_COUNT && mode != E2FM_MODE_LOCATE)) return ss_fail(E2FM_ERR_ARG, "bad argument");
    *out = nullptr;
    NcclApi &api = nccl();
    e2fm_result *local = nullptr;
    uint64_t *m_pos = nullptr, *m_off = nullptr, *m_csr = nullptr;
    uint32_t *m_seq = nullptr;
    unsigned long long *m_counts = nullptr;
    cudaStream_t s = ss->stream;
    try {
        CU(cudaSetDevice(ss->device));
        const uint32_t world = (uint32_t)ss->world;
        const uint64_t unit = packed ? e2fm_packed_stride(length) : length;
        const uint64_t per = (n_total + world - 1) / world;
        uint64_t mine = 0;
        const uint64_t lo = e2fm_shardset_slice(n_total, ss->world, ss->rank, &mine);
        if (mine && !slice) return ss_fail(E2FM_ERR_ARG, "pattern slice missing");
        // ---- this rank's slice to the device, the whole batch over NVLink
        CU(cudaEventRecord(ss->ev[0], s));
        uint8_t *d_batch = (uint8_t *)ss->batch.ensure(per * world * unit + 64, s);
        if (mine) CU(cudaMemcpyAsync(d_batch + lo * unit, slice, mine * unit, cudaMemcpyHostToDevice, s));
        if (world > 1 && per) NC(api.AllGather(d_batch + (uint64_t)ss->rank * per * unit, d_batch, per * unit, ncclUint8, ss->comm, s));
        CU(cudaEventRecord(ss->ev[1], s));
        // ---- local search of the whole batch against this shard
        int rc;
        if (packed) rc = e2fm_search_batch_packed_device(ss->ix, d_batch, n_total, length, mode, &local);
        else rc = e2fm_search_batch_device_fixed(ss->ix, d_batch, n_total, length, mode, &local);
        if (rc != E2FM_OK) return rc;
        CU(cudaEventRecord(ss->ev[2], s));
        const uint64_t *l_counts, *l_off, *l_pos, *l_seqoff;
        const uint32_t *l_seq;
        e2fm_result_device_ptrs(local, &l_counts, &l_off, &l_pos, &l_seq, &l_seqoff);
        const uint64_t l_total = e2fm_result_total_positions(local);
        // ---- counts: all-reduce straight out of the local result into the merged result's array
        CU(cudaMallocAsync((void **)&m_counts, std::max<uint64_t>(n_total, 1) * 8, s));
        if (n_total) {
            if (world > 1) NC(api.AllReduce(l_counts, m_counts, n_total, ncclUint64, ncclSum, ss->comm, s));
            else CU(cudaMemcpyAsync(m_counts, l_counts, n_total * 8, cudaMemcpyDeviceToDevice, s));
        }
        uint64_t m_total = 0;
        if (mode == E2FM_MODE_LOCATE) {
            const uint32_t blocks_n = (uint32_t)((std::max<uint64_t>(n_total, l_total) + 255) / 256);
            uint32_t *nocc_all = (uint32_t *)ss->nocc_all.ensure((size_t)world * std::max<uint64_t>(n_total, 1) * 4, s);
            uint32_t *nocc_mine = nocc_all + (uint64_t)ss->rank * n_total;
            if (blocks_n)
                shard_local_kernel<<<blocks_n, 256, 0, s>>>(l_off, n_total, nocc_mine, const_cast<uint64_t *>(l_pos), const_cast<uint32_t *>(l_seq),
                                                            l_total, ss->text_base, (uint32_t)ss->first_seq);
            // occurrences per pattern of every shard + every shard's total
            unsigned long long *d_tot = (unsigned long long *)ss->totals.ensure((size_t)(world + 1) * 8, s);
            ss->h_totals[ss->rank] = l_total;
            CU(cudaMemcpyAsync(d_tot + ss->rank, ss->h_totals + ss->rank, 8, cudaMemcpyHostToDevice, s));
            if (world > 1) {
                NC(api.GroupStart());
                if (n_total) NC(api.AllGather(nocc_mine, nocc_all, n_total, ncclUint32, ss->comm, s));
                NC(api.AllGather(d_tot + ss->rank, d_tot, 1, ncclUint64, ss->comm, s));
                NC(api.GroupEnd());
            }
            CU(cudaMemcpyAsync(ss->h_totals, d_tot, (size_t)world * 8, cudaMemcpyDeviceToHost, s));
            // merged CSR offsets
            unsigned long long *sum = (unsigned long long *)ss->nocc_sum.ensure(std::max<uint64_t>(n_total, 1) * 8, s);
            uint64_t *scratch = (uint64_t *)ss->scratch.ensure(scan_scratch_elems(std::max<uint64_t>(n_total, 1)) * 8, s);
            CU(cudaMallocAsync((void **)&m_csr, (n_total + 1) * 8, s));
            if (n_total) shard_sum_kernel<<<(uint32_t)((n_total + 255) / 256), 256, 0, s>>>(nocc_all, n_total, world, sum);
            launch_scan_u64(sum, m_csr, n_total, scratch, s);
            CU(cudaStreamSynchronize(s));                       // the shards' totals size the staging and merged arrays
            uint64_t biggest = 0;
            for (uint32_t r = 0; r < world; r++) { m_total += ss->h_totals[r]; biggest = std::max<uint64_t>(biggest, ss->h_totals[r]); }
            CU(cudaMallocAsync((void **)&m_pos, std::max<uint64_t>(m_total, 1) * 8, s));
            CU(cudaMallocAsync((void **)&m_seq, std::max<uint64_t>(m_total, 1) * 4, s));
            CU(cudaMallocAsync((void **)&m_off, std::max<uint64_t>(m_total, 1) * 8, s));
            // every shard's occurrence arrays, once, to every rank
            uint64_t *st_pos = (uint64_t *)ss->stage_pos.ensure(std::max<uint64_t>(m_total, 1) * 8, s);
            uint32_t *st_seq = (uint32_t *)ss->stage_seq.ensure(std::max<uint64_t>(m_total, 1) * 4, s);
            uint64_t *st_off = (uint64_t *)ss->stage_off.ensure(std::max<uint64_t>(m_total, 1) * 8, s);
            std::vector<uint64_t> base(world + 1, 0);
            for (uint32_t r = 0; r < world; r++) base[r + 1] = base[r] + ss->h_totals[r];
            if (world > 1) {
                NC(api.GroupStart());
                for (uint32_t r = 0; r < world; r++) {
                    const uint64_t t = ss->h_totals[r];
                    if (t == 0) continue;
                    const bool me = (int)r == ss->rank;
                    NC(api.Broadcast(me ? (const void *)l_pos : (const void *)(st_pos + base[r]), st_pos + base[r], t, ncclUint64, (int)r, ss->comm, s));
                    NC(api.Broadcast(me ? (const void *)l_seq : (const void *)(st_seq + base[r]), st_seq + base[r], t, ncclUint32, (int)r, ss->comm, s));
                    NC(api.Broadcast(me ? (const void *)l_seqoff : (const void *)(st_off + base[r]), st_off + base[r], t, ncclUint64, (int)r, ss->comm, s));
                }
                NC(api.GroupEnd());
            } else if (l_total) {
                CU(cudaMemcpyAsync(st_pos, l_pos, l_total * 8, cudaMemcpyDeviceToDevice, s));
                CU(cudaMemcpyAsync(st_seq, l_seq, l_total * 4, cudaMemcpyDeviceToDevice, s));
                CU(cudaMemcpyAsync(st_off, l_seqoff, l_total * 8, cudaMemcpyDeviceToDevice, s));
            }
            CU(cudaEventRecord(ss->ev[3], s));
            // per-pattern concatenation in shard order
            uint32_t *before = (uint32_t *)ss->before.ensure(std::max<uint64_t>(n_total, 1) * 4, s);
            CU(cudaMemsetAsync(before, 0, std::max<uint64_t>(n_total, 1) * 4, s));
            uint64_t *roff = (uint64_t *)ss->roff.ensure((n_total + 1) * 8, s);
            for (uint32_t r = 0; r < world && n_total; r++) {
                if (ss->h_totals[r] == 0) continue;
                const uint32_t *nocc_r = nocc_all + (uint64_t)r * n_total;
                launch_ulen_widen(nocc_r, sum, n_total, s);
                launch_scan_u64(sum, roff, n_total, scratch, s);
                shard_merge_kernel<<<(uint32_t)((n_total + 255) / 256), 256, 0, s>>>(nocc_r, roff, m_csr, before, n_total, st_pos + base[r],
                                                                                  st_seq + base[r], st_off + base[r], m_pos, m_seq, m_off);
            }
        } else {
            CU(cudaEventRecord(ss->ev[3], s));
        }
        CU(cudaEventRecord(ss->ev[4], s));
        CU(cudaStreamSynchronize(s));
        CU(cudaGetLastError());
        e2fm_result_free(local);
        local = nullptr;
        *out = e2fm_internal_make_result(ss->ix, n_total, m_total, mode == E2FM_MODE_LOCATE, m_counts, m_csr, m_pos, m_seq, m_off);
        if (!*out) throw std::bad_alloc();
        memset(ss->ms, 0, sizeof(ss->ms));
        cudaEventElapsedTime(&ss->ms[0], ss->ev[0], ss->ev[1]);     // H2D of the slice + all-gather of the batch
        cudaEventElapsedTime(&ss->ms[1], ss->ev[1], ss->ev[2]);     // local search
        cudaEventElapsedTime(&ss->ms[2], ss->ev[2], ss->ev[3]);     // count all-reduce, occurrence-count all-gather, scan, broadcasts
        cudaEventElapsedTime(&ss->ms[3], ss->ev[3], ss->ev[4]);     // merge kernels
        cudaEventElapsedTime(&ss->ms[4], ss->ev[0], ss->ev[4]);     // whole step
    } catch (const NcclError &e) {
        if (local) e2fm_result_free(local);
        cudaStreamSynchronize(s);
        for (void *q : {(void *)m_counts, (void *)m_csr, (void *)m_pos, (void *)m_seq, (void *)m_off}) if (q) cudaFreeAsync(q, s);
        return ss_fail(E2FM_ERR_NCCL, e.what());
    } catch (const std::bad_alloc &) {
        if (local) e2fm_result_free(local);
        cudaStreamSynchronize(s);
        for (void *q : {(void *)m_counts, (void *)m_csr, (void *)m_pos, (void *)m_seq, (void *)m_off}) if (q) cudaFreeAsync(q, s);
        return ss_fail(E2FM_ERR_OOM, "out of host memory");
    } catch (const std::exception &e) {
        if (local) e2fm_result_free(local);
        cudaStreamSynchronize(s);
        for (void *q : {(void *)m_counts, (void *)m_csr, (void *)m_pos, (void *)m_seq, (void *)m_off}) if (q) cudaFreeAsync(q, s);
        return ss_fail(E2FM_ERR_CUDA, e.what());
    }
    return E2FM_OK;
}

}  // extern "C"
