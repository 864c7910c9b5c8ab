// shardset.cu — a collection sharded by sequence over the GPUs of one box, below the C ABI.
//
// One process per GPU. Rank r holds the E2FM index the reference built for the r-th contiguous group of sequences; exact patterns
// never span a terminator and every sequence starts k-aligned (EFMCollection.cpp:218-225), so an occurrence has the same
// (sequence, offset) whether its sequence is indexed alone, in a shard or in the whole collection; adding the shard's text base
// gives the position the unsharded reference index reports, and because shards are contiguous sequence ranges the per-pattern
// concatenation in shard order is already sorted (EFMIndex.cpp:2291). e2fm_shardset_search is ONE exchange step in which every
// rank supplies 1/N of the batch and receives the merged answers for that same 1/N:
//
//   H2D of this rank's slice of the batch  ->  ncclAllGather over NVLink (every rank has the whole batch)
//   local search of the whole batch on the shard (seed search / generic path, api.cu)
//   counts:      ncclReduceScatter(sum, u64): the counts of slice q land on rank q   (EFMIndex::countOccurrences over the collection)
//   occurrences: every rank cuts its CSR at the slice boundaries; ONE grouped ncclSend/ncclRecv round moves the offsets and the
//                (position + text base, sequence + first sequence, offset) triples of slice q to rank q (an all-to-all-v: a rank
//                receives 1/N of the occurrences, nothing is replicated, nothing is zero-filled);
//                shard_merge_kernel then places shard r's occurrences of a pattern behind those of shards 0..r-1. The merged offset of
//                a pattern is the SUM over the shards of its offsets in their CSRs: no scan is needed.
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
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
    *(void **)&api.ReduceScatter = sym("ncclReduceScatter");
    *(void **)&api.Send = sym("ncclSend");
    *(void **)&api.Recv = sym("ncclRecv");
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
    Grow batch, cpad, off32, roff, stage_pos, stage_seq, stage_off, cuts;
    cudaEvent_t ev[6] = {nullptr};
    float ms[8] = {0};
    unsigned long long *h_cuts = nullptr;     // pinned: world x (world + 1) CSR cuts of every rank at the slice boundaries
};

// ---------------------------------------------------------------- kernels of the merge

namespace {

// this shard's occurrences moved into the collection's coordinates, its CSR in 32 bits (what travels), the counts padded to
// world equal slices, and the CSR cut at the slice boundaries: cuts[q] = off[min(n, q * per)]
__global__ void __launch_bounds__(256) shard_local_kernel(const uint64_t *__restrict__ off, const unsigned long long *__restrict__ counts, uint64_t n,
                                                           uint64_t padded, uint32_t *__restrict__ off32, unsigned long long *__restrict__ cpad,
                                                           uint64_t *__restrict__ pos, uint32_t *__restrict__ seq, uint64_t total,
                                                           uint64_t text_base, uint32_t first_seq, uint64_t per, uint32_t world,
                                                           unsigned long long *__restrict__ cuts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (off && i <= n) off32[i] = (uint32_t)off[i];
    if (i < padded) cpad[i] = i < n ? counts[i] : 0ull;
    if (off && i <= world) cuts[i] = off[min(n, i * per)];
    if (off && i < total) {
        pos[i] += text_base;
        const uint32_t s = seq[i];
        if (s != 0xFFFFFFFFu) seq[i] = s + first_seq;     // sequenceIndex -1 (no terminator beyond the position, EFMCollection.cpp:124-133) stays
    }
}

struct MergeSrc {                                       // what one shard sent for this rank's slice
    const uint32_t *off;                                // its CSR offsets of the slice's patterns (mine + 1 entries, absolute in ITS arrays)
    const uint64_t *pos;
    const uint32_t *seq;
    const uint64_t *seqoff;
};
struct MergeArgs { MergeSrc src[16]; };

// pattern i of the slice: its merged offset is the sum over the shards of (their offset - their offset of the slice's first pattern);
// shard r's occurrences go behind those of shards 0..r-1
__global__ void __launch_bounds__(256) shard_merge_kernel(MergeArgs a, uint32_t world, uint64_t mine, uint64_t *__restrict__ m_csr,
                                                           uint64_t *__restrict__ m_pos, uint32_t *__restrict__ m_seq, uint64_t *__restrict__ m_off) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > mine) return;
    uint64_t dst = 0;
    for (uint32_t r = 0; r < world; r++) dst += a.src[r].off[i] - a.src[r].off[0];
    m_csr[i] = dst;
    if (i == mine) return;
    for (uint32_t r = 0; r < world; r++) {
        const uint32_t b = a.src[r].off[i] - a.src[r].off[0], c = a.src[r].off[i + 1] - a.src[r].off[i];
        for (uint32_t q = 0; q < c; q++) {
            m_pos[dst + q] = a.src[r].pos[b + q];
            m_seq[dst + q] = a.src[r].seq[b + q];
            m_off[dst + q] = a.src[r].seqoff[b + q];
        }
        dst += c;
    }
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
        if (world > 16) throw CudaFail("at most 16 ranks per shard set");
        CU(cudaHostAlloc((void **)&ss->h_cuts, sizeof(unsigned long long) * (size_t)world * (size_t)(world + 1), cudaHostAllocDefault));
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
    for (Grow *g : {&ss->batch, &ss->cpad, &ss->off32, &ss->roff, &ss->stage_pos, &ss->stage_seq, &ss->stage_off, &ss->cuts}) g->release();
    for (auto &e : ss->ev) if (e) cudaEventDestroy(e);
    if (ss->h_cuts) cudaFreeHost(ss->h_cuts);
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
        const bool locate = mode == E2FM_MODE_LOCATE;
        const uint64_t padded = per * world;
        // ---- this shard in the collection's coordinates; counts padded to equal slices; CSR cuts at the slice boundaries
        unsigned long long *cpad = (unsigned long long *)ss->cpad.ensure(std::max<uint64_t>(padded, 1) * 8, s);
        uint32_t *off32 = (uint32_t *)ss->off32.ensure((n_total + 2) * 4, s);
        unsigned long long *cuts = (unsigned long long *)ss->cuts.ensure((size_t)world * (world + 1) * 8, s);
        unsigned long long *my_cuts = cuts + (size_t)ss->rank * (world + 1);
        {
            const uint64_t work = std::max<uint64_t>(std::max<uint64_t>(padded, n_total + 1), std::max<uint64_t>(l_total, world + 1));
            shard_local_kernel<<<(uint32_t)((work + 255) / 256), 256, 0, s>>>(locate ? l_off : nullptr, reinterpret_cast<const unsigned long long *>(l_counts),
                                                                            n_total, padded, off32, cpad, const_cast<uint64_t *>(l_pos),
                                                                            const_cast<uint32_t *>(l_seq), l_total, ss->text_base, (uint32_t)ss->first_seq,
                                                                            per, world, my_cuts);
        }
        // ---- counts: the sums of slice q land on rank q
        CU(cudaMallocAsync((void **)&m_counts, std::max<uint64_t>(per, 1) * 8, s));
        if (per) {
            if (world > 1) NC(api.ReduceScatter(cpad, m_counts, per, ncclUint64, ncclSum, ss->comm, s));
            else CU(cudaMemcpyAsync(m_counts, cpad, per * 8, cudaMemcpyDeviceToDevice, s));
        }
        uint64_t m_total = 0;
        if (locate) {
            // ---- how many occurrences every rank holds for every slice
            if (world > 1) NC(api.AllGather(my_cuts, cuts, world + 1, ncclUint64, ss->comm, s));
            CU(cudaMemcpyAsync(ss->h_cuts, cuts, (size_t)world * (world + 1) * 8, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));                       // the cuts size the staging and merged arrays
            auto cut = [&](uint32_t r, uint32_t q) { return ss->h_cuts[(size_t)r * (world + 1) + q]; };
            std::vector<uint64_t> rbase(world + 1, 0);
            for (uint32_t r = 0; r < world; r++) rbase[r + 1] = rbase[r] + (cut(r, ss->rank + 1) - cut(r, ss->rank));
            m_total = rbase[world];
            uint32_t *roff = (uint32_t *)ss->roff.ensure((size_t)world * (per + 1) * 4 + 16, s);
            uint64_t *st_pos = (uint64_t *)ss->stage_pos.ensure(std::max<uint64_t>(m_total, 1) * 8, s);
            uint32_t *st_seq = (uint32_t *)ss->stage_seq.ensure(std::max<uint64_t>(m_total, 1) * 4, s);
            uint64_t *st_off = (uint64_t *)ss->stage_off.ensure(std::max<uint64_t>(m_total, 1) * 8, s);
            // ---- one grouped round: slice q of my CSR and of my occurrence arrays to rank q
            if (world > 1) {
                NC(api.GroupStart());
                for (uint32_t q = 0; q < world; q++) {
                    uint64_t cnt_q = 0;
                    const uint64_t lo_q = e2fm_shardset_slice(n_total, ss->world, (int)q, &cnt_q);
                    const uint64_t a0 = cut(ss->rank, q), sz = cut(ss->rank, q + 1) - a0;
                    NC(api.Send(off32 + lo_q, cnt_q + 1, ncclUint32, (int)q, ss->comm, s));
                    if (sz) {
                        NC(api.Send(l_pos + a0, sz, ncclUint64, (int)q, ss->comm, s));
                        NC(api.Send(l_seq + a0, sz, ncclUint32, (int)q, ss->comm, s));
                        NC(api.Send(l_seqoff + a0, sz, ncclUint64, (int)q, ss->comm, s));
                    }
                }
                for (uint32_t r = 0; r < world; r++) {
                    const uint64_t sz = rbase[r + 1] - rbase[r];
                    NC(api.Recv(roff + (size_t)r * (per + 1), mine + 1, ncclUint32, (int)r, ss->comm, s));
                    if (sz) {
                        NC(api.Recv(st_pos + rbase[r], sz, ncclUint64, (int)r, ss->comm, s));
                        NC(api.Recv(st_seq + rbase[r], sz, ncclUint32, (int)r, ss->comm, s));
                        NC(api.Recv(st_off + rbase[r], sz, ncclUint64, (int)r, ss->comm, s));
                    }
                }
                NC(api.GroupEnd());
            }
            CU(cudaEventRecord(ss->ev[3], s));
            // ---- per-pattern concatenation in shard order
            CU(cudaMallocAsync((void **)&m_csr, (mine + 1) * 8, s));
            CU(cudaMallocAsync((void **)&m_pos, std::max<uint64_t>(m_total, 1) * 8, s));
            CU(cudaMallocAsync((void **)&m_seq, std::max<uint64_t>(m_total, 1) * 4, s));
            CU(cudaMallocAsync((void **)&m_off, std::max<uint64_t>(m_total, 1) * 8, s));
            MergeArgs ma;
            memset(&ma, 0, sizeof(ma));
            for (uint32_t r = 0; r < world; r++) {
                if (world > 1) ma.src[r] = MergeSrc{roff + (size_t)r * (per + 1), st_pos + rbase[r], st_seq + rbase[r], st_off + rbase[r]};
                else ma.src[r] = MergeSrc{off32, l_pos, l_seq, l_seqoff};
            }
            shard_merge_kernel<<<(uint32_t)((mine + 1 + 255) / 256), 256, 0, s>>>(ma, world, mine, m_csr, m_pos, m_seq, m_off);
        } else {
            CU(cudaEventRecord(ss->ev[3], s));
        }
        CU(cudaEventRecord(ss->ev[4], s));
        CU(cudaStreamSynchronize(s));
        CU(cudaGetLastError());
        e2fm_result_free(local);
        local = nullptr;
        *out = e2fm_internal_make_result(ss->ix, mine, m_total, locate, m_counts, m_csr, m_pos, m_seq, m_off);
        if (!*out) throw std::bad_alloc();
        memset(ss->ms, 0, sizeof(ss->ms));
        cudaEventElapsedTime(&ss->ms[0], ss->ev[0], ss->ev[1]);     // H2D of the slice + all-gather of the batch
        cudaEventElapsedTime(&ss->ms[1], ss->ev[1], ss->ev[2]);     // local search
        cudaEventElapsedTime(&ss->ms[2], ss->ev[2], ss->ev[3]);     // count reduce-scatter, CSR cuts, the send/recv round
        cudaEventElapsedTime(&ss->ms[3], ss->ev[3], ss->ev[4]);     // merge kernel
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
