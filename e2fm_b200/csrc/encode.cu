// encode.cu — the build-side inverse of bucket_decode_kernel: the text of every bucket, as the reference builder writes it.
//
// Replaces EFMIndex::encodeBuckets (EFMIndex.cpp:751-773: one BucketEncoder thread per range of buckets) -> Bucket::encodeText
// (Bucket.cpp:442-536: single-pass move-to-front, RLE0 of the runs of rank 0, poly-alphabetic substitution with the bucket's
// Salsa20 key stream, EFMIndex::getCryptoParams :1668) and the symbol loop of Bucket::writeText (Bucket.cpp:232-246: int_log2(
// bucketAlphaSize) bits per symbol, most significant bit first, flushed to a byte). BWT construction, the remappings and the
// headers stay in the reference builder; this is the step that parallelises per bucket exactly like K2.
//
//   bucket_mtf_rle_kernel     one WARP per bucket. The MTF list and its inverse (symbol -> rank) sit in shared memory, so
//                             ArrayMTFList::indexOf is one broadcast read; moveToFront shifts list[0, rank) 32 slots at a time
//                             from the top chunk down and renumbers the inverse on the way. Runs of rank 0 are counted and
//                             written as the digits of RLE0 (m + 1 in binary without its leading 1, most significant first,
//                             :507-520), a rank > 0 as rank + 1. Output: the plain sequence (u16 per symbol) + its length.
//   payload_offsets_kernel    exclusive scan over the buckets: where each bucket's packed text starts (8-byte aligned slots)
//   bucket_substitute_pack_kernel  one warp per bucket: (value + key[i]) mod (alphaSize + 1) with key[i] the i-th
//                             int_log2(alphaSize + 1)-bit field of the bucket's key stream (RandomGenerator.cpp:136), then the
//                             bit packing, a 32-bit word of the payload per lane.
// HBM traffic per row: 4 B code in, 2 B + 2 B scratch, <= 2 B payload out, <= 2 B key stream: the kernels are bound by the
// dependent shared-memory chain of the MTF update (like K2), not by memory.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "bits.h"
#include "encode.cuh"

namespace e2fm {

__device__ __forceinline__ uint32_t enc_bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// n (1..32) bits at bit offset `bit` of a big-endian bit stream in 4-byte aligned memory (as index_build.cu's dev_bits)
__device__ __forceinline__ uint32_t enc_bits(const uint8_t *base, uint64_t bit, uint32_t n) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base) + (bit >> 5);
    const uint64_t v = ((uint64_t)enc_bswap32(__ldg(w)) << 32) | enc_bswap32(__ldg(w + 1));
    const uint32_t o = (uint32_t)bit & 31u;
    const uint32_t r = (uint32_t)(v >> (64 - o - n));
    return n == 32 ? r : (r & ((1u << n) - 1));
}

__device__ __forceinline__ bool bucket_is_odd(const EncodeParams &p, uint64_t bn) {
    return (bn % 2 == 0) && (bn % p.per != 0) && (bn != p.total_buckets - 1);          // EFMIndex.cpp:898-901
}

__global__ void __launch_bounds__(128) bucket_mtf_rle_kernel(EncodeParams p) {
    extern __shared__ __align__(16) uint8_t enc_smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint64_t local = (uint64_t)blockIdx.x * wpc + warp;
    if (local >= p.n_buckets) return;
    const uint64_t bn = p.first_bucket + local;
    uint16_t *list = reinterpret_cast<uint16_t *>(enc_smem) + (size_t)warp * 2 * p.list_elems;
    uint16_t *where = list + p.list_elems;
    const uint64_t row0 = local * p.b_size;
    const uint32_t length = (uint32_t)min(p.b_size, p.n_rows - row0);
    const uint32_t alpha = p.alpha[local];
    const bool odd = bucket_is_odd(p, bn);
    const uint32_t *codes = p.codes + row0;
    uint16_t *vals = p.vals + row0;
    if (alpha == 0 || alpha > p.list_elems) {
        if (lane == 0) { p.clen[local] = 0; atomicOr(p.status, ENCODE_ERR_ALPHABET); }
        return;
    }
    if (alpha == 1) {                                  // a single-character bucket has no text (Bucket.cpp:166-170)
        if (lane == 0) p.clen[local] = 0;
        return;
    }
    for (uint32_t q = lane; q < alpha; q += 32) { list[q] = (uint16_t)q; where[q] = (uint16_t)q; }   // ArrayMTFList ctor
    __syncwarp();
    uint32_t n_out = 0, run = 0;
    bool bad = false;
    auto flush_run = [&]() {                           // m = run: floor(log2(m + 1)) digits of m + 1 - 2^nb (:507-520)
        const uint32_t nb = 31u - (uint32_t)__clz(run + 1);
        const uint32_t N = run + 1 - (1u << nb);
        if (lane < nb) vals[n_out + lane] = (uint16_t)((N >> (nb - 1 - lane)) & 1u);
        n_out += nb;
        run = 0;
    };
    for (uint32_t base = 0; base < length; base += 32) {
        const uint32_t i = base + lane;
        uint32_t c = 0;
        if (i < length) c = __ldg(codes + (odd ? length - 1 - i : i));          // odd buckets are encoded last row first (:446-448)
        bad |= c >= alpha;
        const uint32_t cnt = min(32u, length - base);
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t cj = __shfl_sync(0xFFFFFFFFu, c, j);
            if (cj >= alpha) continue;                 // reported below
            const uint32_t pos = where[cj];            // ArrayMTFList::indexOf
            if (pos == 0) { run++; continue; }
            if (run) flush_run();
            // moveToFront (ArrayMTFList.cpp:32): list[0, pos) one slot up, top chunk first; the inverse follows
            for (int t = (int)((pos - 1) >> 5); t >= 0; --t) {
                const uint32_t idx = ((uint32_t)t << 5) + lane;
                uint16_t v = 0;
                if (idx < pos) v = list[idx];
                __syncwarp();
                if (idx < pos) { list[idx + 1] = v; where[v] = (uint16_t)(idx + 1); }
                __syncwarp();
            }
            if (lane == 0) { list[0] = (uint16_t)cj; where[cj] = 0; vals[n_out] = (uint16_t)(pos + 1); }
            __syncwarp();
            n_out++;
        }
    }
    if (run) flush_run();
    if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicOr(p.status, ENCODE_ERR_CODE);
    if (lane == 0) p.clen[local] = n_out;
}

// payload_off[b] = start of bucket b's packed text in the payload buffer: slots of whole 8-byte words
__global__ void __launch_bounds__(1024) payload_offsets_kernel(EncodeParams p) {
    __shared__ unsigned long long part[1024];
    const uint32_t t = threadIdx.x;
    const uint64_t per = (p.n_buckets + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = min(p.n_buckets, (uint64_t)t * per), hi = min(p.n_buckets, lo + per);
    unsigned long long sum = 0;
    for (uint64_t b = lo; b < hi; b++) {
        const uint32_t a = p.alpha[b];
        sum += a > 1 ? ((unsigned long long)p.clen[b] * int_log2(a) + 63) / 64 * 8 : 0ull;
    }
    part[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (uint32_t i = 0; i < blockDim.x; i++) { const unsigned long long v = part[i]; part[i] = run; run += v; }
        p.payload_off[p.n_buckets] = run;
    }
    __syncthreads();
    unsigned long long run = part[t];
    for (uint64_t b = lo; b < hi; b++) {
        const uint32_t a = p.alpha[b];
        p.payload_off[b] = run;
        run += a > 1 ? ((unsigned long long)p.clen[b] * int_log2(a) + 63) / 64 * 8 : 0ull;
    }
}

__global__ void __launch_bounds__(128) bucket_substitute_pack_kernel(EncodeParams p, uint64_t batch_first, uint64_t batch_count,
                                                                      const uint8_t *__restrict__ ks, uint32_t ks_stride_bytes) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint64_t in_batch = (uint64_t)blockIdx.x * wpc + warp;
    if (in_batch >= batch_count) return;
    const uint64_t local = batch_first + in_batch;
    const uint32_t alpha = p.alpha[local];
    if (alpha <= 1 || alpha > p.list_elems) return;
    const uint32_t clen = (uint32_t)p.clen[local];
    const uint64_t row0 = local * p.b_size;
    uint16_t *vals = p.vals + row0;
    const uint8_t *myks = ks + in_batch * ks_stride_bytes;
    const uint32_t tbits = int_log2(alpha);            // bitsPerTextCharacter (Bucket.cpp:232)
    const uint32_t kbits = int_log2((uint64_t)alpha + 1);   // key width (RandomGenerator.cpp:136)
    const uint32_t mod = alpha + 1;
    if ((uint64_t)clen * kbits + 64 > (uint64_t)ks_stride_bytes * 8) {
        if (lane == 0) atomicOr(p.status, ENCODE_ERR_BOUNDS);
        return;
    }
    // poly-alphabetic substitution (:480-481, :514-518; keyStartOffset = 0, the sequence is never longer than the key)
    for (uint32_t i = lane; i < clen; i += 32) {
        const uint32_t v = ((uint32_t)vals[i] + enc_bits(myks, (uint64_t)i * kbits, kbits)) % mod;
        vals[i] = (uint16_t)v;
        if (p.sequence) p.sequence[row0 + i] = v;
    }
    __syncwarp();
    if (!p.payload) return;
    // Bucket::writeText's symbol loop: lane w builds the w-th 32-bit word of the bucket's bit stream
    const uint64_t off = p.payload_off[local];
    if (off + ((uint64_t)clen * tbits + 63) / 64 * 8 > p.payload_cap) {
        if (lane == 0) atomicOr(p.status, ENCODE_ERR_CAPACITY);
        return;
    }
    uint32_t *out = reinterpret_cast<uint32_t *>(p.payload + off);
    const uint32_t total_bits = clen * tbits;
    const uint32_t n_words = (total_bits + 63) / 64 * 2;
    for (uint32_t w = lane; w < n_words; w += 32) {
        const uint32_t b0 = w * 32, b1 = b0 + 32;
        uint32_t word = 0;
        for (uint32_t sym = b0 / tbits; sym < clen && sym * tbits < b1; sym++) {
            const uint32_t v = vals[sym];
            const int sh = (int)b1 - (int)(sym * tbits + tbits);      // distance of the symbol's last bit from the word's last bit
            word |= sh >= 0 ? (sh < 32 ? v << sh : 0u) : v >> (-sh);
        }
        out[w] = enc_bswap32(word);
    }
}

cudaError_t configure_bucket_encode(size_t smem_per_cta) {
    return cudaFuncSetAttribute(bucket_mtf_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_per_cta);
}

void launch_bucket_mtf_rle(const EncodeParams &p, int warps_per_cta, size_t smem_per_cta, cudaStream_t s) {
    if (p.n_buckets == 0) return;
    const uint32_t grid = (uint32_t)((p.n_buckets + warps_per_cta - 1) / warps_per_cta);
    bucket_mtf_rle_kernel<<<grid, warps_per_cta * 32, smem_per_cta, s>>>(p);
}

void launch_payload_offsets(const EncodeParams &p, cudaStream_t s) { payload_offsets_kernel<<<1, 1024, 0, s>>>(p); }

void launch_bucket_substitute_pack(const EncodeParams &p, uint64_t batch_first, uint64_t batch_count, const uint8_t *d_ks,
                                   uint32_t ks_stride_bytes, cudaStream_t s) {
    if (batch_count == 0) return;
    const int wpc = 4;
    bucket_substitute_pack_kernel<<<(uint32_t)((batch_count + wpc - 1) / wpc), wpc * 32, 0, s>>>(p, batch_first, batch_count, d_ks,
                                                                                                 ks_stride_bytes);
}

}  // namespace e2fm
