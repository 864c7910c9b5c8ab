// index_build.cu — one-time device pipeline: raw index image -> flattened search layout.
//
//   K1  salsa20_keystream_kernel   Salsa20/20 key stream of a batch of buckets (one thread per 64-byte block).
//                                  Replaces salsa20.S behind RandomGenerator::populateKeyStreamBuffer
//                                  (RandomGenerator.cpp:125-172) for Bucket::computeBucketKeyStream (Bucket.cpp:266).
//   K2a sb_parse_kernel            SuperBucket::load (SuperBucket.cpp:37-79): alphabet bitset -> decode map.
//   K2  bucket_decode_kernel       one warp per bucket: Bucket::readHeader (Bucket.cpp:549-636), BitSetRLEEncoder::decompress
//                                  (BitSetRLEEncoder.cpp:67), Bucket::readTextUntilTo (Bucket.cpp:288-403: poly-alphabetic
//                                  substitution^-1, RLE0^-1, MTF^-1 with the MTF list in shared memory), marked-row values
//                                  (Bucket.cpp:823-882). Writes row records {code, marked, value} and markTab.row.
//   K3x rank_hist / rank_scan / rank_fill   build the per-(block,symbol) rank sectors and materialise LF per row.
#include "index_build.cuh"

#include "bits.h"
#include "salsa20.h"
#include "seed.cuh"

#include <algorithm>

namespace e2fm {

// ---------------------------------------------------------------- device bit access

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// n (1..32) bits at absolute bit offset `bit` of a big-endian bit stream stored in 4-byte aligned memory.
__device__ __forceinline__ uint32_t dev_bits(const uint8_t *base, uint64_t bit, uint32_t n) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base) + (bit >> 5);
    const uint64_t v = ((uint64_t)bswap32(__ldg(w)) << 32) | bswap32(__ldg(w + 1));
    const uint32_t o = (uint32_t)bit & 31u;
    const uint32_t r = (uint32_t)(v >> (64 - o - n));
    return n == 32 ? r : (r & ((1u << n) - 1));
}
__device__ __forceinline__ uint64_t dev_bits64(const uint8_t *base, uint64_t bit) {
    return ((uint64_t)dev_bits(base, bit, 32) << 32) | dev_bits(base, bit + 32, 32);
}

// ---------------------------------------------------------------- K1: Salsa20 key stream

// One thread per (bucket, block). Output: ks[(bucket - first_bucket) * blocks_per_bucket + block] as 16 raw
// little-endian words (the standard Salsa20 byte stream when viewed as bytes). Pure integer ALU work:
// 320 add + 320 rotate (SHF) + 320 xor per block.
__global__ void __launch_bounds__(256) salsa20_keystream_kernel(const uint32_t k0, const uint32_t k1, const uint32_t k2,
                                                                 const uint32_t k3, const uint32_t k4, const uint32_t k5,
                                                                 const uint32_t k6, const uint32_t k7, uint64_t first_nonce,
                                                                 uint64_t n_nonces, uint32_t blocks_per_nonce,
                                                                 uint64_t first_block, uint4 *__restrict__ ks) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_nonces * blocks_per_nonce) return;
    const uint64_t nonce = first_nonce + t / blocks_per_nonce;
    const uint64_t blk = first_block + t % blocks_per_nonce;
    const uint32_t key[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
    uint32_t o[16];
    salsa20_block(key, nonce, blk, o);
    uint4 *dst = ks + t * 4;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    dst[2] = make_uint4(o[8], o[9], o[10], o[11]);
    dst[3] = make_uint4(o[12], o[13], o[14], o[15]);
}

void launch_keystream(const uint32_t key[8], uint64_t first_nonce, uint64_t n_nonces, uint32_t blocks_per_nonce,
                      uint64_t first_block, uint4 *d_ks, cudaStream_t s) {
    const uint64_t total = n_nonces * blocks_per_nonce;
    if (total == 0) return;
    const uint32_t grid = (uint32_t)((total + 255) / 256);
    salsa20_keystream_kernel<<<grid, 256, 0, s>>>(key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7], first_nonce,
                                                  n_nonces, blocks_per_nonce, first_block, d_ks);
}

// test hook: `count` values of `bits` bits starting at value index `start` of the stream of `nonce`
__global__ void keystream_values_kernel(const uint8_t *__restrict__ ks_bytes, uint64_t first_block, uint32_t bits, uint64_t start,
                                        uint64_t count, uint32_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t bit = (start + i) * bits - first_block * 512;
    out[i] = dev_bits(ks_bytes, bit, bits);
}

void launch_keystream_debug(const uint32_t key[8], uint64_t nonce, uint32_t bits, uint64_t start, uint64_t count, uint32_t *d_out,
                            cudaStream_t s) {
    const uint64_t first_block = start * bits / 512;
    const uint64_t last_block = ((start + count) * bits + 511) / 512 + 1;
    const uint32_t nb = (uint32_t)(last_block - first_block + 1);
    uint4 *d_ks = nullptr;
    cudaMallocAsync(&d_ks, (size_t)nb * 64 + 64, s);
    launch_keystream(key, nonce, 1, nb, first_block, d_ks, s);
    keystream_values_kernel<<<(uint32_t)((count + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint8_t *>(d_ks), first_block, bits,
                                                                            start, count, d_out);
    cudaFreeAsync(d_ks, s);
}

void launch_salsa_debug(const uint32_t key[8], uint64_t nonce, uint64_t first_block, uint64_t n_blocks, uint32_t *d_out_words,
                        cudaStream_t s) {
    launch_keystream(key, nonce, 1, (uint32_t)n_blocks, first_block, reinterpret_cast<uint4 *>(d_out_words), s);
}

// ---------------------------------------------------------------- K2a: super-bucket alphabets

// "1-bit flag, unaligned i64 size, bits" (SuperBucket.cpp:41-57). One warp per super-bucket; the decode map
// (sb code -> compact code, SuperBucket.cpp:62-68) comes from a ballot-free popcount prefix over 32-bit words.
__global__ void __launch_bounds__(128) sb_parse_kernel(BuildParams p) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t sn = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (sn >= p.S) return;
    uint64_t bit = p.sb_start[sn] * 8;
    const uint32_t compressed = dev_bits(p.file, bit, 1);
    const uint64_t sz = dev_bits64(p.file, bit + 1);
    bit += 65;
    if (compressed || sz != p.A || bit + sz > p.file_bytes * 8) {
        if (lane == 0) atomicOr(p.status, BUILD_ERR_ALPHABET);
        return;
    }
    uint32_t running = 0;
    for (uint32_t j0 = 0; j0 < p.A; j0 += 32) {
        const uint32_t nbits = min(32u, p.A - j0);
        uint32_t w = dev_bits(p.file, bit + j0, nbits) << (32 - nbits);   // bit j0+i at position 31-i
        const uint32_t mine = (w >> (31 - lane)) & 1u;
        const uint32_t before = lane == 0 ? 0u : __popc(w >> (32 - lane));
        if (mine) p.sb_decode[sn * p.A + running + before] = (uint16_t)(j0 + lane);
        running += __popc(w);
    }
    if (lane == 0) p.sb_alpha[sn] = running;
}

void launch_sb_parse(const BuildParams &p, cudaStream_t s) {
    const uint32_t grid = (uint32_t)((p.S + 3) / 4);
    sb_parse_kernel<<<grid, 128, 0, s>>>(p);
}

// ---------------------------------------------------------------- K2: bucket decode

// MTF list in shared memory, one warp: move the element at `pos` to the front (ArrayMTFList::moveToFront,
// ArrayMTFList.cpp:32). The shift of list[0..pos) by one slot is done 32 elements at a time from the top
// chunk down, so that no chunk overwrites a slot a later chunk still has to read.
__device__ __forceinline__ uint32_t mtf_move_to_front(uint16_t *list, uint32_t pos, uint32_t lane) {
    const uint32_t e = list[pos];
    for (int t = (int)((pos == 0 ? 0 : pos - 1) >> 5); t >= 0; --t) {
        const uint32_t idx = ((uint32_t)t << 5) + lane;
        uint16_t v = 0;
        if (idx < pos) v = list[idx];
        __syncwarp();
        if (idx < pos) list[idx + 1] = v;
        __syncwarp();
    }
    if (lane == 0) list[0] = (uint16_t)e;
    __syncwarp();
    return e;
}

template <bool kDummy>
__global__ void __launch_bounds__(128) bucket_decode_kernel(BuildParams p, uint64_t first_bucket, uint64_t n_buckets,
                                                             const uint8_t *__restrict__ ks, uint32_t ks_stride_bytes) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uint64_t local = (uint64_t)blockIdx.x * wpc + warp;
    if (local >= n_buckets) return;
    const uint64_t bn = first_bucket + local;
    const uint32_t per_warp = p.list_elems * 4 + p.mark_words * 4;
    uint16_t *list = reinterpret_cast<uint16_t *>(smem_raw + (size_t)warp * per_warp);
    uint16_t *dec = list + p.list_elems;
    uint32_t *markbits = reinterpret_cast<uint32_t *>(dec + p.list_elems);

    const uint64_t file_bits = p.file_bytes * 8;
    const uint64_t sn = bn / p.per;
    const uint32_t sbA = p.sb_alpha[sn];
    const uint64_t row0 = bn * p.b_size;
    const uint32_t length = (uint32_t)min(p.b_size, p.n - row0);
    const bool first = (bn % p.per) == 0;
    const bool odd = (bn % 2 == 0) && !first && (bn != p.B - 1);   // EFMIndex.cpp:898-901

    // ---- header (every lane parses the same fields: uniform, no broadcast needed) ----
    uint64_t bit = p.b_start[bn] * 8;
    const uint64_t m_start = dev_bits64(p.file, bit), t_start = dev_bits64(p.file, bit + 64);
    bit += 128;
    if (!first && !odd) {
        // skip this bucket's own counter table (Bucket.cpp:570-573): u8 head bits + sbAlpha integerDecode fields.
        // The rank structure is rebuilt from the decoded text, so the values are not needed, only their length.
        const uint32_t h = dev_bits(p.file, bit, 8);
        bit += 8;
        for (uint32_t j = 0; j < sbA; j++) {
            if (bit + 64 > file_bits) break;
            const uint32_t i = h ? dev_bits(p.file, bit, h) : 0;
            bit += h;
            if (i == 2) bit += 1;
            else if (i > 2) bit += i - 1;
        }
    }
    bool bad = bit + 65 + sbA > file_bits;
    uint32_t alpha = 0;
    if (!bad) {
        // bucket alphabet bitset (Bucket.cpp:581-606) -> dec[bucket code] = compact code
        const uint32_t compressed = dev_bits(p.file, bit, 1);
        const uint64_t sz = dev_bits64(p.file, bit + 1);
        bit += 65;
        if (compressed || sz != sbA || sbA > p.list_elems) {
            if (lane == 0) atomicOr(p.status, BUILD_ERR_ALPHABET);
            return;
        }
        const uint16_t *sbd = p.sb_decode + sn * p.A;
        for (uint32_t j0 = 0; j0 < sbA; j0 += 32) {
            const uint32_t nbits = min(32u, sbA - j0);
            const uint32_t w = dev_bits(p.file, bit + j0, nbits) << (32 - nbits);
            const uint32_t mine = (w >> (31 - lane)) & 1u;
            const uint32_t before = lane == 0 ? 0u : __popc(w >> (32 - lane));
            if (mine) dec[alpha + before] = sbd[j0 + lane];
            alpha += __popc(w);
        }
        bit += sbA;
    }
    // ---- marked-rows bitmap: BitSetRLEEncoder::decompress (BitSetRLEEncoder.cpp:67-135), truncated to `length` ----
    for (uint32_t w = lane; w < p.mark_words; w += 32) markbits[w] = 0;
    __syncwarp();
    if (!bad && p.mrp > 0) {
        const uint64_t csz = dev_bits64(p.file, bit);
        bit += 64;
        if (bit + csz > file_bits) bad = true;
        else {
            uint64_t outp = 0, N = 0;
            uint32_t x = 0;
            bool pending = false;
            for (uint64_t i0 = 0; i0 < csz; i0 += 32) {
                const uint32_t nb = (uint32_t)min((uint64_t)32, csz - i0);
                const uint32_t w = dev_bits(p.file, bit + i0, nb) << (32 - nb);
                for (uint32_t q = 0; q < nb; q++) {
                    const uint32_t b = (w >> (31 - q)) & 1u;
                    if (!pending) {
                        if (b == 0) { N = 2 * N; x++; } else pending = true;
                    } else {
                        if (b == 0) { N = 2 * N + 1; x++; }
                        else {
                            if (x > 0) outp += (x < 40 ? (1ull << x) : (1ull << 40)) + N - 1;
                            if (outp < length && lane == 0) markbits[outp >> 5] |= 1u << (outp & 31);
                            outp++; N = 0; x = 0;
                        }
                        pending = false;
                    }
                }
            }
            if (x > 0) {   // left-over digits are flushed and followed by a set bit (:116-133)
                outp += (x < 40 ? (1ull << x) : (1ull << 40)) + N - 1;
                if (outp < length && lane == 0) markbits[outp >> 5] |= 1u << (outp & 31);
            }
            bit += csz;
        }
    }
    __syncwarp();
    if (bad) {
        if (lane == 0) atomicOr(p.status, BUILD_ERR_BOUNDS);
        return;
    }
    bit = (bit + 7) & ~7ull;                                   // gotoNextByteStart (Bucket.cpp:621)
    const uint64_t clen = dev_bits64(p.file, bit);             // junk when alpha == 1 (Bucket.cpp:168-169)
    const uint32_t tbits = int_log2(alpha);                    // bitsPerTextCharacter (Bucket.cpp:628)
    const uint32_t kbits = int_log2((uint64_t)alpha + 1);      // key width (RandomGenerator.cpp:136)
    const uint64_t text_bit = t_start * 8;
    uint64_t *rec = p.rec + row0;

    // ---- text: poly^-1 -> RLE0^-1 -> MTF^-1 (Bucket.cpp:288-403) ----
    uint32_t produced = 0;
    if (alpha <= 1) {                                          // expandSingleCharacterUntilTo (Bucket.cpp:414)
        const uint64_t v = (uint64_t)dec[0] << 32;
        for (uint32_t q = lane; q < length; q += 32) rec[q] = v;
        produced = length;
    } else {
        if (clen > length || text_bit + clen * tbits > file_bits || (clen * kbits + 7) / 8 + 8 > ks_stride_bytes) {
            if (lane == 0) atomicOr(p.status, BUILD_ERR_BOUNDS);
            return;
        }
        for (uint32_t q = lane; q < alpha; q += 32) list[q] = (uint16_t)q;   // ArrayMTFList ctor
        __syncwarp();
        const uint8_t *myks = ks + local * ks_stride_bytes;
        const uint32_t mod = alpha + 1;
        uint32_t pp = 0;                                       // physical offset of the next symbol
        uint32_t x = 0;                                        // pending RLE0 digits: count and value
        uint64_t N = 0;
        for (uint64_t base = 0; base < clen && pp < length; base += 32) {
            const uint64_t i = base + lane;
            uint32_t c = 0;
            if (i < clen) {
                const int32_t v = (int32_t)dev_bits(p.file, text_bit + i * tbits, tbits);
                const int32_t key = (int32_t)dev_bits(myks, i * kbits, kbits);
                int32_t d = (v - key) % (int32_t)mod;          // invertPoly (Bucket.cpp:406)
                if (d < 0) d += (int32_t)mod;
                c = (uint32_t)d;
            }
            const uint32_t cnt = (uint32_t)min((uint64_t)32, clen - base);
            for (uint32_t j = 0; j < cnt && pp < length; j++) {
                const uint32_t cj = __shfl_sync(0xFFFFFFFFu, c, j);
                if (cj <= 1) { N = 2 * N + cj; x++; continue; }
                if (x > 0) {                                   // run of the front symbol: 2^x + N - 1 copies (:377)
                    uint64_t m = (x < 40 ? (1ull << x) : (1ull << 40)) + N - 1;
                    if (m > length - pp) m = length - pp;
                    const uint64_t v = (uint64_t)dec[list[0]] << 32;
                    for (uint32_t q = lane; q < (uint32_t)m; q += 32) rec[odd ? length - 1 - (pp + q) : pp + q] = v;
                    pp += (uint32_t)m; x = 0; N = 0;
                    if (pp >= length) break;
                }
                const uint32_t pos = cj - 1;                   // RLE0^-1 of an MTF rank
                const uint32_t e = mtf_move_to_front(list, pos, lane);
                if (lane == 0) rec[odd ? length - 1 - pp : pp] = (uint64_t)dec[e] << 32;
                pp++;
            }
        }
        if (x > 0 && pp < length) {                            // digits at the very end of the text
            uint64_t m = (x < 40 ? (1ull << x) : (1ull << 40)) + N - 1;
            if (m > length - pp) m = length - pp;
            const uint64_t v = (uint64_t)dec[list[0]] << 32;
            for (uint32_t q = lane; q < (uint32_t)m; q += 32) rec[odd ? length - 1 - (pp + q) : pp + q] = v;
            pp += (uint32_t)m;
        }
        produced = pp;
    }
    if (produced != length) {
        if (lane == 0) atomicOr(p.status, BUILD_ERR_LENGTH);
        return;
    }
    __syncwarp();
    // ---- marked values (Bucket.cpp:823-882): the j-th set bit owns the j-th val_bits field at m_start ----
    if (p.mrp > 0) {
        const uint64_t mark_bit = m_start * 8;
        uint32_t running = 0;
        for (uint32_t w0 = 0; w0 < p.mark_words; w0 += 32) {
            const uint32_t wi = w0 + lane;
            uint32_t w = wi < p.mark_words ? markbits[wi] : 0;
            const uint32_t pc = __popc(w);
            uint32_t incl = pc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            uint32_t j = running + incl - pc;
            running += __shfl_sync(0xFFFFFFFFu, incl, 31);
            while (w) {
                const uint32_t b = __ffs(w) - 1;
                w &= w - 1;
                const uint32_t o = (wi << 5) + b;
                const uint64_t vb = mark_bit + (uint64_t)j * p.val_bits;
                if (vb + p.val_bits > file_bits) { atomicOr(p.status, BUILD_ERR_BOUNDS); break; }
                const uint32_t v = dev_bits(p.file, vb, p.val_bits);
                if (v >= p.mrn) { atomicOr(p.status, BUILD_ERR_MARKS); }
                else {
                    rec[o] = rec[o] | REC_MARKED | v;
                    p.mark_tab[v].x = (uint32_t)(row0 + o);
                }
                j++;
            }
        }
    }
}

void launch_bucket_decode(const BuildParams &p, uint64_t first_bucket, uint64_t n_buckets, const uint8_t *d_ks,
                          uint32_t ks_stride_bytes, int warps_per_cta, size_t smem_per_cta, cudaStream_t s) {
    if (n_buckets == 0) return;
    const uint32_t grid = (uint32_t)((n_buckets + warps_per_cta - 1) / warps_per_cta);
    bucket_decode_kernel<true><<<grid, warps_per_cta * 32, smem_per_cta, s>>>(p, first_bucket, n_buckets, d_ks, ks_stride_bytes);
}

cudaError_t configure_bucket_decode(size_t smem_per_cta) {
    return cudaFuncSetAttribute(bucket_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_per_cta);
}

// ---------------------------------------------------------------- rank sectors + LF

// one CTA per rank block: histogram of the block's codes into the sectors' base fields
__global__ void __launch_bounds__(256) rank_hist_kernel(RankBuildParams p) {
    extern __shared__ uint32_t hist[];
    const uint32_t blk = blockIdx.x;
    for (uint32_t c = threadIdx.x; c < p.A; c += blockDim.x) hist[c] = 0;
    __syncthreads();
    const uint64_t r0 = (uint64_t)blk << p.blk_shift;
    const uint32_t rows = (uint32_t)min((uint64_t)1 << p.blk_shift, p.n - r0);
    for (uint32_t i = threadIdx.x; i < rows; i += blockDim.x) atomicAdd(&hist[rec_code(p.rec[r0 + i])], 1u);
    __syncthreads();
    unsigned long long need = 0;
    for (uint32_t c = threadIdx.x; c < p.A; c += blockDim.x) {
        const uint32_t h = hist[c];
        reinterpret_cast<uint32_t *>(p.occ + ((size_t)blk * p.A + c) * 2)[0] = h;
        if (h > SECTOR_SLOTS) need += h - SECTOR_INLINE_OVF + 1;
    }
    if (need) atomicAdd(p.ovf_cursor, need);
}

void launch_rank_hist(const RankBuildParams &p, cudaStream_t s) {
    rank_hist_kernel<<<p.n_blk, 256, (size_t)p.A * 4, s>>>(p);
}
cudaError_t configure_rank_hist(size_t smem) {
    return cudaFuncSetAttribute(rank_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// column-wise exclusive scan of the per-block counts (threads run along the symbol axis so that every
// step of every thread touches a different sector of the same block row).
__global__ void __launch_bounds__(256) rank_scan_totals_kernel(RankBuildParams p, uint32_t *__restrict__ totals, uint32_t chunk_blocks) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x, chunk = blockIdx.y;
    if (c >= p.A) return;
    const uint32_t b0 = chunk * chunk_blocks, b1 = min(p.n_blk, b0 + chunk_blocks);
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; b++) sum += reinterpret_cast<const uint32_t *>(p.occ + ((size_t)b * p.A + c) * 2)[0];
    totals[(size_t)chunk * p.A + c] = sum;
}
__global__ void __launch_bounds__(256) rank_scan_chunks_kernel(RankBuildParams p, uint32_t *__restrict__ totals, uint32_t n_chunks) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.A) return;
    uint32_t run = (uint32_t)p.C[c];
    for (uint32_t ch = 0; ch < n_chunks; ch++) {
        const uint32_t t = totals[(size_t)ch * p.A + c];
        totals[(size_t)ch * p.A + c] = run;
        run += t;
    }
    if (run != (uint32_t)p.C[c + 1]) atomicOr(p.status, BUILD_ERR_LENGTH);   // decoded text disagrees with the header's C[]
}
__global__ void __launch_bounds__(256) rank_scan_apply_kernel(RankBuildParams p, const uint32_t *__restrict__ totals, uint32_t chunk_blocks) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x, chunk = blockIdx.y;
    if (c >= p.A) return;
    const uint32_t b0 = chunk * chunk_blocks, b1 = min(p.n_blk, b0 + chunk_blocks);
    uint32_t run = totals[(size_t)chunk * p.A + c];
    for (uint32_t b = b0; b < b1; b++) {
        uint32_t *q = reinterpret_cast<uint32_t *>(p.occ + ((size_t)b * p.A + c) * 2);
        const uint32_t t = *q;
        *q = run;
        run += t;
    }
}

void launch_rank_scan(const RankBuildParams &p, uint32_t *chunk_totals, uint32_t n_chunks, uint32_t chunk_blocks, cudaStream_t s) {
    dim3 grid((p.A + 255) / 256, n_chunks);
    rank_scan_totals_kernel<<<grid, 256, 0, s>>>(p, chunk_totals, chunk_blocks);
    rank_scan_chunks_kernel<<<(p.A + 255) / 256, 256, 0, s>>>(p, chunk_totals, n_chunks);
    rank_scan_apply_kernel<<<grid, 256, 0, s>>>(p, chunk_totals, chunk_blocks);
}

// one CTA per rank block, one thread per symbol: collect the symbol's in-block offsets (ascending), write its
// sector and the LF value of each of its rows (LF = base + j for the j-th occurrence in the block).
__global__ void __launch_bounds__(256) rank_fill_kernel(RankBuildParams p) {
    extern __shared__ __align__(16) uint16_t codes[];
    const uint32_t blk = blockIdx.x;
    const uint32_t BLK = 1u << p.blk_shift;
    const uint64_t r0 = (uint64_t)blk << p.blk_shift;
    const uint32_t rows = (uint32_t)min((uint64_t)BLK, p.n - r0);
    for (uint32_t i = threadIdx.x; i < BLK; i += blockDim.x) codes[i] = i < rows ? (uint16_t)rec_code(p.rec[r0 + i]) : (uint16_t)0xFFFF;
    __syncthreads();
    const uint4 *cv = reinterpret_cast<const uint4 *>(codes);
    for (uint32_t c = threadIdx.x; c < p.A; c += blockDim.x) {
        uint4 *sec = p.occ + ((size_t)blk * p.A + c) * 2;
        const uint32_t base = reinterpret_cast<const uint32_t *>(sec)[0];
        const uint32_t cc = c | (c << 16);
        uint16_t off[SECTOR_SLOTS];
#pragma unroll
        for (int i = 0; i < SECTOR_SLOTS; i++) off[i] = (uint16_t)SLOT_PAD;
        uint32_t cnt = 0;
        for (uint32_t v = 0; v < BLK / 8; v++) {
            const uint4 w = cv[v];
            const uint32_t m0 = __vcmpeq2(w.x, cc), m1 = __vcmpeq2(w.y, cc), m2 = __vcmpeq2(w.z, cc), m3 = __vcmpeq2(w.w, cc);
            if ((m0 | m1 | m2 | m3) == 0) continue;
            const uint32_t ms[4] = {m0, m1, m2, m3};
#pragma unroll
            for (int h = 0; h < 8; h++) {
                if ((ms[h >> 1] >> ((h & 1) * 16)) & 1u) {
                    const uint32_t o = v * 8 + h;
                    // LF of this row
                    const uint64_t r = p.rec[r0 + o];
                    const uint32_t lf = base + cnt;
                    if (rec_marked(r)) p.mark_tab[rec_low(r)].y = lf;
                    else p.rec[r0 + o] = r | lf;
                    if (cnt < SECTOR_SLOTS) {
#pragma unroll
                        for (int i = 0; i < SECTOR_SLOTS; i++) if (i == (int)cnt) off[i] = (uint16_t)o;
                    }
                    cnt++;
                }
            }
        }
        uint32_t w[8];
        w[0] = base;
        if (cnt <= SECTOR_SLOTS) {
#pragma unroll
            for (int i = 0; i < 7; i++) w[1 + i] = (uint32_t)off[2 * i] | ((uint32_t)off[2 * i + 1] << 16);
        } else {
            // overflow: 12 inline offsets (bit 15 of the first flags it) + pool index; the pool gets the rest + 0xFFFF
            const uint32_t extra = cnt - SECTOR_INLINE_OVF + 1;
            const unsigned long long at = atomicAdd(p.ovf_cursor + 1, (unsigned long long)extra);
#pragma unroll
            for (int i = 0; i < 6; i++) w[1 + i] = (uint32_t)off[2 * i] | ((uint32_t)off[2 * i + 1] << 16);
            w[1] |= 0x80000000u;   // overflow flag: bit 15 of slot 1 (the high half, so a SWAR borrow cannot leak)
            w[7] = (uint32_t)at;
            uint32_t seen = 0, q = 0;
            for (uint32_t i = 0; i < rows; i++) {
                if (codes[i] == c) {
                    if (seen >= SECTOR_INLINE_OVF) p.ovf[at + q++] = (uint16_t)i;
                    seen++;
                }
            }
            p.ovf[at + q] = 0xFFFF;
        }
        sec[0] = make_uint4(w[0], w[1], w[2], w[3]);
        sec[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

void launch_rank_fill(const RankBuildParams &p, cudaStream_t s) {
    rank_fill_kernel<<<p.n_blk, 256, (size_t)2 << p.blk_shift, s>>>(p);
}
cudaError_t configure_rank_fill(size_t smem) {
    return cudaFuncSetAttribute(rank_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

// ---------------------------------------------------------------- dense locate tables

__global__ void __launch_bounds__(256) dense_build_kernel(const uint64_t *__restrict__ rec, const uint2 *__restrict__ mark_tab, uint64_t mrn,
                                                           uint32_t rate, uint64_t n, uint32_t *__restrict__ sa, uint16_t *__restrict__ text,
                                                           uint32_t *status) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= mrn) return;
    const uint2 mt = mark_tab[v];
    if (mt.x >= n) { atomicOr(status, BUILD_ERR_MARKS); return; }   // a marked value without a row
    uint64_t pos = v * rate;                                        // text position of the marked row's suffix
    uint32_t row = mt.x;
    uint64_t r = rec[row];
    sa[row] = (uint32_t)pos;
    text[pos == 0 ? n - 1 : pos - 1] = (uint16_t)rec_code(r);       // BWT[row] precedes the suffix in the text
    row = mt.y;                                                     // LF of the marked row
    for (uint32_t step = 0; step <= rate; step++) {
        r = rec[row];
        if (rec_marked(r)) return;                                  // the previous marked row: its own thread continues
        pos = pos == 0 ? n - 1 : pos - 1;
        sa[row] = (uint32_t)pos;
        text[pos == 0 ? n - 1 : pos - 1] = (uint16_t)rec_code(r);
        row = rec_low(r);
    }
    atomicOr(status, BUILD_ERR_MARKS);                              // no marked row within `rate` steps: corrupt index
}

void launch_dense_build(const uint64_t *rec, const uint2 *mark_tab, uint64_t mrn, uint32_t rate, uint64_t n, uint32_t *sa,
                        uint16_t *text, uint32_t *status, cudaStream_t s) {
    if (mrn == 0) return;
    dense_build_kernel<<<(uint32_t)((mrn + 255) / 256), 256, 0, s>>>(rec, mark_tab, mrn, rate, n, sa, text, status);
}

// bwt16 has 16 entries of padding (firstcheck_kernel reads aligned groups of 8 rows and looks every code up in a table before it
// masks the rows outside the interval): they are zeroed, a code must never be garbage
__global__ void __launch_bounds__(256) bwt16_build_kernel(const uint64_t *__restrict__ rec, uint64_t n, uint16_t *__restrict__ bwt16) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bwt16[i] = (uint16_t)rec_code(rec[i]);
    else if (i < n + 16) bwt16[i] = 0;
}
void launch_bwt16_build(const uint64_t *rec, uint64_t n, uint16_t *bwt16, cudaStream_t s) {
    if (n) bwt16_build_kernel<<<(uint32_t)((n + 16 + 255) / 256), 256, 0, s>>>(rec, n, bwt16);
}

// Verification records (device_index.cuh): one 32-byte sector per row with everything seed_verify_rec_kernel needs to finish
// EFMIndex::exactMatch for a row the seed table produced — the row's text position and the super-characters around it — so that
// a row task costs ONE gather instead of sa[row] + one or two sectors of text[].
//   w0      sa[row]
//   w1..w5  text[sa - 1], text[sa - 2], .., text[sa - 10]   (indexes mod n; 16 bits each, the nearer one in the low half)
//   w6, w7  text[sa + 2], text[sa + 3], text[sa + 4], 0: the super-character behind a seed of 2, 3 or 4 (the '?'-suffixed last one);
//           0xFFFF beyond the end of the text
__global__ void __launch_bounds__(256) vrec_build_kernel(uint64_t n, const uint32_t *__restrict__ sa, const uint16_t *__restrict__ text,
                                                          uint4 *__restrict__ vrec) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const uint32_t pos = sa[row], nn = (uint32_t)n;
    uint32_t w[8];
    w[0] = pos;
#pragma unroll
    for (int q = 0; q < 5; q++) {
        const uint32_t r0 = 2 * q + 1, r1 = 2 * q + 2;
        const uint32_t a = pos >= r0 ? pos - r0 : pos + (nn - r0), b = pos >= r1 ? pos - r1 : pos + (nn - r1);
        w[1 + q] = (uint32_t)__ldg(text + a) | ((uint32_t)__ldg(text + b) << 16);
    }
    uint32_t right[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const uint64_t ph = (uint64_t)pos + 2 + q;
        right[q] = ph < n ? (uint32_t)__ldg(text + ph) : 0xFFFFu;
    }
    w[6] = right[0] | (right[1] << 16);
    w[7] = right[2];
    vrec[2 * row] = make_uint4(w[0], w[1], w[2], w[3]);
    vrec[2 * row + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}
void launch_vrec_build(uint64_t n, const uint32_t *sa, const uint16_t *text, uint4 *vrec, cudaStream_t s) {
    if (n) vrec_build_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(n, sa, text, vrec);
}

__global__ void __launch_bounds__(256) pair_build_kernel(DevIndex ix, uint2 *__restrict__ pair_tab) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ix.A * ix.A) return;
    const uint32_t a = t / ix.A, b = t - a * ix.A;
    const uint32_t s = (uint32_t)ix.C[a], e = (uint32_t)ix.C[a + 1];
    uint2 out = make_uint2(0u, 0u);
    if (e > s) {
        const uint32_t ns = s == 0 ? (uint32_t)ix.C[b] : rank_c(ix, b, s - 1);
        const uint32_t ne = rank_c(ix, b, e - 1);
        out = make_uint2(ns, ne > ns ? ne : ns);
    }
    pair_tab[t] = out;
}
void launch_pair_build(const DevIndex &ix, uint2 *pair_tab, cudaStream_t s) {
    const uint32_t n = ix.A * ix.A;
    if (n) pair_build_kernel<<<(n + 255) / 256, 256, 0, s>>>(ix, pair_tab);
}

// ---------------------------------------------------------------- seed table (seed.cuh)

// every 16-byte entry: base = n (no row at or beyond the group's first seed), all counters 0
__global__ void __launch_bounds__(256) seed_fill_kernel(uint4 *__restrict__ tab, uint64_t n_sectors, uint32_t n_rows) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sectors; i += (uint64_t)gridDim.x * blockDim.x)
        tab[i] = make_uint4(n_rows, 0u, 0u, 0u);
}

// the first j super-characters of the rotation that row `row` stands for: all but the last packed in `hi`, the last in `cl`
__device__ __forceinline__ void seed_of_row(const SeedBuildParams &p, uint32_t row, uint32_t &hi, uint32_t &cl) {
    uint64_t pos = p.sa[row];
    uint32_t h = 0;
    for (uint32_t i = 0; i + 1 < p.j; i++) {
        h = h * p.A + p.text[pos];
        pos = pos + 1 == p.n ? 0 : pos + 1;
    }
    hi = h;
    cl = p.text[pos];
}

// One thread per row. Rows are sorted by rotation, so rows with the same seed are consecutive and seeds never decrease
// with the row number (checked: a violation sets BUILD_ERR_SEED and the table is dropped). The first row of every run
// writes the run's length into the seed's counter and the base of every group that begins between the previous row's
// seed and its own.
__global__ void __launch_bounds__(256) seed_build_kernel(SeedBuildParams p) {
    const uint64_t r64 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r64 >= p.n) return;
    const uint32_t r = (uint32_t)r64;
    uint32_t hi, cl;
    seed_of_row(p, r, hi, cl);
    const uint64_t key = (uint64_t)hi * p.A + cl;
    const uint64_t g = (uint64_t)hi * p.gph + cl / SEED_GROUP;
    uint64_t g_first = 0;            // first group whose base this row sets
    bool start = true;
    if (r > 0) {
        uint32_t hp, cp;
        seed_of_row(p, r - 1, hp, cp);
        const uint64_t keyp = (uint64_t)hp * p.A + cp;
        if (keyp > key) atomicOr(p.status, BUILD_ERR_SEED);
        start = keyp != key;
        g_first = (uint64_t)hp * p.gph + cp / SEED_GROUP + 1;
    }
    if (!start) return;
    uint32_t cnt = 1;
    for (uint64_t rr = r64 + 1; cnt < SEED_SAT && rr < p.n; rr++) {
        uint32_t h2, c2;
        seed_of_row(p, (uint32_t)rr, h2, c2);
        if ((uint64_t)h2 * p.A + c2 != key) break;
        cnt++;
    }
    uint32_t *words = reinterpret_cast<uint32_t *>(p.tab);
    const uint32_t slot = cl % SEED_GROUP;
    atomicOr(words + g * 4 + 1 + (slot >> 3), cnt << ((slot & 7u) * 4u));
    for (uint64_t gg = g_first; gg <= g; gg++) words[gg * 4] = r;
}

// last pass: bit 31 of an entry's base flags a saturated counter (rows are below 2^31, host_format.cpp), so that seed_decode
// needs no per-counter test
__global__ void __launch_bounds__(256) seed_flag_kernel(uint4 *__restrict__ tab, uint64_t n_sectors) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sectors; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 a = tab[i];
        const uint32_t w[3] = {a.y, a.z, a.w};
        uint32_t sat = 0;
#pragma unroll
        for (int q = 0; q < 3; q++) sat |= w[q] & (w[q] >> 1) & (w[q] >> 2) & (w[q] >> 3);
        if (sat & 0x11111111u) reinterpret_cast<uint32_t *>(tab)[i * 4] = a.x | 0x80000000u;
    }
}

void launch_seed_build(const SeedBuildParams &p, uint64_t n_sectors, cudaStream_t s) {
    if (p.n == 0 || n_sectors == 0) return;
    const uint32_t grid = (uint32_t)std::min<uint64_t>((n_sectors + 255) / 256, 148 * 32);
    seed_fill_kernel<<<grid, 256, 0, s>>>(p.tab, n_sectors, (uint32_t)p.n);
    seed_build_kernel<<<(uint32_t)((p.n + 255) / 256), 256, 0, s>>>(p);
    seed_flag_kernel<<<grid, 256, 0, s>>>(p.tab, n_sectors);
}

}  // namespace e2fm
