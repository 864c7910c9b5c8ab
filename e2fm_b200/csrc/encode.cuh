// encode.cuh — build-side bucket text encoder (encode.cu): parameters and launchers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace e2fm {

struct EncodeParams {
    const uint32_t *codes;        // [n_rows] bucket codes of the rows of buckets first_bucket .. first_bucket + n_buckets - 1
    uint64_t n_rows, b_size;
    uint64_t first_bucket, n_buckets;
    uint64_t per, total_buckets;  // buckets per super-bucket, buckets of the whole index: which buckets are reversed
    const uint32_t *alpha;        // [n_buckets] bucketAlphaSize
    uint16_t *vals;               // [n_rows] scratch: the plain, then the substituted sequence of each bucket at its row offset
    uint64_t *clen;               // [n_buckets] out: EncodingResult::realLength (0: single-character bucket, no text)
    uint64_t *payload_off;        // [n_buckets + 1] out: byte offset of each bucket's packed text (8-byte aligned slots)
    uint8_t *payload;             // out (may be null: sequence only)
    uint64_t payload_cap;
    uint32_t *sequence;           // [n_rows] optional out: EncodingResult::sequence of each bucket at its row offset
    uint32_t *status;             // ENCODE_ERR_* flags
    uint32_t list_elems;          // per-warp MTF list capacity (elements)
};

constexpr uint32_t ENCODE_ERR_ALPHABET = 1;   // bucketAlphaSize 0 or above the kernel's list capacity
constexpr uint32_t ENCODE_ERR_CODE = 2;       // a row's code is not below its bucket's alphabet size
constexpr uint32_t ENCODE_ERR_BOUNDS = 4;     // key stream too short
constexpr uint32_t ENCODE_ERR_CAPACITY = 8;   // payload buffer too small

cudaError_t configure_bucket_encode(size_t smem_per_cta);
void launch_bucket_mtf_rle(const EncodeParams &p, int warps_per_cta, size_t smem_per_cta, cudaStream_t s);
void launch_payload_offsets(const EncodeParams &p, cudaStream_t s);
void launch_bucket_substitute_pack(const EncodeParams &p, uint64_t batch_first, uint64_t batch_count, const uint8_t *d_ks,
                                   uint32_t ks_stride_bytes, cudaStream_t s);

}  // namespace e2fm
