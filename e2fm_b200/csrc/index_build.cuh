// index_build.cuh — one-time device pipeline that turns the raw (encrypted, compressed) index image
// into the flattened search layout of device_index.cuh.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "device_index.cuh"

namespace e2fm {

struct BuildParams {
    const uint8_t *file;        // device copy of the index image, 16-byte aligned, zero padded by >= 16 bytes
    uint64_t file_bytes;
    const uint64_t *b_start;    // [B]  byte offsets of buckets
    const uint64_t *sb_start;   // [S]
    uint64_t n, B, S, b_size, per;   // per = buckets per super-bucket
    uint64_t mrn;
    uint32_t A, mrp, rate, val_bits;
    uint32_t poly_key[8];
    uint16_t *sb_decode;        // [S*A] super-bucket code -> compact code
    uint32_t *sb_alpha;         // [S]
    uint64_t *rec;              // [n]
    uint2 *mark_tab;            // [mrn]
    uint32_t *status;           // bit flags, see BUILD_ERR_*
    uint32_t ks_bytes;          // per-warp shared key-stream bytes
    uint32_t list_elems;        // per-warp MTF list / decode map capacity (elements)
    uint32_t mark_words;        // per-warp marked-bitmap words
};

constexpr uint32_t BUILD_ERR_BOUNDS = 1;     // a field points outside the file
constexpr uint32_t BUILD_ERR_LENGTH = 2;     // decoded bucket length != expected (wrong key / corrupt text)
constexpr uint32_t BUILD_ERR_ALPHABET = 4;   // bitset size mismatch
constexpr uint32_t BUILD_ERR_MARKS = 8;      // marked value out of range / duplicate
constexpr uint32_t BUILD_ERR_SEED = 16;      // rows are not sorted by their first super-characters: the seed table is dropped

struct SeedBuildParams {
    uint64_t n;
    uint32_t A, j, gph;         // j super-characters per seed, gph sectors per value of the first j-1
    const uint32_t *sa;
    const uint16_t *text;
    uint4 *tab;
    uint32_t *status;
};

struct RankBuildParams {
    uint64_t n;
    uint32_t A, blk_shift, n_blk;
    uint64_t *rec;
    uint4 *occ;
    uint16_t *ovf;
    unsigned long long *ovf_cursor;   // [0] halfwords needed/used in the overflow pool
    const uint64_t *C;
    uint2 *mark_tab;
    uint64_t mrn;
    uint32_t *status;
};

void launch_sb_parse(const BuildParams &p, cudaStream_t s);
// K1 for a batch of buckets: ks[(nonce - first_nonce) * blocks_per_nonce + block] = Salsa20 block `first_block + block` of `nonce`
void launch_keystream(const uint32_t key[8], uint64_t first_nonce, uint64_t n_nonces, uint32_t blocks_per_nonce, uint64_t first_block,
                      uint4 *d_ks, cudaStream_t s);
void launch_bucket_decode(const BuildParams &p, uint64_t first_bucket, uint64_t n_buckets, const uint8_t *d_ks, uint32_t ks_stride_bytes,
                          int warps_per_cta, size_t smem_per_cta, cudaStream_t s);
cudaError_t configure_bucket_decode(size_t smem_per_cta);
cudaError_t configure_rank_hist(size_t smem);
cudaError_t configure_rank_fill(size_t smem);
void launch_rank_hist(const RankBuildParams &p, cudaStream_t s);
void launch_rank_scan(const RankBuildParams &p, uint32_t *chunk_totals, uint32_t n_chunks, uint32_t chunk_blocks, cudaStream_t s);
void launch_rank_fill(const RankBuildParams &p, cudaStream_t s);

// dense locate tables: one thread per marked row walks LF until the next marked row (<= rate steps), writing
// sa[row] = text position and text[position-1] = BWT[row] on the way (EFMIndex::getRowPosition, EFMIndex.cpp:2468, in bulk)
void launch_dense_build(const uint64_t *rec, const uint2 *mark_tab, uint64_t mrn, uint32_t rate, uint64_t n, uint32_t *sa,
                        uint16_t *text, uint32_t *status, cudaStream_t s);

// bwt16[row] = compact code of the row (from rec)
void launch_bwt16_build(const uint64_t *rec, uint64_t n, uint16_t *bwt16, cudaStream_t s);
// pair_tab[a*A + b] = {s,e}: [C[a], C[a+1]) after one backward step with b
void launch_pair_build(const DevIndex &ix, uint2 *pair_tab, cudaStream_t s);

// verification records: 32 bytes per row (sa[row], the 10 super-characters before that position, the ones 2, 3 and 4 behind it)
void launch_vrec_build(uint64_t n, const uint32_t *sa, const uint16_t *text, uint4 *vrec, cudaStream_t s);

// seed table (seed.cuh) from the dense tables: fill + one thread per row
void launch_seed_build(const SeedBuildParams &p, uint64_t n_sectors, cudaStream_t s);

// standalone K1 test hook
void launch_keystream_debug(const uint32_t key[8], uint64_t nonce, uint32_t bits, uint64_t start, uint64_t count,
                            uint32_t *d_out, cudaStream_t s);
void launch_salsa_debug(const uint32_t key[8], uint64_t nonce, uint64_t first_block, uint64_t n_blocks, uint32_t *d_out_words,
                        cudaStream_t s);

}  // namespace e2fm
