/* e2fm_b200.h — C ABI of the B200-native E2FM search path (libe2fm_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no exceptions.
 * Every entry point names the reference interface it replaces (file:line into montecuollo/E2FM).
 * The reference has no FFI of its own; INTEGRATION.md shows the few lines a maintainer adds to
 * EFMIndex / EFMCollection to call these instead of the CPU path.
 *
 * All functions return E2FM_OK (0) or a negative error class; e2fm_last_error() gives the text
 * (thread-local). There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef E2FM_B200_H_
#define E2FM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E2FM_OK 0
#define E2FM_ERR_CUDA (-1)      /* CUDA runtime error / no device */
#define E2FM_ERR_FORMAT (-2)    /* not an E2FM index, corrupt file, or wrong key */
#define E2FM_ERR_ARG (-3)       /* bad argument */
#define E2FM_ERR_UNSUPPORTED (-4) /* e.g. locate on an index without marked rows (EFMIndex.cpp:2305) */
#define E2FM_ERR_OOM (-5)
#define E2FM_ERR_NCCL (-6)      /* NCCL error / libnccl not found (e2fm_shardset_* only) */

#define E2FM_MODE_COUNT 0        /* EFMIndex::exactMatch + countOccurrences           (EFMIndex.cpp:1694,2273) */
#define E2FM_MODE_LOCATE 1       /* ... + locateOccurrences + (sequence, offset) map   (EFMIndex.cpp:2281, EFMCollection.cpp:115) */

typedef struct e2fm_index e2fm_index;     /* device-resident flattened index (one GPU) */
typedef struct e2fm_result e2fm_result;   /* device-resident result of one batch; may be fetched until its index is closed, freed at any time */
typedef struct e2fm_shardset e2fm_shardset; /* one rank's shard of a collection sharded by sequence + the NCCL communicator */

typedef struct e2fm_info {
    uint64_t text_length;        /* n, super-text length (EFMIndex.h:372) */
    uint64_t bucket_size;        /* EFMIndex::getBucketSize      */
    uint64_t superbucket_size;   /* EFMIndex::getSuperBucketSize */
    uint64_t n_buckets;          /* EFMIndex::getBucketsNumber   */
    uint64_t n_superbuckets;     /* EFMIndex::getSuperBucketsNumber */
    uint64_t n_sequences;        /* FMCollection::getSize        */
    uint64_t initial_n;          /* initialNCharacters           */
    uint64_t device_bytes;       /* bytes of HBM held by the index */
    uint32_t k;                  /* EFMIndex::getSuperAlphabetOrder */
    uint32_t n_symbols;          /* original alphabet size       */
    uint32_t compact_alphabet;   /* compactAlphabetSize          */
    uint32_t marked_rows_percentage; /* EFMIndex::getMarkedRowsPercentage */
    uint32_t marking_rate;
    uint32_t rank_block;         /* rows per rank block of the device layout */
    uint32_t quirks;             /* bit0: floor(n/rate) is a power of two (reference locate is corrupt, EFMIndex.cpp:399 vs :1280)
                                    bit1: n % rate == 0 (reference reads one junk marked-row entry, EFMIndex.cpp:1278) */
    uint32_t dense;              /* 0: no dense locate tables (HBM short); 1: built and in use; 2: built, switched off (E2FM_OPT_DENSE) */
    uint32_t seed_chars;         /* super-characters per entry of the seed table (0: not built) */
    uint32_t seed_table_mb;      /* its size in MiB (included in device_bytes) */
    uint32_t vrec_mb;            /* MiB of verification records (32 bytes per row: what the seed search reads per candidate row); 0: not built */
    uint32_t reserved;
    float load_ms[8];            /* one-time costs: [0] host parse, [1] H2D, [2] superbucket parse, [3] bucket decrypt+decode,
                                    [4] histogram, [5] scan, [6] rank sectors + LF fill, [7] total device */
} e2fm_info;

/* Replaces EFMIndex::load / EFMCollection::load (EFMIndex.cpp:1536, EFMCollection.cpp:333) + the lazy
 * SuperBucket::load / Bucket::readHeader / Bucket::readTextUntilTo (SuperBucket.cpp:37, Bucket.cpp:549,288):
 * `file` is the whole .efm index image (what MemoryBitReader::loadDataFromFile returns, MemoryBitReader.cpp:93),
 * `key64` the 64-byte key (EncryptionManager::initSALSA20Keys, EncryptionManager.cpp:47).
 * Header secrets are derived on the host; every bucket is decrypted + decoded on the device. */
int e2fm_index_open(const uint8_t *file, size_t size, const uint8_t key64[64], int device, e2fm_index **out);
/* The host half of e2fm_index_open alone (no GPU needed): EFMIndex::readHeader + EFMCollection::readHeader (EFMIndex.cpp:1150,
 * EFMCollection.cpp:289) incl. the key-dependent secrets — scrambling permutation, decrypted C[] (compact_alphabet + 1 entries,
 * up to c_cap copied), terminators (up to term_cap copied). Same error classes and messages as e2fm_index_open. */
int e2fm_parse_header(const uint8_t *file, size_t size, const uint8_t key64[64], e2fm_info *info, uint64_t *c_out, uint64_t c_cap,
                      int64_t *term_out, uint64_t term_cap);
/* EFMIndex::close (EFMIndex.cpp:1602) */
void e2fm_index_close(e2fm_index *ix);
int e2fm_index_info(const e2fm_index *ix, e2fm_info *out);
/* sequence terminators (EFMCollection::readHeader, EFMCollection.cpp:289): copies min(cap, n_sequences) entries */
int e2fm_index_terminators(const e2fm_index *ix, int64_t *out, uint64_t cap);
/* FASTA headers blob (EFMCollection::loadFastaHeaders, EFMCollection.cpp:97): returns its size; copies up to cap bytes */
uint64_t e2fm_index_fasta_headers(const e2fm_index *ix, char *out, uint64_t cap);
/* the CUDA stream (cudaStream_t) every kernel of this index is launched on */
void *e2fm_index_stream(const e2fm_index *ix);

/* Batched replacement of the per-pattern loop around EFMCollection::search (Main.cpp:1055-1079,
 * EFMCollection.cpp:115): patterns are `n` byte strings, pattern i = bytes[offsets[i] .. offsets[i+1]).
 * HOST buffers; the call copies them to the device, runs the search and leaves the result on the device.
 * offsets[] must be non-decreasing and offsets[n] the size of bytes[] (they are the caller's own bookkeeping and are not
 * re-validated per batch); an empty pattern has zero occurrences. */
int e2fm_search_batch(e2fm_index *ix, const uint8_t *bytes, const uint64_t *offsets, uint64_t n, int mode,
                      e2fm_result **out);
/* Fixed-length batch: pattern i = bytes[i*length .. (i+1)*length) (no offsets array to ship). */
int e2fm_search_batch_fixed(e2fm_index *ix, const uint8_t *bytes, uint64_t n, uint32_t length, int mode,
                            e2fm_result **out);
/* The pipelined form of the two calls above: offsets == NULL means fixed-length patterns of `length` bytes. When
 * counts_out (host, n entries, ideally from e2fm_host_alloc) is given, the counts of every finished chunk are copied
 * back while later chunks are still being copied in and searched, so that H2D, kernels and D2H overlap. */
int e2fm_search_batch_stream(e2fm_index *ix, const uint8_t *bytes, const uint64_t *offsets, uint64_t n, uint32_t length,
                             int mode, uint64_t *counts_out, e2fm_result **out);
/* Same two with DEVICE-resident pattern buffers (no host<->device traffic). */
int e2fm_search_batch_device(e2fm_index *ix, const uint8_t *d_bytes, const uint64_t *d_offsets, uint64_t n,
                             uint64_t total_bytes, int mode, e2fm_result **out);
int e2fm_search_batch_device_fixed(e2fm_index *ix, const uint8_t *d_bytes, uint64_t n, uint32_t length, int mode,
                                   e2fm_result **out);
/* 2-bit packed fixed-length patterns: A C G T = 0 1 2 3, base i of a pattern at bits [2i, 2i+2) of its little-endian bit string,
 * e2fm_packed_stride(length) = 16 * ceil(length / 64) bytes per pattern, unused bits zero. A quarter of the PCIe bytes of the ASCII
 * entry points, and no pattern pre-pass on the device. Only for batches the seed search takes (e2fm_info.seed_chars > 0, k <= 7,
 * length <= 128 and long enough for every shift to hold a whole seed: about (seed_chars + 2) * k bases); E2FM_ERR_FORMAT otherwise.
 * Same search, same results as the per-pattern loop around EFMCollection::search (Main.cpp:1055-1079). counts_out: optional host
 * array of n counts (NULL: fetch them with the result). */
uint32_t e2fm_packed_stride(uint32_t length);
/* host helper: ASCII (n x length) -> packed; *n_unpackable = patterns holding a byte other than A C G T (packed as if it were A:
 * send such batches through the ASCII entry points, which handle every symbol of the index's alphabet) */
int e2fm_pack_patterns(const uint8_t *ascii, uint64_t n, uint32_t length, uint8_t *packed_out, uint64_t *n_unpackable);
/* The same bit string without the padding, for the PCIe-bound host-to-host call (e2fm_search_hits with packed =
 * E2FM_PACKED_TIGHT): ceil(length / 4) bytes per pattern, base i at bits 2(i mod 4) of byte i / 4 — i.e. the first bytes of the
 * 16-byte form. A length-50 pattern travels as 13 bytes instead of 16 (or 50 as ASCII); the device widens it stage by stage. */
#define E2FM_ASCII 0
#define E2FM_PACKED_WORDS 1
#define E2FM_PACKED_TIGHT 2
uint32_t e2fm_packed_tight_stride(uint32_t length);
int e2fm_pack_patterns_tight(const uint8_t *ascii, uint64_t n, uint32_t length, uint8_t *packed_out, uint64_t *n_unpackable);
int e2fm_search_batch_packed(e2fm_index *ix, const uint8_t *packed, uint64_t n, uint32_t length, int mode, uint64_t *counts_out,
                             e2fm_result **out);
int e2fm_search_batch_packed_device(e2fm_index *ix, const uint8_t *d_packed, uint64_t n, uint32_t length, int mode,
                                    e2fm_result **out);
/* Host to host in ONE pipelined call: the batched replacement of the loop around EFMCollection::search (Main.cpp:1055-1079) that hands
 * back what that loop collects — per pattern the number of matches (EFMIndex::countOccurrences) and, in locate mode, the
 * (Occurrence::sequenceIndex, Occurrence::patternPosition) pairs of FMCollection.h:23-44 — in 32 bits, straight into the caller's
 * host arrays. Fixed-length patterns, ASCII (packed == E2FM_ASCII, n x length bytes), 2-bit packed in 16-byte words (E2FM_PACKED_WORDS,
 * e2fm_packed_stride bytes each) or 2-bit packed without padding (E2FM_PACKED_TIGHT, e2fm_packed_tight_stride bytes each). The batch
 * is cut into stages; stage c travels to the device while stage c-1 is searched and the results of stage c-2 travel back, so the call
 * takes about as long as the slower of the two PCIe directions. Page-locked buffers (e2fm_host_alloc) are needed for the overlap.
 *   counts        [n]      may be NULL
 *   occ_offsets   [n + 1]  locate: occurrences of pattern i are entries [occ_offsets[i], occ_offsets[i+1]) of the two arrays below
 *   seq_index, seq_offset  [capacity] locate; sorted by text position within a pattern (EFMIndex.cpp:2291)
 *   total         out: occurrences written. E2FM_ERR_ARG with total = the capacity needed when `capacity` is too small, or when a value
 *                 does not fit 32 bits (use e2fm_search_batch_* + e2fm_result_fetch then).
 * Batches the seed search does not take (see e2fm_search_batch_packed) run unpipelined through the generic path, same results. */
typedef struct e2fm_hits {
    uint32_t *counts;
    uint32_t *occ_offsets;
    uint32_t *seq_index;
    uint32_t *seq_offset;
    uint64_t capacity;
    uint64_t total;
} e2fm_hits;
int e2fm_search_hits(e2fm_index *ix, const uint8_t *patterns, uint64_t n, uint32_t length, int packed, int mode, e2fm_hits *hits);
/* Page-locked host memory for pattern / result buffers (plain malloc'ed buffers work too, only slower). */
void *e2fm_host_alloc(size_t bytes);
void e2fm_host_free(void *p);
/* number of visible CUDA devices (0 when there is none: every compute call then fails with E2FM_ERR_CUDA) */
int e2fm_device_count(void);

/* EFMIndex::extractSnippet(from, to) (EFMIndex.cpp:2325): the original-alphabet text in [from, to] (positions as search reports
 * them), '&' padding and terminators included. *out_len receives the length; E2FM_ERR_ARG if it exceeds cap. An empty range,
 * or one past the end, gives length 0 as in the reference. */
int e2fm_extract(e2fm_index *ix, uint64_t from, uint64_t to, char *out, uint64_t cap, uint64_t *out_len);

uint64_t e2fm_result_patterns(const e2fm_result *r);
uint64_t e2fm_result_total_positions(const e2fm_result *r);
/* D2H copy of whichever arrays are non-NULL:
 *   counts[n]            EFMIndex::countOccurrences
 *   pos_offsets[n+1]     CSR offsets into the arrays below
 *   positions[total]     sorted, de-duplicated global text positions (EFMIndex::locateOccurrences, EFMIndex.cpp:2281-2301)
 *   seq_index[total], seq_offset[total]   Occurrence{sequenceIndex, patternPosition} (EFMCollection.cpp:124-148) */
int e2fm_result_fetch(e2fm_result *r, uint64_t *counts, uint64_t *pos_offsets, uint64_t *positions,
                      uint32_t *seq_index, uint64_t *seq_offset);
/* The same in 32 bits (a third of the PCIe bytes): counts[n], pos_offsets[n+1], seq_index[total], seq_offset[total] — whichever are
 * non-NULL. What EFMCollection::search returns per pattern is (sequenceIndex, patternPosition) of every occurrence
 * (FMCollection.h:23-44): pos_offsets + seq_index + seq_offset. E2FM_ERR_ARG when a value does not fit (2^32 occurrences, a
 * sequence of 2^32 bases); counts above 2^32-1 are truncated (a count is bounded by k times the text length). */
int e2fm_result_fetch_compact(e2fm_result *r, uint32_t *counts, uint32_t *pos_offsets, uint32_t *seq_index, uint32_t *seq_offset);
/* device pointers of the same arrays (valid until e2fm_result_free) */
int e2fm_result_device_ptrs(e2fm_result *r, const uint64_t **d_counts, const uint64_t **d_pos_offsets,
                            const uint64_t **d_positions, const uint32_t **d_seq_index,
                            const uint64_t **d_seq_offset);
void e2fm_result_free(e2fm_result *r);

/* ---- collections sharded by sequence over the GPUs of one box (one process per GPU) ----
 * The reference has no distribution. What makes sharding exact is EFMCollection.cpp:218-225: every sequence is padded to a multiple
 * of k and followed by a k-long terminator, so a pattern never spans two sequences and an occurrence has the same (sequence, offset)
 * in a shard as in the whole collection. Rank r opens the index the reference built for the r-th contiguous group of sequences
 * (first_sequence = index of its first sequence in the collection, text_base = padded length of everything before it).
 *   e2fm_shardset_unique_id   one rank creates the NCCL id and hands the 128 bytes to the others (file, socket, MPI, ...)
 *   e2fm_shardset_create      ncclCommInitRank; the shard set borrows ix (close it after e2fm_shardset_destroy)
 *   e2fm_shardset_search      ONE exchange step for a fixed-length batch of n_total patterns: this rank supplies ITS slice of the
 *                             batch from host memory — patterns [lo, lo + count) with lo = e2fm_shardset_slice(n_total, world, rank,
 *                             &count), `packed` != 0: 2-bit packed (e2fm_packed_stride bytes each), else ASCII —, the slices are
 *                             all-gathered over NVLink, every rank searches the whole batch against its shard, and the answers for
 *                             slice q travel to rank q: counts by ncclReduceScatter, occurrence lists by one grouped ncclSend/ncclRecv
 *                             round, concatenated per pattern in shard order (already sorted, EFMIndex.cpp:2291). The result holds
 *                             THIS RANK'S `count` patterns: the same arrays, bit for bit, as the unsharded index gives for them
 *                             (e2fm_result_fetch / _fetch_compact / _free as usual).
 *   e2fm_shardset_last_ms     [0] H2D + batch all-gather, [1] local search, [2] NCCL traffic of the merge, [3] merge kernel, [4] whole step
 * All ranks must call create / search / destroy in the same order. */
int e2fm_shardset_unique_id(uint8_t id_out[128]);
int e2fm_shardset_create(e2fm_index *ix, const uint8_t id[128], int rank, int world, uint64_t first_sequence, uint64_t text_base,
                         e2fm_shardset **out);
void e2fm_shardset_destroy(e2fm_shardset *ss);
uint64_t e2fm_shardset_slice(uint64_t n_total, int world, int rank, uint64_t *count);
int e2fm_shardset_search(e2fm_shardset *ss, const uint8_t *slice, uint64_t n_total, uint32_t length, int packed, int mode,
                         e2fm_result **out);
int e2fm_shardset_last_ms(const e2fm_shardset *ss, float out[8]);

/* Per-phase device times (ms, CUDA events on the index stream) of the most recent batch:
 * [0] pattern H2D, [1] backward search kernel, [2] interval scan + expansion, [3] row-walk kernel, [4] pattern pre-pass (k-mer codes),
 * [5] CSR build + sort + dedupe + sequence map, [6] whole batch (first copy to last kernel), [7] firstcheck kernel (row scans of long wildcard intervals) */
int e2fm_last_batch_ms(const e2fm_index *ix, float out[8]);
/* Operation counters of the most recent batch:
 * [0] rank (= Bucket::getOffsetInformations(CountOnly) calls the reference would make, Bucket.cpp:656),
 * [1] lf (= ...(RetrieveCharacter)) and [2] mark (= Bucket::getMarkedTextPosition, Bucket.cpp:823): only the per-query walks
 *     (E2FM_OPT_DENSE 0) execute these one for one; with the dense tables they stay 0,
 * [3] row tasks, [4] distinct 32-byte sectors gathered by the backward-search kernel (rank sectors, pair-table entries, records),
 * [5] kernels launched by the batch, [6] located occurrences, [7] gathers issued by the walk / resolve kernel,
 * [8] rows streamed by firstcheck_kernel (2 bytes each, read twice), [9] chunks redone through the windowed path,
 * [10] patterns the seed search handed to the generic path, [11] 1 when the batch went through the seed search,
 * [12] pipeline stages of e2fm_search_hits (0: the call ran unpipelined), [13] why its pipeline gave up (1 row tasks, 2 occurrences,
 *      4 duplicates, 8 a pattern for the generic path, 16 output capacity), [14..15] reserved */
int e2fm_last_batch_counters(const e2fm_index *ix, uint64_t out[16]);
#define E2FM_OPT_CHUNK 2         /* patterns per internal chunk (0 = default 2^22) */
#define E2FM_OPT_WINDOW 3        /* row tasks per walk launch (0 = default 2^26) */
#define E2FM_OPT_PIPE 5          /* patterns per pipeline stage of a host-buffer batch (0 = default 2^18) */
#define E2FM_OPT_TAPER 7         /* 1: pipeline stages of a host batch halve in size towards its end (down to 1/8 of the first), so that only a
                                    small stage is left to search when the last bytes arrive; 0 (default): equal stages */
#define E2FM_OPT_PHASE_TIMERS 8  /* per-phase CUDA-event pairs (e2fm_last_batch_ms slots 1-5, 7): 0 never, 1 always, 2 (default) only for
                                    device-resident batches — on a pipelined host batch the extra event calls cost more than they show */
#define E2FM_OPT_DENSITY 6       /* test hook: preset the learned row-task / occurrence densities (per-mille per pattern) that size the
                                    fast path's buffers; too small a value forces the overflow -> windowed-path redo */
#define E2FM_OPT_SEED 9          /* 1 (default when the seed table was built): fixed-length batches whose every shift holds a whole seed
                                    go through seed_search_kernel (seed table + dense tables); 0: always the generic backward search.
                                    Same results either way. */
#define E2FM_OPT_HITS_STAGE 10   /* patterns per full pipeline stage of e2fm_search_hits; the last stage of a batch is cut into 1/2, 1/4, 1/4.
                                    0 = default: 2^20, a quarter of the batch (>= 2^17) when the batch is shorter than four stages, and
                                    the first stage cut into 1/4, 1/4, 1/2 so that the search starts early */
#define E2FM_OPT_DENSE 4         /* 1 (default when the tables were built): locate and hat verification read the dense sa[] / text[]
                                    tables that the load pipeline fills by walking LF from every marked row once (6 bytes per row
                                    of HBM); 0: walk from the marked rows per query (EFMIndex::getRowPosition, EFMIndex.cpp:2468).
                                    Same results either way. */
int e2fm_set_option(e2fm_index *ix, int option, uint64_t value);

/* ---- build side: the bucket text encoder ----
 * EFMIndex::encodeBuckets (EFMIndex.cpp:751-773) -> Bucket::encodeText (Bucket.cpp:442-536) for a run of consecutive buckets, and
 * the symbol loop of Bucket::writeText (Bucket.cpp:232-246): move-to-front, RLE0 of the runs of rank 0, poly-alphabetic substitution
 * with the bucket's Salsa20 key stream (EFMIndex::getCryptoParams, EFMIndex.cpp:1668; key = bytes 32..63 of key64, nonce = bucket
 * number), int_log2(bucketAlphaSize) bits per symbol, most significant bit first. BWT construction, the two remappings and the
 * bucket headers stay in the reference builder; this is its per-bucket step, a warp per bucket on the GPU.
 *   bucket_codes        host, n_rows: sbwt[] of the rows of buckets [first_bucket, first_bucket + n_buckets) AFTER the super-bucket
 *                       and bucket remappings (EFMIndex.cpp:841-858); every bucket holds bucket_size rows, the last may be shorter
 *   buckets_per_super_bucket, total_buckets   which buckets are encoded last row first (EFMIndex.cpp:898-901)
 *   bucket_alpha_size   [n_buckets] Bucket::bucketAlphaSize
 *   compressed_length   out [n_buckets]: EncodingResult::realLength (0 for a single-character bucket: Bucket::write writes no text)
 *   payload_offset      out [n_buckets + 1] (may be NULL without payload): byte offset of each bucket's packed text in `payload`
 *                       (8-byte aligned slots; the text is (compressed_length * int_log2(alpha) + 7) / 8 bytes, zero padded like
 *                       BitWriter::flush) — what Bucket::writeText writes after the sequenceLength field
 *   payload             out, payload_cap bytes (may be NULL); *payload_bytes = bytes needed (E2FM_ERR_ARG when the cap is too small)
 *   sequence            optional out [n_rows]: EncodingResult::sequence of each bucket at its row offset (what the reference's own
 *                       writeText would be handed instead of the packed bytes)
 *   device_ms           optional out: CUDA-event time of the kernels */
int e2fm_encode_buckets(int device, const uint8_t key64[64], const uint32_t *bucket_codes, uint64_t n_rows, uint64_t bucket_size,
                        uint64_t first_bucket, uint64_t n_buckets, uint64_t buckets_per_super_bucket, uint64_t total_buckets,
                        const uint32_t *bucket_alpha_size, uint64_t *compressed_length, uint64_t *payload_offset, uint8_t *payload,
                        uint64_t payload_cap, uint64_t *payload_bytes, uint32_t *sequence, float *device_ms);

/* ---- test hooks (parity of the individual kernels against the oracle) ---- */
/* K1: Salsa20 keystream kernel. Replaces RandomGenerator::populateKeyStreamBuffer + nextInt over salsa20.S
 * (RandomGenerator.cpp:125,60): `count` values of int_log2(max_value+1) bits starting at value `start`
 * of the stream (key = bytes 32..63 of key64, nonce = `nonce`). */
int e2fm_debug_keystream(int device, const uint8_t key64[64], uint64_t nonce, uint32_t max_value,
                         uint64_t start, uint64_t count, uint32_t *out);
/* host only: the pipeline stages e2fm_search_hits cuts a batch of n patterns into (stage_option as E2FM_OPT_HITS_STAGE, 0 = automatic):
 * bounds_out[c] .. bounds_out[c + 1] is stage c; *n_bounds = entries written (stages + 1) */
int e2fm_debug_hits_stages(uint64_t n, uint64_t stage_option, uint64_t *bounds_out, uint64_t cap, uint64_t *n_bounds);
/* HBM random-gather roofline: independent 8-byte loads at hashed addresses of a buffer_bytes buffer (make it much larger than
 * L2); returns 10^9 gathers per second (best of 3). Every gather moves one 32-byte sector over the memory system. */
int e2fm_debug_gather_bench(int device, uint64_t buffer_bytes, uint64_t n_gathers, double *giga_gathers_per_s);
/* the same with the load instruction chosen: 0 ld.global.nc.u64 (the default above), 1 ld.global.u64, 2 ld.global.cg.u64, 3 one 32-byte
 * ld.global.nc.v8.u32, 4 two 16-byte loads of one sector, 5 / 7 ld.global.nc.L2::64B / ::128B.u64, 6 ld.global.nc.L1::no_allocate.u64,
 * 8 one 32-byte ld.global.nc.L2::64B.v8.u32 (what the search kernels use for rank, seed and text sectors) */
int e2fm_debug_gather_bench_ex(int device, uint64_t buffer_bytes, uint64_t n_gathers, int variant, double *giga_gathers_per_s);
/* raw Salsa20/20 blocks [first_block, first_block + n_blocks) for a 32-byte key */
int e2fm_debug_salsa20(int device, const uint8_t key32[32], uint64_t nonce, uint64_t first_block,
                       uint64_t n_blocks, uint8_t *out);
/* K2 output: decoded BWT (compact codes, logical row order) rows [row0, row0+count) */
int e2fm_debug_bwt(const e2fm_index *ix, uint64_t row0, uint64_t count, uint32_t *out);
/* marked value of each row in [row0,row0+count) or UINT64_MAX if the row is not marked */
int e2fm_debug_marks(const e2fm_index *ix, uint64_t row0, uint64_t count, uint64_t *out);
/* seed table against the rank structure: `count` seeds of seed_chars compact codes each (text order) -> out[3*i] = interval start,
 * out[3*i+1] = rows, out[3*i+2] = 1 when exact (0: a saturated counter, the search falls back to backward steps) */
int e2fm_debug_seed(const e2fm_index *ix, const uint32_t *codes, uint64_t count, uint32_t *out);
/* LF-mapping of rows (device rank structure) */
int e2fm_debug_lf(const e2fm_index *ix, const uint64_t *rows, uint64_t count, uint64_t *out);
/* C[c] + occ(c,row) for `count` (code,row) pairs — the rank primitive of EFMIndex::backwardStep (EFMIndex.cpp:1966) */
int e2fm_debug_rank(const e2fm_index *ix, const uint32_t *codes, const uint64_t *rows, uint64_t count, uint64_t *out);

const char *e2fm_last_error(void);
const char *e2fm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* E2FM_B200_H_ */
