/* e2fm_oracle.h — CPU restatement of the reference's search path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (e2fm_b200/, include/) never does.
 *
 * Parity status: PINNED. Every function is checked against the compiled, unmodified reference
 * (oracle/_ref/e2fm_ref, built by oracle/Makefile from /root/reference) — known-answer vectors
 * (tests/golden/kat.txt), byte-identical internal dumps (`e2fm_ref dump` vs orc_dump) and search
 * results (the .res files under tests/golden) — and against naive string search.
 */
#ifndef E2FM_ORACLE_H_
#define E2FM_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_index orc_index;

/* primitives (exposed so the tests can pin them one by one) */
uint32_t orc_int_log2(uint64_t u);                                         /* Utils.cpp:58 */
void orc_salsa20_block(const uint8_t key[32], uint64_t nonce, uint64_t counter, uint8_t out[64]);
/* RandomGenerator::populateKeyStreamBuffer(nonce,max,start,end)+nextInt() x count  (RandomGenerator.cpp:125,60) */
void orc_keystream_ints(const uint8_t key[32], uint64_t nonce, uint32_t max_value, uint64_t start,
                        uint64_t count, uint32_t *out);
void orc_scrambling_key(const uint8_t key64[64], uint32_t cardinality, uint32_t *out);      /* EncryptionManager.cpp:92 */
void orc_freq_permutation(const uint8_t key64[64], uint32_t n, uint32_t *out);              /* EncryptionManager.cpp:170 */
void orc_freq_keystream(const uint8_t key64[64], uint32_t n, uint64_t max_freq, uint64_t *out); /* EncryptionManager.cpp:150 */
/* BitSetRLEEncoder::decompress (BitSetRLEEncoder.cpp:67): bits are one byte each (0/1); returns output length */
size_t orc_rle_decompress(const uint8_t *in_bits, size_t n_in, uint8_t *out_bits, size_t out_cap);

/* index: load = EFMIndex::load + readHeader + (eager) SuperBucket::load / Bucket::readHeader / readText */
orc_index *orc_open(const uint8_t *file, size_t size, const uint8_t key64[64], char *err, size_t errcap);
void orc_close(orc_index *ix);
int orc_dump(const orc_index *ix, const char *path);      /* same layout as `e2fm_ref dump` */

typedef struct {
    uint64_t n;            /* super-text length */
    uint32_t k, n_sym, mrp, rate;
    uint64_t bucket_size, superbucket_size, n_buckets, n_superbuckets;
    uint32_t compact_alphabet;
    uint64_t n_sequences;
    uint64_t initial_n;
} orc_info;
void orc_get_info(const orc_index *ix, orc_info *out);

/* one pattern: EFMIndex::exactMatch + countOccurrences (+ locateOccurrences if mode==1).
 * positions: caller buffer of pos_cap entries (sorted, deduped global text positions).
 * returns number of positions that would be written (may exceed pos_cap), or -1 on error. */
int64_t orc_search(orc_index *ix, const uint8_t *pattern, uint32_t len, int mode, uint64_t *count,
                   uint64_t *positions, uint64_t pos_cap);
/* EFMCollection::search position -> (sequenceIndex, offset)   (EFMCollection.cpp:124-148) */
void orc_map_position(const orc_index *ix, uint64_t pos, uint32_t *seq, uint64_t *off);

void orc_get_C(const orc_index *ix, uint64_t *out);                 /* compact_alphabet + 1 entries */
void orc_get_terminators(const orc_index *ix, int64_t *out);        /* n_sequences entries */
/* EFMIndex::extractSnippet(from,to) (EFMIndex.cpp:2325): returns the length written (may exceed cap), -1 without marked rows */
int64_t orc_extract(orc_index *ix, uint64_t from, uint64_t to, uint8_t *out, uint64_t cap);

/* per-row views of the decoded index (kernel-level parity tests): compact BWT codes, marked values (~0 = unmarked),
 * LF(row), and C[c] + occ(c,row) */
void orc_get_bwt(const orc_index *ix, uint64_t row0, uint64_t count, uint32_t *out);
void orc_get_marks(const orc_index *ix, uint64_t row0, uint64_t count, uint64_t *out);
void orc_lf_rows(orc_index *ix, const uint64_t *rows, uint64_t count, uint64_t *out);
void orc_rank_rows(orc_index *ix, const uint32_t *codes, const uint64_t *rows, uint64_t count, uint64_t *out);

/* op counters since open: rank = getOffsetInformations(CountOnly),
 * lf = getOffsetInformations(RetrieveCharacter), mark = getMarkedTextPosition */
void orc_get_counters(const orc_index *ix, uint64_t *n_rank, uint64_t *n_lf, uint64_t *n_mark);
void orc_reset_counters(orc_index *ix);

/* ---- build side: the bucket text encoder, pinned against the bytes the reference builder wrote ---- */
/* [0] length [1] bucketAlphaSize [2] isOdd [3] isFirstInSuperBucket [4] compressedLength [5] text bit offset in the file [6] bits per symbol */
void orc_bucket_info(const orc_index *ix, uint64_t bucket, uint64_t out[8]);
/* rows of the bucket as bucket codes, i.e. sbwt[] after both remappings (EFMIndex.cpp:841-858) */
void orc_bucket_codes(const orc_index *ix, uint64_t bucket, uint32_t *out);
/* Bucket::encodeText (Bucket.cpp:442-536); seq_out has `length` entries; returns EncodingResult::realLength */
uint64_t orc_encode_bucket(const uint8_t key32[32], uint64_t bucket_number, const uint32_t *codes, uint64_t length, uint32_t alpha,
                           int is_odd, uint32_t *seq_out);
/* symbol loop of Bucket::writeText (Bucket.cpp:232-246) + flush */
void orc_pack_sequence(const uint32_t *seq, uint64_t n, uint32_t bits, uint8_t *out);
/* Bucket::readTextUntilTo (Bucket.cpp:288-403) over a sequence in memory: the inverse of orc_encode_bucket; returns rows produced */
uint64_t orc_decode_sequence(const uint8_t key32[32], uint64_t bucket_number, const uint32_t *seq, uint64_t clen, uint64_t length,
                             uint32_t alpha, int is_odd, uint32_t *codes_out);

#ifdef __cplusplus
}
#endif
#endif
