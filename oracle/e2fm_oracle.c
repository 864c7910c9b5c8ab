/* e2fm_oracle.c — plain-C restatement of the reference's load-and-search path.
 *
 * TEST INFRASTRUCTURE ONLY (see e2fm_oracle.h). Written from the reference's behaviour, not from
 * its text; every function cites the reference file:line it restates. Deliberately naive: it
 * follows the reference's data flow (file counters + in-bucket scans), not the GPU layout.
 */
#include "e2fm_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small helpers */

/* Utils::int_log2 (Utils.cpp:58-69): number of bits needed to represent u; 0 and 1 both give 1. */
uint32_t orc_int_log2(uint64_t u) {
    uint32_t i = 1;
    uint64_t r = 1;
    while (i <= 64 && r < u) { r = 2 * r + 1; i++; }
    return i;
}

/* MSB-first bit field of width n (<=57) at absolute bit position `bit`
 * (MemoryBitReader::read, MemoryBitReader.cpp:126-174; the file image is zero-padded). */
static uint64_t rd_bits(const uint8_t *d, size_t size, uint64_t bit, uint32_t n) {
    if (n == 0) return 0;
    uint64_t byte = bit >> 3;
    uint32_t head = (uint32_t)(bit & 7);
    uint64_t w = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t idx = byte + (uint64_t)i;
        w = (w << 8) | (idx < size ? d[idx] : 0);
    }
    return (w >> (64 - head - n)) & ((n == 64) ? ~0ull : ((1ull << n) - 1));
}

typedef struct { const uint8_t *d; size_t size; uint64_t bit; } rdr;
static uint64_t R(rdr *r, uint32_t n) { uint64_t v = rd_bits(r->d, r->size, r->bit, n); r->bit += n; return v; }
static uint64_t R64(rdr *r) { uint64_t hi = R(r, 56); return (hi << 8) | R(r, 8); }   /* MemoryBitReader::getInt64 :51 */
static void Ralign(rdr *r) { r->bit = (r->bit + 7) & ~7ull; }                         /* gotoNextByteStart :82 */
/* BitReader::integerDecode (BitReader.cpp:104-133) */
static uint64_t Ridec(rdr *r, uint32_t h) {
    uint64_t i = R(r, h);
    if (i < 2) return i;
    if (i == 2) return 2 + R(r, 1);
    i--;
    return (1ull << i) + R(r, (uint32_t)i);
}

/* ------------------------------------------------------------------ Salsa20/20 (replaces salsa20.S) */

#define ROTL(v, c) (((v) << (c)) | ((v) >> (32 - (c))))
#define QR(a, b, c, d) do { b ^= ROTL(a + d, 7); c ^= ROTL(b + a, 9); d ^= ROTL(c + b, 13); a ^= ROTL(d + c, 18); } while (0)

static uint32_t ld32le(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* One 64-byte block. State layout is the standard one; the reference's asm keeps a permuted copy
 * (salsa20.S:4567-4882) but emits the standard keystream (pinned by the ECRYPT Set 1 vector 0 KAT).
 * 256-bit key, 64-bit nonce as little-endian u64 (RandomGenerator.cpp:27-28,128), 64-bit block counter
 * (RandomGenerator.cpp:175-179: low word = counter, high word = 0). */
void orc_salsa20_block(const uint8_t key[32], uint64_t nonce, uint64_t counter, uint8_t out[64]) {
    static const uint8_t sigma[16] = {'e','x','p','a','n','d',' ','3','2','-','b','y','t','e',' ','k'};
    uint32_t in[16], x[16];
    in[0] = ld32le(sigma);      in[5] = ld32le(sigma + 4); in[10] = ld32le(sigma + 8); in[15] = ld32le(sigma + 12);
    for (int i = 0; i < 4; i++) { in[1 + i] = ld32le(key + 4 * i); in[11 + i] = ld32le(key + 16 + 4 * i); }
    in[6] = (uint32_t)nonce; in[7] = (uint32_t)(nonce >> 32);
    in[8] = (uint32_t)counter; in[9] = (uint32_t)(counter >> 32);
    memcpy(x, in, sizeof x);
    for (int r = 0; r < 10; r++) {
        QR(x[0], x[4], x[8], x[12]);  QR(x[5], x[9], x[13], x[1]);  QR(x[10], x[14], x[2], x[6]);  QR(x[15], x[3], x[7], x[11]);
        QR(x[0], x[1], x[2], x[3]);   QR(x[5], x[6], x[7], x[4]);   QR(x[10], x[11], x[8], x[9]);  QR(x[15], x[12], x[13], x[14]);
    }
    for (int i = 0; i < 16; i++) {
        uint32_t v = x[i] + in[i];
        out[4 * i] = (uint8_t)v; out[4 * i + 1] = (uint8_t)(v >> 8); out[4 * i + 2] = (uint8_t)(v >> 16); out[4 * i + 3] = (uint8_t)(v >> 24);
    }
}

/* ------------------------------------------------------------------ RandomGenerator restatement */

typedef struct {
    const uint8_t *key;      /* 32 bytes */
    uint64_t nonce;
    uint32_t counter;        /* start block of the current buffer (RandomGenerator.h:52) */
    uint32_t buf_size;       /* keyStreamBufferSize */
    uint8_t *buf;
    uint32_t generated, live_bytes;
    uint32_t bits_buffer;    /* 32-bit on purpose: RandomGenerator.h:35 (quirk 5) */
    uint32_t live_bits;
    uint32_t bits_to_extract;
} rg_t;

static void rg_init(rg_t *g, uint32_t buf_size, const uint8_t *key, uint64_t nonce) {
    memset(g, 0, sizeof *g);
    g->key = key; g->nonce = nonce; g->buf_size = buf_size;
    g->buf = (uint8_t *)malloc((size_t)buf_size + 64);
}
static void rg_free(rg_t *g) { free(g->buf); g->buf = NULL; }

/* RandomGenerator::populateKeyStreamBuffer(bytesToGenerate) (:157-172): always restarts at `counter`,
 * so a refill REPEATS the same stream (quirk 6). */
static void rg_fill(rg_t *g, uint32_t bytes) {
    uint32_t blocks = (bytes + 63) / 64;
    for (uint32_t b = 0; b < blocks; b++) orc_salsa20_block(g->key, g->nonce, (uint64_t)g->counter + b, g->buf + 64 * (size_t)b);
    g->generated = bytes; g->live_bytes = bytes;
}
static uint32_t rg_byte(rg_t *g) {                      /* extractByteFromKeyStreamBuffer :116-122 */
    if (g->live_bytes == 0) rg_fill(g, g->buf_size);
    uint32_t c = g->buf[g->generated - g->live_bytes];
    g->live_bytes--;
    return c;
}
static uint32_t rg_next(rg_t *g, uint32_t n) {          /* RandomGenerator::next :96-113 */
    uint32_t live = g->live_bits, buf = g->bits_buffer;
    if (live < n) {
        do { buf = (buf << 8) | rg_byte(g); live += 8; } while (live < n);
        g->bits_buffer = buf;
    }
    g->live_bits = live - n;
    uint32_t sh = live - n;
    uint32_t v = sh >= 32 ? 0 : (buf >> sh);
    return v & (n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1));
}
static uint32_t rg_next_below(rg_t *g, uint32_t n) {    /* nextInt(n) :69-78: rejection sampling in [0,n-1] */
    uint32_t bits = orc_int_log2((uint64_t)(uint32_t)(n - 1)), v;
    do v = rg_next(g, bits); while (v > n - 1);
    return v;
}
/* populateKeyStreamBuffer(nonce,max,start,end) :125-154 + gotoBlockOffset :184-192 */
static void rg_populate(rg_t *g, uint64_t nonce, uint32_t max_value, uint64_t start, uint64_t end) {
    g->nonce = nonce;
    g->bits_to_extract = orc_int_log2((uint64_t)max_value + 1);
    uint64_t sbits = start * g->bits_to_extract, ebits = (end + 1) * g->bits_to_extract;
    uint64_t sblk = sbits / 512, soff = sbits % 512, eblk = ebits / 512;
    g->counter = (uint32_t)sblk;
    uint64_t bytes = 64 * (eblk - sblk + 1);
    if (bytes > g->buf_size) bytes = g->buf_size;
    rg_fill(g, (uint32_t)bytes);
    uint32_t byte_no = (uint32_t)(soff / 8), bit_no = (uint32_t)(soff % 8);
    g->live_bytes = g->generated - byte_no;
    g->bits_buffer = g->buf[g->generated - g->live_bytes];
    g->live_bytes--;
    g->live_bits = 8 - bit_no;
}

void orc_keystream_ints(const uint8_t key[32], uint64_t nonce, uint32_t max_value, uint64_t start, uint64_t count, uint32_t *out) {
    if (count == 0) return;
    rg_t g; rg_init(&g, 131072, key, 0);                 /* EFMIndex.cpp:1547 */
    rg_populate(&g, nonce, max_value, start, start + count - 1);
    for (uint64_t i = 0; i < count; i++) out[i] = rg_next(&g, g.bits_to_extract);   /* nextInt() :60 */
    rg_free(&g);
}

/* EncryptionManager::computeScramblingKey + shuffleExcludingFirstElement (EncryptionManager.cpp:92-117) */
void orc_scrambling_key(const uint8_t key64[64], uint32_t card, uint32_t *out) {
    for (uint32_t i = 0; i < card; i++) out[i] = i;
    rg_t g; rg_init(&g, 16384, key64 /* monoSeed = key[0:32] */, 0);
    for (uint32_t i = card; i > 2; i--) {
        uint32_t j;
        do j = rg_next_below(&g, i); while (j == 0);
        uint32_t t = out[i - 1]; out[i - 1] = out[j]; out[j] = t;
    }
    rg_free(&g);
}
/* computeFrequenciesPermutation + shuffle (EncryptionManager.cpp:170-180,119-129) */
void orc_freq_permutation(const uint8_t key64[64], uint32_t n, uint32_t *out) {
    for (uint32_t i = 0; i < n; i++) out[i] = i;
    rg_t g; rg_init(&g, 16384, key64, 1);
    for (uint32_t i = n; i > 1; i--) {
        uint32_t j;
        do j = rg_next_below(&g, i); while (j == 0);
        uint32_t t = out[i - 1]; out[i - 1] = out[j]; out[j] = t;
    }
    rg_free(&g);
}
/* computeCumulativeFreqsKeyStream (EncryptionManager.cpp:150-161); nextInt takes uint32_t so max+1 truncates */
void orc_freq_keystream(const uint8_t key64[64], uint32_t n, uint64_t max_freq, uint64_t *out) {
    rg_t g; rg_init(&g, 16384, key64 + 32 /* polySeed */, 2);
    for (uint32_t i = 0; i < n; i++) out[i] = rg_next_below(&g, (uint32_t)(max_freq + 1));
    rg_free(&g);
}

/* BitSetRLEEncoder::decompress (BitSetRLEEncoder.cpp:67-135). Tokens: "0" digit 0, "10" digit 1,
 * "11" = flush the pending digits (x digits of value N -> 2^x+N-1 zeros) then a set bit. Digits left
 * over at the end are flushed and followed by a set bit as well (:116-133). */
size_t orc_rle_decompress(const uint8_t *in, size_t n_in, uint8_t *out, size_t cap) {
    size_t p = 0; uint64_t N = 0; uint32_t x = 0; int pending = 0;
    for (size_t i = 0; i < n_in; i++) {
        if (!pending) {
            if (in[i] == 0) { N = 2 * N; x++; } else pending = 1;
        } else {
            if (in[i] == 0) { N = 2 * N + 1; x++; }
            else {
                if (x > 0) { uint64_t m = (1ull << x) + N - 1; for (uint64_t j = 0; j < m; j++) { if (p < cap) out[p] = 0; p++; } }
                if (p < cap) out[p] = 1;
                p++; N = 0; x = 0;
            }
            pending = 0;
        }
    }
    if (x > 0) {
        uint64_t m = (1ull << x) + N - 1;
        for (uint64_t j = 0; j < m; j++) { if (p < cap) out[p] = 0; p++; }
        if (p < cap) out[p] = 1;
        p++;
    }
    return p;
}

/* ------------------------------------------------------------------ index */

typedef struct {
    uint64_t length; uint32_t alpha, tbits; int is_odd, is_first;
    uint64_t clen, text_bit, mark_bit;
    uint64_t *table;      /* sbAlpha counters: own (even) or the next bucket's (odd); NULL if first in SB */
    uint8_t *bits;        /* sbAlpha */
    uint64_t n_marked;
} bucket_t;

struct orc_index {
    uint8_t *file; size_t size;
    uint8_t key[64];
    uint32_t n_sym; uint8_t syms[256]; uint32_t k;
    uint64_t n, initial_n, final_n, last_pos, sb_size, b_size; uint32_t mrp, rate;
    uint64_t c_start, mrb_start, sbpos_start, bpos_start, fasta_start;
    uint32_t sigma_k, A;
    uint32_t *perm, *inv_perm;           /* natural -> scrambled, scrambled -> natural (ScrambledSuperAlphabet) */
    int64_t *compact_of_scr;             /* charactersRemappings, -1 if absent */
    uint32_t *scr_of_compact;            /* decodeMap */
    uint64_t *C;
    uint64_t mrn; uint32_t val_bits; uint64_t *mr_buckets;
    uint64_t S, B, per; uint64_t *sb_start, *b_start;
    uint8_t *sb_bits; uint64_t *sb_occ; uint32_t *sb_alpha;
    bucket_t *bk;
    uint32_t *bwt;                       /* n compact codes, logical order */
    uint8_t *marked; uint64_t *mark_val; /* per row */
    uint64_t *row_of_mark;               /* mrn, ~0 if unset */
    uint64_t n_term; int64_t *term;
    uint64_t c_rank, c_lf, c_mark;
};

static void seterr(char *err, size_t cap, const char *m) { if (err && cap) { strncpy(err, m, cap - 1); err[cap - 1] = 0; } }

/* reads "1-bit flag, unaligned i64 size, bits" (EFMIndex.cpp:1200-1215, SuperBucket.cpp:41-57, Bucket.cpp:581-595) */
static uint8_t *read_bitset(rdr *r, uint64_t *out_size) {
    int compressed = (int)R(r, 1);
    uint64_t sz = R64(r);
    if (sz > (1ull << 32)) return NULL;
    uint8_t *raw = (uint8_t *)malloc(sz + 1);
    for (uint64_t i = 0; i < sz; i++) raw[i] = (uint8_t)R(r, 1);
    if (!compressed) { *out_size = sz; return raw; }
    size_t need = orc_rle_decompress(raw, sz, NULL, 0);
    uint8_t *o = (uint8_t *)calloc(need + 1, 1);
    orc_rle_decompress(raw, sz, o, need);
    free(raw); *out_size = need; return o;
}

static uint32_t sym_index(const orc_index *ix, uint8_t c) {   /* ScrambledSuperAlphabet::originalSymbolsIndexOf :65 */
    for (uint32_t i = 0; i < ix->n_sym; i++) if (ix->syms[i] == c) return i;
    return ix->n_sym;
}

/* Bucket::readTextUntilTo (Bucket.cpp:288-403) for the whole bucket + decode maps (Bucket.cpp:598-606,
 * SuperBucket.cpp:62-68): emits compact codes at logical offsets. */
static int decode_bucket(orc_index *ix, uint64_t bn) {
    bucket_t *b = &ix->bk[bn];
    uint64_t sn = bn / ix->per;
    uint32_t A = ix->A, sbA = ix->sb_alpha[sn];
    uint32_t *sb_decode = (uint32_t *)malloc(sizeof(uint32_t) * (sbA + 1));
    for (uint32_t j = 0, c = 0; j < A; j++) if (ix->sb_bits[sn * A + j]) sb_decode[c++] = j;
    uint32_t *bk_decode = (uint32_t *)malloc(sizeof(uint32_t) * (b->alpha + 1));
    for (uint32_t j = 0, c = 0; j < sbA; j++) if (b->bits[j]) bk_decode[c++] = sb_decode[j];
    uint32_t *out = ix->bwt + bn * ix->b_size;
    uint64_t len = b->length, p = 0;
    if (b->alpha <= 1) {                                   /* expandSingleCharacterUntilTo :414 */
        for (p = 0; p < len; p++) out[p] = bk_decode[0];
    } else {
        uint32_t *key = (uint32_t *)malloc(sizeof(uint32_t) * (b->clen + 1));
        orc_keystream_ints(ix->key + 32, bn, b->alpha, 0, b->clen, key);       /* computeBucketKeyStream :266 */
        uint32_t *mtf = (uint32_t *)malloc(sizeof(uint32_t) * b->alpha);       /* ArrayMTFList */
        for (uint32_t i = 0; i < b->alpha; i++) mtf[i] = i;
        int64_t mod = (int64_t)b->alpha + 1;
        uint64_t i = 0;
        while (p < len && i < b->clen) {
#define SYM(ii) ((int64_t)((((int64_t)rd_bits(ix->file, ix->size, b->text_bit + (ii) * b->tbits, b->tbits) - (int64_t)key[ii]) % mod + mod) % mod))
            int64_t c = SYM(i);                             /* invertPoly :406 */
            if (c > 1) {
                uint32_t pos = (uint32_t)(c - 1), e = mtf[pos];
                memmove(mtf + 1, mtf, sizeof(uint32_t) * pos);                 /* ArrayMTFList::moveToFront :32 */
                mtf[0] = e;
                out[b->is_odd ? len - p - 1 : p] = bk_decode[e];
                p++; i++;
            } else {
                uint64_t x = 0, N = 0;
                while (i + x < b->clen) { int64_t d = SYM(i + x); if (d >= 2) break; N = 2 * N + (uint64_t)d; x++; }
                uint64_t m = (1ull << x) + N - 1;           /* :377 */
                for (uint64_t j = 0; j < m && p < len; j++) { out[b->is_odd ? len - p - 1 : p] = bk_decode[mtf[0]]; p++; }
                i += x;
            }
#undef SYM
        }
        free(key); free(mtf);
    }
    free(sb_decode); free(bk_decode);
    return p == len ? 0 : -1;
}

orc_index *orc_open(const uint8_t *file, size_t size, const uint8_t key64[64], char *err, size_t errcap) {
    orc_index *ix = (orc_index *)calloc(1, sizeof *ix);
    ix->file = (uint8_t *)calloc(size + 16, 1);
    memcpy(ix->file, file, size); ix->size = size;
    memcpy(ix->key, key64, 64);
    rdr r = { ix->file, size, 0 };
    /* EFMIndex::readHeader (EFMIndex.cpp:1150-1329) */
    if (size < 64 || R(&r, 8) != 'E' || R(&r, 8) != 'G') { seterr(err, errcap, "File type is not Encrypted and Compressed Genomic Index"); orc_close(ix); return NULL; }
    ix->n_sym = (uint32_t)R(&r, 8);
    for (uint32_t i = 0; i < ix->n_sym; i++) ix->syms[i] = (uint8_t)R(&r, 8);
    ix->k = (uint32_t)R(&r, 8);
    ix->n = R64(&r); ix->initial_n = R64(&r); ix->final_n = R64(&r); ix->last_pos = R64(&r);
    ix->sb_size = R64(&r); ix->b_size = R64(&r); ix->mrp = (uint32_t)R(&r, 8);
    (void)R64(&r);                                         /* maximumBucketAlphaSize: uninitialised at build (quirk 7) */
    ix->c_start = R64(&r); ix->mrb_start = R64(&r); ix->sbpos_start = R64(&r); ix->bpos_start = R64(&r);
    ix->fasta_start = R64(&r);                             /* EFMCollection::readSuspendedInformations, EFMCollection.cpp:242 */
    if (ix->b_size == 0 || ix->sb_size % ix->b_size || ix->n == 0 || ix->k == 0 || ix->n_sym == 0) { seterr(err, errcap, "corrupt header"); orc_close(ix); return NULL; }
    ix->per = ix->sb_size / ix->b_size;
    ix->B = (ix->n + ix->b_size - 1) / ix->b_size;
    ix->S = (ix->n + ix->sb_size - 1) / ix->sb_size;
    uint64_t card = 1; for (uint32_t i = 0; i < ix->k; i++) card *= ix->n_sym;
    ix->sigma_k = (uint32_t)card;
    /* scrambled super-alphabet (EFMIndex.cpp:516-539, ScrambledSuperAlphabet.cpp:16-43) */
    ix->perm = (uint32_t *)malloc(sizeof(uint32_t) * card);
    ix->inv_perm = (uint32_t *)malloc(sizeof(uint32_t) * card);
    orc_scrambling_key(ix->key, ix->sigma_k, ix->perm);
    if (sym_index(ix, 'N') < ix->n_sym) {
        uint32_t all_n = ix->sigma_k - 1, nat = 0, pos = 0;
        for (uint32_t i = 0; i < ix->k; i++) nat = nat * ix->n_sym + sym_index(ix, 'N');
        uint32_t rep = ix->perm[nat];
        for (uint32_t i = 0; i < ix->sigma_k; i++) if (ix->perm[i] == all_n) { pos = i; break; }
        ix->perm[pos] = rep; ix->perm[nat] = all_n;
    }
    for (uint32_t i = 0; i < ix->sigma_k; i++) ix->inv_perm[ix->perm[i]] = i;
    /* compact alphabet bitset :1200-1239 */
    uint64_t bs_size = 0;
    uint8_t *abits = read_bitset(&r, &bs_size);
    if (!abits || bs_size != card) { seterr(err, errcap, "alphabet bitset size mismatch"); free(abits); orc_close(ix); return NULL; }
    ix->compact_of_scr = (int64_t *)malloc(sizeof(int64_t) * card);
    ix->scr_of_compact = (uint32_t *)malloc(sizeof(uint32_t) * (card + 1));
    ix->A = 0;
    for (uint32_t i = 0; i < ix->sigma_k; i++) {
        if (abits[i]) { ix->compact_of_scr[i] = ix->A; ix->scr_of_compact[ix->A] = i; ix->A++; } else ix->compact_of_scr[i] = -1;
    }
    free(abits);
    uint32_t A = ix->A;
    /* C[]: inverse transposition + inverse poly-alphabetic substitution :1241-1272 */
    Ralign(&r);
    if (r.bit != ix->c_start * 8) { seterr(err, errcap, "characterOccurrencesStart mismatch"); orc_close(ix); return NULL; }
    uint32_t nb = (uint32_t)R(&r, 8);
    uint64_t maxf = (nb >= 64 ? ~0ull : ((1ull << nb) - 1));
    uint32_t *fperm = (uint32_t *)malloc(sizeof(uint32_t) * (A + 1));
    uint64_t *fks = (uint64_t *)malloc(sizeof(uint64_t) * (A + 1));
    uint64_t *enc = (uint64_t *)malloc(sizeof(uint64_t) * (A + 1));
    orc_freq_permutation(ix->key, A + 1, fperm);
    for (uint32_t i = 0; i <= A; i++) enc[fperm[i]] = R(&r, nb);
    orc_freq_keystream(ix->key, A + 1, maxf, fks);
    ix->C = (uint64_t *)malloc(sizeof(uint64_t) * (A + 1));
    for (uint32_t i = 0; i <= A; i++) ix->C[i] = (enc[i] - fks[i]) & maxf;   /* unsigned wrap mod 2^nb == the reference's signed fix-up */
    free(fperm); free(fks); free(enc);
    Ralign(&r);
    /* markedRowsBuckets :1276-1287 */
    if (ix->mrp > 0) {
        ix->rate = 100 / ix->mrp;
        ix->mrn = ix->n / ix->rate + 1;
        ix->val_bits = orc_int_log2(ix->n / ix->rate - 1);          /* load-side formula :1280 (quirk 2) */
        uint32_t bb = orc_int_log2(ix->B - 1);
        ix->mr_buckets = (uint64_t *)malloc(sizeof(uint64_t) * ix->mrn);
        r.bit = ix->mrb_start * 8;
        for (uint64_t i = 0; i < ix->mrn; i++) ix->mr_buckets[i] = R(&r, bb);
    }
    ix->sb_start = (uint64_t *)malloc(sizeof(uint64_t) * ix->S);
    ix->b_start = (uint64_t *)malloc(sizeof(uint64_t) * (ix->B + 1));
    r.bit = ix->sbpos_start * 8; for (uint64_t i = 0; i < ix->S; i++) ix->sb_start[i] = R64(&r);
    r.bit = ix->bpos_start * 8;  for (uint64_t i = 0; i < ix->B; i++) ix->b_start[i] = R64(&r);
    /* EFMCollection::readHeader (EFMCollection.cpp:289-303): terminators follow the bucket-start table */
    ix->n_term = R(&r, 32);
    uint32_t tb = (uint32_t)R(&r, 8);
    ix->term = (int64_t *)malloc(sizeof(int64_t) * (ix->n_term + 1));
    { int64_t tp = 0; for (uint64_t i = 0; i < ix->n_term; i++) { tp += (int64_t)R(&r, tb); ix->term[i] = tp; } }

    /* SuperBucket::load (SuperBucket.cpp:37-79) for every SB */
    ix->sb_bits = (uint8_t *)calloc(ix->S * A, 1);
    ix->sb_occ = (uint64_t *)calloc(ix->S * A, sizeof(uint64_t));
    ix->sb_alpha = (uint32_t *)calloc(ix->S, sizeof(uint32_t));
    for (uint64_t sn = 0; sn < ix->S; sn++) {
        r.bit = ix->sb_start[sn] * 8;
        uint64_t sz; uint8_t *bits = read_bitset(&r, &sz);
        if (!bits || sz != A) { seterr(err, errcap, "superbucket bitset size mismatch"); free(bits); orc_close(ix); return NULL; }
        memcpy(ix->sb_bits + sn * A, bits, A); free(bits);
        uint32_t cnt = 0; for (uint32_t j = 0; j < A; j++) cnt += ix->sb_bits[sn * A + j];
        ix->sb_alpha[sn] = cnt;
        if (sn > 0) { uint32_t b2 = (uint32_t)R(&r, 8); for (uint32_t j = 0; j < A; j++) ix->sb_occ[sn * A + j] = R(&r, b2); }
    }
    /* Bucket::readHeader (Bucket.cpp:549-636) for every bucket */
    ix->bk = (bucket_t *)calloc(ix->B, sizeof(bucket_t));
    ix->bwt = (uint32_t *)malloc(sizeof(uint32_t) * ix->n);
    ix->marked = (uint8_t *)calloc(ix->n, 1);
    ix->mark_val = (uint64_t *)malloc(sizeof(uint64_t) * ix->n);
    memset(ix->mark_val, 0xFF, sizeof(uint64_t) * ix->n);
    if (ix->mrn) { ix->row_of_mark = (uint64_t *)malloc(sizeof(uint64_t) * ix->mrn); memset(ix->row_of_mark, 0xFF, sizeof(uint64_t) * ix->mrn); }
    for (uint64_t bn = 0; bn < ix->B; bn++) {
        bucket_t *b = &ix->bk[bn];
        uint64_t sn = bn / ix->per; uint32_t sbA = ix->sb_alpha[sn];
        uint64_t start = bn * ix->b_size, end = (bn + 1) * ix->b_size - 1;
        if (end > ix->n - 1) end = ix->n - 1;
        b->length = end - start + 1;
        b->is_first = (bn % ix->per) == 0;
        b->is_odd = (bn % 2 == 0) && !b->is_first && (bn != ix->B - 1);        /* EFMIndex.cpp:898-901 */
        r.bit = ix->b_start[bn] * 8;
        uint64_t m_start = R64(&r), t_start = R64(&r);
        if (!b->is_first) {
            uint64_t save = 0;
            if (b->is_odd) { save = r.bit; r.bit = (ix->b_start[bn + 1] + 16) * 8; }   /* :565-569 */
            uint32_t h = (uint32_t)R(&r, 8);
            b->table = (uint64_t *)malloc(sizeof(uint64_t) * (sbA + 1));
            for (uint32_t j = 0; j < sbA; j++) b->table[j] = Ridec(&r, h);
            if (b->is_odd) r.bit = save;
        }
        uint64_t sz; b->bits = read_bitset(&r, &sz);
        if (!b->bits || sz != sbA) { seterr(err, errcap, "bucket bitset size mismatch"); orc_close(ix); return NULL; }
        b->alpha = 0; for (uint32_t j = 0; j < sbA; j++) b->alpha += b->bits[j];
        if (ix->mrp > 0) {                                  /* :610-618 */
            uint64_t csz = R64(&r);
            uint8_t *cb = (uint8_t *)malloc(csz + 1);
            for (uint64_t i = 0; i < csz; i++) cb[i] = (uint8_t)R(&r, 1);
            orc_rle_decompress(cb, csz, ix->marked + start, b->length);      /* truncate/pad to length */
            free(cb);
        }
        Ralign(&r);
        b->clen = R64(&r);                                  /* junk when alpha==1 (Bucket.cpp:168-169) */
        b->tbits = orc_int_log2(b->alpha);
        b->text_bit = t_start * 8; b->mark_bit = m_start * 8;
        if (b->alpha > 1 && (b->clen > b->length || b->text_bit + b->clen * b->tbits > (uint64_t)size * 8)) { seterr(err, errcap, "corrupt bucket (compressedLength)"); orc_close(ix); return NULL; }
        if (decode_bucket(ix, bn) != 0) { seterr(err, errcap, "bucket text decode came up short (wrong key?)"); orc_close(ix); return NULL; }
        /* marked values: Bucket::getMarkedTextPosition / readMarkedRows (Bucket.cpp:823-882) */
        uint64_t rk = 0;
        for (uint64_t o = 0; o < b->length; o++) if (ix->marked[start + o]) {
            uint64_t v = rd_bits(ix->file, size, b->mark_bit + rk * ix->val_bits, ix->val_bits);
            ix->mark_val[start + o] = v;
            if (v < ix->mrn) ix->row_of_mark[v] = start + o;   /* markedRowsReverseCache */
            rk++;
        }
        b->n_marked = rk;
    }
    return ix;
}

void orc_close(orc_index *ix) {
    if (!ix) return;
    if (ix->bk) for (uint64_t b = 0; b < ix->B; b++) { free(ix->bk[b].table); free(ix->bk[b].bits); }
    free(ix->bk); free(ix->file); free(ix->perm); free(ix->inv_perm); free(ix->compact_of_scr); free(ix->scr_of_compact);
    free(ix->C); free(ix->mr_buckets); free(ix->sb_start); free(ix->b_start); free(ix->sb_bits); free(ix->sb_occ);
    free(ix->sb_alpha); free(ix->bwt); free(ix->marked); free(ix->mark_val); free(ix->row_of_mark); free(ix->term);
    free(ix);
}

void orc_get_info(const orc_index *ix, orc_info *o) {
    o->n = ix->n; o->k = ix->k; o->n_sym = ix->n_sym; o->mrp = ix->mrp; o->rate = ix->rate;
    o->bucket_size = ix->b_size; o->superbucket_size = ix->sb_size; o->n_buckets = ix->B; o->n_superbuckets = ix->S;
    o->compact_alphabet = ix->A; o->n_sequences = ix->n_term; o->initial_n = ix->initial_n;
}
void orc_get_counters(const orc_index *ix, uint64_t *a, uint64_t *b, uint64_t *c) { *a = ix->c_rank; *b = ix->c_lf; *c = ix->c_mark; }
void orc_reset_counters(orc_index *ix) { ix->c_rank = ix->c_lf = ix->c_mark = 0; }

static void dw(FILE *f, uint64_t v) { fwrite(&v, 8, 1, f); }
int orc_dump(const orc_index *ix, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    uint32_t A = ix->A;
    dw(f, 0x31504d444d463245ull); dw(f, ix->n); dw(f, ix->k); dw(f, ix->n_sym);
    for (uint32_t i = 0; i < ix->n_sym; i++) dw(f, ix->syms[i]);
    dw(f, ix->initial_n); dw(f, ix->final_n); dw(f, ix->last_pos); dw(f, ix->sb_size); dw(f, ix->b_size); dw(f, ix->mrp);
    dw(f, A);
    for (uint32_t i = 0; i <= A; i++) dw(f, ix->C[i]);
    dw(f, ix->sigma_k);
    for (uint32_t i = 0; i < ix->sigma_k; i++) dw(f, ix->compact_of_scr[i] < 0 ? ~0ull : (uint64_t)ix->compact_of_scr[i]);
    for (uint32_t i = 0; i < ix->sigma_k; i++) dw(f, ix->perm[i]);
    dw(f, ix->mrn);
    for (uint64_t i = 0; i < ix->mrn; i++) dw(f, ix->mr_buckets[i]);
    dw(f, ix->S); dw(f, ix->B);
    for (uint64_t s = 0; s < ix->S; s++) {
        dw(f, ix->sb_alpha[s]);
        for (uint32_t i = 0; i < A; i++) dw(f, ix->sb_bits[s * A + i]);
        for (uint32_t i = 0; i < A; i++) dw(f, ix->sb_occ[s * A + i]);
    }
    for (uint64_t bn = 0; bn < ix->B; bn++) {
        const bucket_t *b = &ix->bk[bn];
        uint32_t sbA = ix->sb_alpha[bn / ix->per];
        uint64_t start = bn * ix->b_size;
        dw(f, b->length); dw(f, b->alpha); dw(f, (uint64_t)b->is_odd); dw(f, (uint64_t)b->is_first);
        dw(f, b->alpha > 1 ? b->clen : 0);
        dw(f, b->table != NULL);
        if (b->table) for (uint32_t i = 0; i < sbA; i++) dw(f, b->table[i]);
        for (uint32_t i = 0; i < sbA; i++) dw(f, b->bits[i]);
        for (uint64_t i = 0; i < b->length; i++) dw(f, ix->marked[start + i]);
        for (uint64_t i = 0; i < b->length; i++) dw(f, ix->bwt[start + i]);
        dw(f, b->n_marked);
        for (uint64_t i = 0; i < b->length; i++) if (ix->marked[start + i]) { dw(f, i); dw(f, ix->mark_val[start + i]); }
    }
    dw(f, ix->n_term);
    for (uint64_t i = 0; i < ix->n_term; i++) dw(f, (uint64_t)ix->term[i]);
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------ rank / LF */

/* SuperBucket::getRemappedCode (SuperBucket.cpp:82-90) / Bucket::getRemappedCode (Bucket.cpp:783-791) */
static int64_t remap(const uint8_t *bits, uint32_t code) {
    if (!bits[code]) return -1;
    int64_t r = 0;
    for (uint32_t i = 0; i < code; i++) r += bits[i];
    return r;
}

/* Bucket::getOffsetInformations + occ (Bucket.cpp:656-779): inclusive in-superbucket rank of compact
 * code c up to logical offset `off` of bucket bn. sb_code must be >= 0. */
static uint64_t bucket_rank(const orc_index *ix, uint64_t bn, uint32_t c, int64_t sb_code, uint64_t off) {
    const bucket_t *b = &ix->bk[bn];
    if (!b->bits[sb_code]) return b->is_first ? 0 : b->table[sb_code];            /* :657-658 */
    const uint32_t *t = ix->bwt + bn * ix->b_size;
    uint64_t cnt = 0;
    if (b->is_odd) {                                                             /* :746-761 */
        for (uint64_t p = off + 1; p < b->length; p++) cnt += (t[p] == c);
        return b->table[sb_code] - cnt;
    }
    for (uint64_t p = 0; p <= off; p++) cnt += (t[p] == c);                       /* :763-778 */
    return (b->is_first ? 0 : b->table[sb_code]) + cnt;
}

/* occ(c,row) part of EFMIndex::backwardStep (EFMIndex.cpp:1966-2001) */
static uint64_t occ_upto(orc_index *ix, uint32_t c, uint64_t row) {
    uint64_t bn = row / ix->b_size, bo = row % ix->b_size, sn = bn / ix->per;
    uint64_t in_sb = 0;
    int64_t sbc = remap(ix->sb_bits + sn * ix->A, c);
    if (sbc != -1) { in_sb = bucket_rank(ix, bn, c, sbc, bo); ix->c_rank++; }
    return ix->sb_occ[sn * ix->A + c] + in_sb;
}
static void backward_step(orc_index *ix, uint32_t c, uint64_t *sp, uint64_t *ep) {
    uint64_t nsp = ix->C[c] + occ_upto(ix, c, *sp - 1);
    uint64_t nep = ix->C[c] + occ_upto(ix, c, *ep) - 1;
    *sp = nsp; *ep = nep;
}
/* LF-mapping as in getRowPosition / backwardStepIfMatch / extractSuperCharacter (RetrieveCharacter mode) */
static uint64_t lf(orc_index *ix, uint64_t row, uint32_t *code_out) {
    uint64_t bn = row / ix->b_size, bo = row % ix->b_size, sn = bn / ix->per;
    uint32_t c = ix->bwt[row];
    int64_t sbc = remap(ix->sb_bits + sn * ix->A, c);
    ix->c_lf++;
    if (code_out) *code_out = c;
    return ix->C[c] + ix->sb_occ[sn * ix->A + c] + bucket_rank(ix, bn, c, sbc, bo) - 1;
}
/* EFMIndex::getRowPosition (EFMIndex.cpp:2468-2497) */
static uint64_t row_position(orc_index *ix, uint64_t row) {
    uint64_t t = 0;
    for (;;) {
        ix->c_mark++;
        if (ix->marked[row]) return (ix->mark_val[row] * ix->rate + t) % ix->n;
        row = lf(ix, row, NULL); t++;
    }
}
/* k-mer text of a compact code: decodeMap -> ScrambledSuperAlphabet::getSymbol (ScrambledSuperAlphabet.cpp:90-112) */
static void kmer_of(const orc_index *ix, uint32_t compact, uint8_t *out) {
    uint32_t nat = ix->inv_perm[ix->scr_of_compact[compact]];
    for (int i = (int)ix->k - 1; i >= 0; i--) { out[i] = ix->syms[nat % ix->n_sym]; nat /= ix->n_sym; }
}
/* EFMIndex::extractSuperCharacter (EFMIndex.cpp:2216-2269) */
static void super_char_at(orc_index *ix, uint64_t position, uint8_t *out) {
    uint64_t nv = position / ix->rate + 1, v = nv % ix->mrn;
    uint64_t row = ix->row_of_mark[v];                       /* markedRowsBuckets[v] + getMarkedBwtPosition */
    uint64_t nxt = nv * ix->rate;
    if (nxt > ix->n - 1) nxt = ix->n;
    uint64_t skip = nxt - position - 1;
    for (uint64_t s = 0; s < skip; s++) row = lf(ix, row, NULL);
    ix->c_lf++;                                              /* the final RetrieveCharacter :2263 */
    kmer_of(ix, ix->bwt[row], out);
}

/* ------------------------------------------------------------------ search */

typedef struct { uint64_t *v; size_t n, cap; } vec64;
static void push(vec64 *a, uint64_t x) { if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 64; a->v = (uint64_t *)realloc(a->v, 8 * a->cap); } a->v[a->n++] = x; }
static int cmp64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }

typedef struct { int fixed; int64_t code; uint8_t text[16]; } spchar;

/* one found SA interval of one shift: hat verification or direct match
 * (EFMIndex::findWholePatternMatches :1862-1899, locateOccurrences(Match*) :2304-2323) */
static void emit_interval(orc_index *ix, uint64_t sp, uint64_t ep, int hat, const spchar *last, uint32_t L, uint32_t sa,
                          int mode, uint64_t *count, vec64 *pos) {
    if (!hat) {
        *count += ep - sp + 1;
        if (mode == 1) for (uint64_t r = sp; r <= ep; r++) push(pos, row_position(ix, r) * ix->k + sa + ix->initial_n);
        return;
    }
    uint8_t found[16];
    for (uint64_t r = sp; r <= ep; r++) {
        uint64_t tp = row_position(ix, r);
        super_char_at(ix, tp + L - 1, found);
        int ok = 1;
        for (uint32_t j = 0; j < ix->k && ok; j++) if (last->text[j] != '?' && last->text[j] != found[j]) ok = 0;
        if (ok) { *count += 1; if (mode == 1) push(pos, tp * ix->k + sa + ix->initial_n); }
    }
}

int64_t orc_search(orc_index *ix, const uint8_t *pat, uint32_t m, int mode, uint64_t *count_out, uint64_t *positions, uint64_t pos_cap) {
    uint64_t count = 0; vec64 pos = {0, 0, 0};
    uint32_t k = ix->k, no = ix->n_sym;
    int valid = m > 0;                                         /* the reference reads out of bounds on an empty pattern */
    for (uint32_t i = 0; i < m && valid; i++) if (sym_index(ix, pat[i]) >= no) valid = 0;   /* EFMIndex.cpp:1702-1717 */
    if (mode == 1 && ix->mrp == 0) { *count_out = 0; return -1; }   /* :2305-2306 */
    spchar *sc = (spchar *)malloc(sizeof(spchar) * ((m + k) / k + 2));
    for (uint32_t sa = 0; sa < k && valid; sa++) {
        /* EFMIndex::computeSuperPatterns (:2055-2207) */
        uint32_t L = (m + sa + k - 1) / k;
        int dropped = 0;
        for (uint32_t c = 0; c < L && !dropped; c++) {
            int64_t start = (int64_t)c * k - sa, end = (int64_t)(c + 1) * k - sa - 1;
            uint32_t prefix = 0, suffix = 0; int64_t fs = start, fe = end;
            if (start < 0) { prefix = (uint32_t)(-start); fs = 0; }
            if (end > (int64_t)m - 1) { suffix = (uint32_t)(end - (int64_t)m + 1); fe = (int64_t)m - 1; }
            uint64_t fv = 0, w = 1;
            for (uint32_t e = 0; e < suffix; e++) w *= no;
            for (int64_t i = fe; i >= fs; i--) { fv += sym_index(ix, pat[i]) * w; w *= no; }
            spchar *s = &sc[c];
            if (prefix || suffix) {
                s->fixed = 0; s->code = -1;
                uint32_t q = 0;
                for (uint32_t i = 0; i < prefix; i++) s->text[q++] = '?';
                for (int64_t i = fs; i <= fe; i++) s->text[q++] = pat[i];
                for (uint32_t i = 0; i < suffix; i++) s->text[q++] = '?';
            } else {
                s->fixed = 1;
                s->code = ix->compact_of_scr[ix->perm[fv]];    /* :2186-2191 */
                if (s->code < 0) dropped = 1;                  /* :2194-2197 */
            }
        }
        if (dropped) continue;
        /* start interval selection, EFMIndex::exactMatch(string) :1728-1748 */
        int hat = 0; uint32_t l; int64_t a;
        if (sc[L - 1].fixed) { a = sc[L - 1].code; l = L; }
        else if (L > 1 && sc[L - 2].fixed) { hat = 1; a = sc[L - 2].code; l = L - 1; }
        else continue;                                         /* quirk 1: no explicit intervals were ever built */
        /* EFMIndex::exactMatch(CompactAlphabetInterval*) :1901-1963 */
        uint64_t sp = ix->C[a], ep = ix->C[a + 1] - 1;
        const spchar *cur = &sc[l - 1];
        int64_t i = (int64_t)l - 2;
        while ((int64_t)sp <= (int64_t)ep && i >= 0) {
            cur = &sc[i];
            if (cur->fixed) backward_step(ix, (uint32_t)cur->code, &sp, &ep);
            else {
                for (uint64_t row = sp; row <= ep; row++) {    /* backwardStepIfMatch :2011-2041 */
                    uint8_t km[16]; uint32_t code;
                    uint64_t nxt = lf(ix, row, &code);
                    kmer_of(ix, code, km);
                    int ok = 1;
                    for (uint32_t j = 0; j < k && ok; j++) if (cur->text[j] != '?' && cur->text[j] != km[j]) ok = 0;
                    if (ok) emit_interval(ix, nxt, nxt, hat, &sc[L - 1], L, sa, mode, &count, &pos);
                }
            }
            i--;
        }
        if ((cur->fixed || l == 1) && (int64_t)ep >= (int64_t)sp) emit_interval(ix, sp, ep, hat, &sc[L - 1], L, sa, mode, &count, &pos);
    }
    free(sc);
    *count_out = count;
    /* EFMIndex::locateOccurrences(vector<Match*>*) :2281-2301: sort + adjacent dedupe */
    size_t w = 0;
    if (pos.n) {
        qsort(pos.v, pos.n, 8, cmp64);
        for (size_t i = 0; i < pos.n; i++) if (i == 0 || pos.v[i] != pos.v[i - 1]) pos.v[w++] = pos.v[i];
        for (size_t i = 0; i < w && i < pos_cap; i++) positions[i] = pos.v[i];
    }
    free(pos.v);
    return (int64_t)w;
}

/* header views for the host-parse tests: C[0..A] and the terminators */
void orc_get_C(const orc_index *ix, uint64_t *out) { for (uint32_t i = 0; i <= ix->A; i++) out[i] = ix->C[i]; }
void orc_get_terminators(const orc_index *ix, int64_t *out) { for (uint64_t i = 0; i < ix->n_term; i++) out[i] = ix->term[i]; }

/* EFMIndex::extractSnippet(from, to) (EFMIndex.cpp:2325-2452), character by character through super_char_at (the reference walks LF
 * once from the marked row after stTo; the k-mers read are the same). Returns the length (may exceed cap), -1 without marked rows. */
int64_t orc_extract(orc_index *ix, uint64_t from, uint64_t to, uint8_t *out, uint64_t cap) {
    if (ix->mrp == 0) return -1;
    from -= ix->initial_n; to -= ix->initial_n;
    if (from > to) return 0;
    uint64_t k = ix->k, st_from = from / k, off_from = from % k, st_to = to / k, off_to = to % k;
    if (st_to > ix->n - 2) st_to = ix->n - 2;
    if (st_from > st_to || (st_from == st_to && off_from > off_to)) return 0;
    uint64_t count = st_to - st_from + 1, keep_lf = ix->c_lf, w = 0;
    for (uint64_t c = 0; c < count; c++) {
        uint8_t km[16];
        super_char_at(ix, st_from + c, km);
        uint64_t a = 0, b = k;                                   /* [a,b) of this k-mer */
        if (count == 1) { a = off_from; b = off_from + off_to + 1; if (b > k) b = k; }   /* substr(pos, LENGTH) quirk, :2441 */
        else if (c == 0) a = off_from;
        else if (c == count - 1) b = off_to + 1;
        for (uint64_t j = a; j < b; j++) { if (w < cap) out[w] = km[j]; w++; }
    }
    ix->c_lf = keep_lf;
    return (int64_t)w;
}

/* ---- per-row views of the decoded index, for the kernel-level parity tests ---- */
void orc_get_bwt(const orc_index *ix, uint64_t row0, uint64_t count, uint32_t *out) {
    for (uint64_t i = 0; i < count; i++) out[i] = row0 + i < ix->n ? ix->bwt[row0 + i] : 0xFFFFFFFFu;
}
void orc_get_marks(const orc_index *ix, uint64_t row0, uint64_t count, uint64_t *out) {
    for (uint64_t i = 0; i < count; i++) out[i] = (row0 + i < ix->n && ix->marked[row0 + i]) ? ix->mark_val[row0 + i] : ~0ull;
}
void orc_lf_rows(orc_index *ix, const uint64_t *rows, uint64_t count, uint64_t *out) {
    uint64_t keep = ix->c_lf;
    for (uint64_t i = 0; i < count; i++) out[i] = rows[i] < ix->n ? lf(ix, rows[i], NULL) : ~0ull;
    ix->c_lf = keep;
}
/* C[c] + occ(c,row): the quantity EFMIndex::backwardStep adds up (EFMIndex.cpp:1966-2001) */
void orc_rank_rows(orc_index *ix, const uint32_t *codes, const uint64_t *rows, uint64_t count, uint64_t *out) {
    uint64_t keep = ix->c_rank;
    for (uint64_t i = 0; i < count; i++)
        out[i] = (codes[i] < ix->A && rows[i] < ix->n) ? ix->C[codes[i]] + occ_upto(ix, codes[i], rows[i]) : ~0ull;
    ix->c_rank = keep;
}

void orc_map_position(const orc_index *ix, uint64_t p, uint32_t *seq, uint64_t *off) {
    int64_t s = -1;
    for (uint64_t i = 0; i < ix->n_term && s == -1; i++) if ((uint64_t)ix->term[i] > p) s = (int64_t)i;
    *seq = (uint32_t)s;
    *off = p - (s > 0 ? (uint64_t)ix->term[s - 1] + ix->k : 0);
}

/* ------------------------------------------------------------------ build side: bucket text encoder */

/* What the builder wrote for bucket bn: [0] length, [1] bucketAlphaSize, [2] isOdd, [3] isFirstInSuperBucket, [4] compressedLength
 * (0 for a single-character bucket: Bucket.cpp:166-170 writes no text), [5] bit offset of the text in the file, [6] bits per text symbol */
void orc_bucket_info(const orc_index *ix, uint64_t bn, uint64_t out[8]) {
    const bucket_t *b = &ix->bk[bn];
    out[0] = b->length; out[1] = b->alpha; out[2] = (uint64_t)b->is_odd; out[3] = (uint64_t)b->is_first;
    out[4] = b->alpha > 1 ? b->clen : 0; out[5] = b->text_bit; out[6] = b->tbits; out[7] = 0;
}

/* The rows of bucket bn as the builder saw them when it encoded the bucket: sbwt[] after the super-bucket and the bucket
 * remapping (EFMIndex.cpp:841-858: SuperBucket::charactersRemappings, then Bucket::charactersRemappings), in row order.
 * Both remappings are order preserving, so a bucket code is the rank of the row's compact code among the bucket's symbols. */
void orc_bucket_codes(const orc_index *ix, uint64_t bn, uint32_t *out) {
    const bucket_t *b = &ix->bk[bn];
    uint64_t sn = bn / ix->per;
    uint32_t A = ix->A;
    int64_t *code_of = (int64_t *)malloc(sizeof(int64_t) * A);
    for (uint32_t j = 0; j < A; j++) code_of[j] = -1;
    for (uint32_t j = 0, sbc = 0, c = 0; j < A; j++) {
        if (!ix->sb_bits[sn * A + j]) continue;
        if (b->bits[sbc]) code_of[j] = c++;
        sbc++;
    }
    const uint32_t *bwt = ix->bwt + bn * ix->b_size;
    for (uint64_t i = 0; i < b->length; i++) out[i] = (uint32_t)code_of[bwt[i]];
    free(code_of);
}

/* Bucket::encodeText (Bucket.cpp:442-536): single-pass move-to-front, RLE0 of the runs of rank 0 and poly-alphabetic
 * substitution with the bucket's key stream (EFMIndex::getCryptoParams :1668 -> computeBucketKeyStream(bucketNumber, 0,
 * length - 1, bucketAlphaSize)). codes: the bucket's rows as bucket codes (orc_bucket_codes); odd buckets are encoded from
 * their last row to their first (:446-448). seq_out: `length` entries; returns EncodingResult::realLength. */
uint64_t orc_encode_bucket(const uint8_t key32[32], uint64_t bucket_number, const uint32_t *codes, uint64_t length, uint32_t alpha,
                           int is_odd, uint32_t *seq_out) {
    uint32_t *key = (uint32_t *)malloc(sizeof(uint32_t) * (length + 1));
    orc_keystream_ints(key32, bucket_number, alpha, 0, length, key);
    uint32_t *text = (uint32_t *)malloc(sizeof(uint32_t) * (length + 1));
    for (uint64_t j = 0; j < length; j++) text[is_odd ? length - j - 1 : j] = codes[j];
    uint32_t *mtf = (uint32_t *)malloc(sizeof(uint32_t) * alpha);              /* ArrayMTFList(bucketAlphaSize) */
    for (uint32_t i = 0; i < alpha; i++) mtf[i] = i;
    const uint64_t mod = (uint64_t)alpha + 1;
    uint64_t n_out = 0, i = 0;
    while (i < length) {
        uint32_t pos = 0;
        while (mtf[pos] != text[i]) pos++;                                     /* ArrayMTFList::indexOf */
        if (pos > 0) {
            uint32_t e = mtf[pos];
            memmove(mtf + 1, mtf, sizeof(uint32_t) * pos);                     /* moveToFront */
            mtf[0] = e;
            seq_out[n_out] = (uint32_t)(((uint64_t)pos + 1 + key[n_out % length]) % mod);   /* :480-481, keyStartOffset = 0 */
            n_out++;
            i++;
        } else {
            uint64_t m = 1;                                                     /* run of rank 0 (:499-506) */
            while (i + m < length && text[i + m] == mtf[0]) m++;
            uint32_t nb = 0;                                                    /* floor(log2(m + 1)) (:508) */
            while (((m + 1) >> (nb + 1)) != 0) nb++;
            const uint64_t N = m + 1 - (1ull << nb);
            for (int64_t j = (int64_t)nb - 1; j >= 0; j--) {                    /* digits, most significant first (:512-519) */
                const uint64_t digit = (N >> j) & 1u;
                seq_out[n_out] = (uint32_t)((digit + key[n_out % length]) % mod);
                n_out++;
            }
            i += m;
        }
    }
    free(key); free(text); free(mtf);
    return n_out;
}

/* the symbol loop of Bucket::writeText (Bucket.cpp:232-246) + BitWriter::flush: n symbols of `bits` bits each, most significant
 * bit first, the last byte padded with zeros. out: (n * bits + 7) / 8 bytes. */
void orc_pack_sequence(const uint32_t *seq, uint64_t n, uint32_t bits, uint8_t *out) {
    memset(out, 0, (size_t)((n * bits + 7) / 8));
    uint64_t bit = 0;
    for (uint64_t i = 0; i < n; i++)
        for (int32_t b = (int32_t)bits - 1; b >= 0; b--, bit++)
            if ((seq[i] >> b) & 1u) out[bit >> 3] |= (uint8_t)(0x80u >> (bit & 7));
}

/* The inverse, for round-trip properties: Bucket::readTextUntilTo (Bucket.cpp:288-403) over a sequence held in memory instead of
 * the file's bit stream — invertPoly (:406), RLE0^-1 (:377), ArrayMTFList::moveToFront (ArrayMTFList.cpp:32), odd buckets written
 * back last row first. Returns the number of rows produced (== length for a well-formed sequence). */
uint64_t orc_decode_sequence(const uint8_t key32[32], uint64_t bucket_number, const uint32_t *seq, uint64_t clen, uint64_t length,
                             uint32_t alpha, int is_odd, uint32_t *codes_out) {
    uint32_t *key = (uint32_t *)malloc(sizeof(uint32_t) * (clen + 1));
    orc_keystream_ints(key32, bucket_number, alpha, 0, clen, key);
    uint32_t *mtf = (uint32_t *)malloc(sizeof(uint32_t) * alpha);
    for (uint32_t i = 0; i < alpha; i++) mtf[i] = i;
    const int64_t mod = (int64_t)alpha + 1;
    uint64_t p = 0, i = 0;
    while (p < length && i < clen) {
        int64_t c = (((int64_t)seq[i] - (int64_t)key[i]) % mod + mod) % mod;
        if (c > 1) {
            uint32_t pos = (uint32_t)(c - 1), e = mtf[pos];
            memmove(mtf + 1, mtf, sizeof(uint32_t) * pos);
            mtf[0] = e;
            codes_out[is_odd ? length - p - 1 : p] = e;
            p++; i++;
        } else {
            uint64_t x = 0, N = 0;
            while (i + x < clen) {
                int64_t d = (((int64_t)seq[i + x] - (int64_t)key[i + x]) % mod + mod) % mod;
                if (d >= 2) break;
                N = 2 * N + (uint64_t)d; x++;
            }
            uint64_t m = (1ull << x) + N - 1;
            for (uint64_t j = 0; j < m && p < length; j++) { codes_out[is_odd ? length - p - 1 : p] = mtf[0]; p++; }
            i += x;
        }
    }
    free(key); free(mtf);
    return p;
}

