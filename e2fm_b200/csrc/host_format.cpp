// host_format.cpp — parse the global header of an E2FM index and derive the load-time secrets.
//
// What is read here (small, sequential, key-dependent) — reference EFMIndex::readHeader
// (EFMIndex.cpp:1150-1329), EFMCollection::readHeader (EFMCollection.cpp:289-308):
//   magic, original symbols, k, lengths, bucket sizes, the 5 bookmarks, the compact-alphabet bitset,
//   the encrypted cumulative frequencies C[], super-bucket / bucket start tables, terminators, FASTA
//   headers. Everything per-bucket (alphabet bitsets, marked-row bitmaps, encrypted text, marked-row
//   values) is parsed and decoded by the CUDA kernels in index_build.cu.
#include "host_format.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "bits.h"
#include "salsa20.h"

namespace e2fm {

// ---------------------------------------------------------------- KeyStream

KeyStream::KeyStream(uint32_t buffer_bytes, const uint8_t key32[32], uint64_t nonce) : nonce_(nonce) {
    for (int i = 0; i < 8; i++)
        key_[i] = (uint32_t)key32[4 * i] | ((uint32_t)key32[4 * i + 1] << 8) | ((uint32_t)key32[4 * i + 2] << 16) |
                  ((uint32_t)key32[4 * i + 3] << 24);
    buf_.resize((buffer_bytes + 63) / 64 * 64);
}

// RandomGenerator::populateKeyStreamBuffer(bytes) (RandomGenerator.cpp:157-172): the block counter is reset
// to `counter` (0 for these generators) on every refill, so the stream repeats every buffer_bytes bytes.
void KeyStream::refill() {
    uint32_t w[16];
    for (size_t b = 0; b < buf_.size() / 64; b++) {
        salsa20_block(key_, nonce_, (uint64_t)counter_ + b, w);
        for (int i = 0; i < 16; i++) {
            buf_[64 * b + 4 * i + 0] = (uint8_t)w[i];
            buf_[64 * b + 4 * i + 1] = (uint8_t)(w[i] >> 8);
            buf_[64 * b + 4 * i + 2] = (uint8_t)(w[i] >> 16);
            buf_[64 * b + 4 * i + 3] = (uint8_t)(w[i] >> 24);
        }
    }
    live_bytes_ = (uint32_t)buf_.size();
}

uint32_t KeyStream::next(uint32_t n) {
    uint32_t live = live_bits_, buf = bits_buffer_;   // 32-bit on purpose (RandomGenerator.h:35)
    if (live < n) {
        do {
            if (live_bytes_ == 0) refill();
            uint32_t c = buf_[buf_.size() - live_bytes_];
            live_bytes_--;
            buf = (buf << 8) | c;
            live += 8;
        } while (live < n);
        bits_buffer_ = buf;
    }
    live_bits_ = live - n;
    uint32_t sh = live - n;
    uint32_t v = sh >= 32 ? 0u : (buf >> sh);
    return n >= 32 ? v : (v & ((1u << n) - 1));
}

uint32_t KeyStream::next_below(uint32_t n) {
    uint32_t bits = int_log2((uint64_t)(uint32_t)(n - 1));
    uint32_t v;
    do v = next(bits); while (v > n - 1);
    return v;
}

// ---------------------------------------------------------------- header

static uint32_t symbol_index(const HostImage &img, uint8_t c) {
    for (uint32_t i = 0; i < img.n_sym; i++)
        if (img.syms[i] == c) return i;
    return img.n_sym;
}

// "1-bit flag, unaligned i64 size, bits" (EFMIndex.cpp:1200-1215). The flag selects the RLE form
// (BitSetRLEEncoder), which the reference never writes for alphabets (tryToStoreCompressedBitmaps=false,
// EFMIndex.h:406) but can read.
// Tail-safe reader for the host parse: the caller's buffer is NOT padded (unlike the device copy), so a field that starts
// within 8 bytes of the end — or, in a truncated or corrupt file, beyond it — reads zeros instead of foreign memory.
struct HostCursor {
    const uint8_t *p;
    uint64_t size;     // bytes
    uint64_t bit;
    uint64_t get(uint32_t n) {
        const uint64_t byte = bit >> 3;
        uint64_t v;
        if (byte + 8 <= size) v = read_bits(p, bit, n);
        else {
            uint8_t tmp[8] = {0};
            for (uint64_t i = 0; i < 8 && byte + i < size; i++) tmp[i] = p[byte + i];
            v = read_bits(tmp, bit & 7, n);
        }
        bit += n;
        return v;
    }
    uint64_t get64() { uint64_t hi = get(56); return (hi << 8) | get(8); }   // MemoryBitReader::getInt64 :51
    void align() { bit = (bit + 7) & ~7ull; }                                // gotoNextByteStart :82
};

static std::vector<uint8_t> read_bitset(HostCursor &r, uint64_t max_bits) {
    bool compressed = r.get(1) == 1;
    uint64_t sz = r.get64();
    // the stored form is never longer than the bitset it stands for (max_bits, known from the header) plus the RLE's own
    // overhead, and it must lie inside the file: checked BEFORE anything is allocated
    if (sz > 2 * max_bits + 64 || r.bit + sz > r.size * 8) throw std::runtime_error("corrupt index: bitset larger than the file");
    std::vector<uint8_t> raw(sz);
    for (uint64_t i = 0; i < sz; i++) raw[i] = (uint8_t)r.get(1);
    if (!compressed) return raw;
    std::vector<uint8_t> out;   // BitSetRLEEncoder::decompress (BitSetRLEEncoder.cpp:67-135)
    uint64_t N = 0;
    uint32_t x = 0;
    bool pending = false;
    auto flush = [&]() {
        if (x > 0) {
            if (x >= 40 || out.size() + ((1ull << x) + N - 1) > max_bits + 1) throw std::runtime_error("corrupt index: bitset run past its size");
            out.insert(out.end(), (size_t)((1ull << x) + N - 1), 0);
        }
        out.push_back(1);
        N = 0;
        x = 0;
    };
    for (uint64_t i = 0; i < sz; i++) {
        if (!pending) {
            if (raw[i] == 0) { N = 2 * N; x++; } else pending = true;
        } else {
            if (raw[i] == 0) { N = 2 * N + 1; x++; } else flush();
            pending = false;
        }
    }
    if (x > 0) flush();
    return out;
}

void parse_index_header(const uint8_t *file, size_t size, const uint8_t key64[64], HostImage &img) {
    if (size < 2 || file[0] != 'E' || file[1] != 'G')
        throw std::runtime_error("File type is not Encrypted and Compressed Genomic Index");   // EFMIndex.cpp:1155
    if (size < 64) throw std::runtime_error("corrupt index: truncated header");
    const uint64_t file_bits = (uint64_t)size * 8;
    HostCursor r{file, (uint64_t)size, 16};
    img.n_sym = (uint32_t)r.get(8);
    if (img.n_sym == 0 || 3 + (uint64_t)img.n_sym + 100 > size) throw std::runtime_error("corrupt index: bad symbol table");
    for (uint32_t i = 0; i < img.n_sym; i++) img.syms[i] = (uint8_t)r.get(8);
    img.k = (uint32_t)r.get(8);
    img.n = r.get64();
    img.initial_n = r.get64();
    img.final_n = r.get64();
    img.last_pos = r.get64();
    img.sb_size = r.get64();
    img.b_size = r.get64();
    img.mrp = (uint32_t)r.get(8);
    (void)r.get64();   // maximumBucketAlphaSize: never initialised by the builder (EFMIndex.h:395, EFMIndex.cpp:864)
    uint64_t c_start = r.get64(), mrb_start = r.get64(), sbpos_start = r.get64(), bpos_start = r.get64();
    uint64_t fasta_start = r.get64();   // EFMCollection::readSuspendedInformations (EFMCollection.cpp:242)
    if (img.k == 0 || img.k > 12 || img.n == 0 || img.b_size == 0 || img.sb_size == 0 || img.sb_size % img.b_size != 0)
        throw std::runtime_error("corrupt index: bad geometry");
    if (c_start >= size || mrb_start > size || sbpos_start >= size || bpos_start >= size || fasta_start > size)
        throw std::runtime_error("corrupt index: bookmark outside the file");
    // rows are 32-bit on the device and bit 31 of an interval length flags a long wildcard interval (search.cu, IV_LONG)
    if (img.n >= (1ull << 31)) throw std::runtime_error("unsupported index: super-text longer than 2^31-1 rows per device shard");
    img.B = (img.n + img.b_size - 1) / img.b_size;
    img.S = (img.n + img.sb_size - 1) / img.sb_size;

    uint64_t card = 1;
    for (uint32_t i = 0; i < img.k; i++) {
        card *= img.n_sym;
        if (card > (1ull << 24)) throw std::runtime_error("unsupported index: super-alphabet larger than 2^24");
    }
    img.sigma_k = (uint32_t)card;

    // key split (EncryptionManager::initSALSA20Keys, EncryptionManager.cpp:47-55)
    const uint8_t *mono = key64, *poly = key64 + 32;
    for (int i = 0; i < 8; i++)
        img.poly_key[i] = (uint32_t)poly[4 * i] | ((uint32_t)poly[4 * i + 1] << 8) | ((uint32_t)poly[4 * i + 2] << 16) |
                          ((uint32_t)poly[4 * i + 3] << 24);

    // scrambling permutation: Fisher-Yates that never moves element 0 (EncryptionManager.cpp:92-117)
    std::vector<uint32_t> perm(card);
    for (uint32_t i = 0; i < card; i++) perm[i] = i;
    {
        KeyStream ks(16384, mono, 0);
        for (uint32_t i = (uint32_t)card; i > 2; i--) {
            uint32_t j;
            do j = ks.next_below(i); while (j == 0);
            std::swap(perm[i - 1], perm[j]);
        }
    }
    // the all-'N' k-mer is forced onto the last code (ScrambledSuperAlphabet.cpp:23-31)
    uint32_t n_idx = symbol_index(img, 'N');
    if (n_idx < img.n_sym) {
        uint32_t all_n = img.sigma_k - 1, nat = 0;
        for (uint32_t i = 0; i < img.k; i++) nat = nat * img.n_sym + n_idx;
        uint32_t rep = perm[nat];
        uint32_t pos = (uint32_t)(std::find(perm.begin(), perm.end(), all_n) - perm.begin());
        perm[pos] = rep;
        perm[nat] = all_n;
    }

    // compact alphabet (EFMIndex.cpp:1200-1239): compact code = rank of the scrambled code in the bitset
    std::vector<uint8_t> abits = read_bitset(r, card);
    if (abits.size() != card) throw std::runtime_error("corrupt index or wrong alphabet: bitset size mismatch");
    std::vector<int32_t> compact_of_scr(card, -1);
    img.A = 0;
    for (uint32_t i = 0; i < card; i++)
        if (abits[i]) compact_of_scr[i] = (int32_t)img.A++;
    if (img.A == 0 || img.A > 0xFFFE) throw std::runtime_error("unsupported index: compact alphabet must have 1..65534 symbols");
    img.compact_of_natural.assign(card, 0xFFFF);
    img.natural_of_compact.assign(img.A, 0);
    for (uint32_t nat = 0; nat < card; nat++) {
        int32_t c = compact_of_scr[perm[nat]];
        if (c >= 0) { img.compact_of_natural[nat] = (uint16_t)c; img.natural_of_compact[c] = nat; }
    }

    // C[] (EFMIndex.cpp:1241-1272): inverse transposition, then inverse poly-alphabetic substitution
    r.align();
    if (r.bit != c_start * 8) throw std::runtime_error("corrupt index: characterOccurrencesStart mismatch");
    uint32_t nb = (uint32_t)r.get(8);
    if (nb == 0 || nb > 57) throw std::runtime_error("corrupt index: bad counter width");
    const uint32_t A1 = img.A + 1;
    if (r.bit + (uint64_t)A1 * nb > file_bits) throw std::runtime_error("corrupt index: truncated C[]");
    const uint64_t maxf = (1ull << nb) - 1;
    std::vector<uint32_t> fperm(A1);
    for (uint32_t i = 0; i < A1; i++) fperm[i] = i;
    {
        KeyStream ks(16384, mono, 1);   // computeFrequenciesPermutation (EncryptionManager.cpp:170-180)
        for (uint32_t i = A1; i > 1; i--) {
            uint32_t j;
            do j = ks.next_below(i); while (j == 0);
            std::swap(fperm[i - 1], fperm[j]);
        }
    }
    std::vector<uint64_t> enc(A1);
    for (uint32_t i = 0; i < A1; i++) enc[fperm[i]] = r.get(nb);
    img.C.resize(A1);
    {
        KeyStream ks(16384, poly, 2);   // computeCumulativeFreqsKeyStream (EncryptionManager.cpp:150-161)
        for (uint32_t i = 0; i < A1; i++) {
            uint64_t key = ks.next_below((uint32_t)(maxf + 1));   // nextInt takes uint32_t: the bound truncates
            img.C[i] = (enc[i] - key) & maxf;
        }
    }
    // a wrong key decrypts C[] to garbage: it must be a non-decreasing prefix-sum ending at n
    if (img.C[0] != 0 || img.C[img.A] != img.n) throw std::runtime_error("wrong encryption key (or corrupt index): C[] does not decrypt");
    for (uint32_t i = 0; i < img.A; i++)
        if (img.C[i] > img.C[i + 1]) throw std::runtime_error("wrong encryption key (or corrupt index): C[] does not decrypt");

    // marked rows (EFMIndex.cpp:1276-1287)
    if (img.mrp > 0) {
        img.rate = 100 / img.mrp;
        if (img.rate == 0) throw std::runtime_error("corrupt index: marked rows percentage > 100");
        uint64_t q = img.n / img.rate;
        img.mrn = q + (img.n % img.rate != 0);          // marks actually written: multiples of rate in [0,n)
        img.val_bits = int_log2(q);                      // width used by the builder (EFMIndex.cpp:399)
        if (q >= 1 && int_log2(q - 1) != img.val_bits) img.quirks |= 1;   // the reference reads them with the wrong width (:1280)
        if (img.n % img.rate == 0) img.quirks |= 2;      // the reference reads one junk table entry (:1278)
    }

    img.sb_start.resize(img.S);
    img.b_start.resize(img.B);
    if (sbpos_start + 8 * img.S > size || bpos_start + 8 * img.B + 5 > size) throw std::runtime_error("corrupt index: truncated start tables");
    r.bit = sbpos_start * 8;
    for (uint64_t i = 0; i < img.S; i++) img.sb_start[i] = r.get64();
    r.bit = bpos_start * 8;
    for (uint64_t i = 0; i < img.B; i++) img.b_start[i] = r.get64();
    for (uint64_t i = 0; i < img.S; i++)
        if (img.sb_start[i] + 9 > size) throw std::runtime_error("corrupt index: super-bucket start outside the file");
    img.max_bucket_bytes = 0;
    for (uint64_t i = 0; i < img.B; i++) {
        if (img.b_start[i] + 17 > size || (i > 0 && img.b_start[i] <= img.b_start[i - 1]))
            throw std::runtime_error("corrupt index: bucket start table");
        uint64_t end = i + 1 < img.B ? img.b_start[i + 1] : size;
        if (end > img.b_start[i]) img.max_bucket_bytes = std::max(img.max_bucket_bytes, end - img.b_start[i]);
    }

    // collection part (EFMCollection.cpp:289-303): terminators follow the bucket start table
    uint64_t cnt = r.get(32);
    uint32_t tb = (uint32_t)r.get(8);
    if (tb == 0 || tb > 57 || r.bit + cnt * tb > file_bits) throw std::runtime_error("corrupt index: terminators table");
    img.terminators.resize(cnt);
    int64_t tp = 0;
    for (uint64_t i = 0; i < cnt; i++) { tp += (int64_t)r.get(tb); img.terminators[i] = tp; }
    // FASTA headers (EFMCollection::loadFastaHeaders, EFMCollection.cpp:97-104): i32 size + bytes
    if (fasta_start + 4 <= size) {
        HostCursor h{file, (uint64_t)size, fasta_start * 8};
        uint64_t hs = h.get(32);
        if (fasta_start + 4 + hs <= size) img.fasta_headers.assign((const char *)file + fasta_start + 4, hs);
    }
}

}  // namespace e2fm
