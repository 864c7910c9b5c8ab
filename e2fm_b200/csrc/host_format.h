// host_format.h — host-side parse of an E2FM index image (header only; buckets are decoded on the GPU).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace e2fm {

// Key-stream integer source with the reference's exact semantics (RandomGenerator.cpp): MSB-first
// fields of a Salsa20/20 stream, a 32-bit bit buffer, and a byte buffer that is REGENERATED FROM THE
// SAME START COUNTER when exhausted. Used for the load-time secrets only (scrambling permutation,
// C[] permutation and key stream); bucket key streams are generated on the device.
class KeyStream {
public:
    KeyStream(uint32_t buffer_bytes, const uint8_t key32[32], uint64_t nonce);
    uint32_t next(uint32_t n_bits);          // RandomGenerator::next :96
    uint32_t next_below(uint32_t n);         // RandomGenerator::nextInt(n) :69 (rejection sampling)
private:
    void refill();
    uint32_t key_[8];
    uint64_t nonce_;
    uint32_t counter_ = 0;
    std::vector<uint8_t> buf_;
    uint32_t live_bytes_ = 0;
    uint32_t bits_buffer_ = 0, live_bits_ = 0;
};

struct HostImage {
    uint32_t n_sym = 0;
    uint8_t syms[256] = {0};
    uint32_t k = 0;
    uint64_t n = 0, initial_n = 0, final_n = 0, last_pos = 0, sb_size = 0, b_size = 0;
    uint32_t mrp = 0, rate = 0;
    uint32_t sigma_k = 0;                       // n_sym^k
    uint32_t A = 0;                             // compact alphabet size
    std::vector<uint16_t> compact_of_natural;   // [sigma_k] natural k-mer value -> compact code, 0xFFFF if absent
    std::vector<uint32_t> natural_of_compact;   // [A]
    std::vector<uint64_t> C;                    // [A+1]
    uint64_t mrn = 0;                           // markedRowsNumber
    uint32_t val_bits = 0;                      // width of a stored marked-row value
    uint32_t quirks = 0;
    uint64_t S = 0, B = 0;
    std::vector<uint64_t> sb_start, b_start;    // byte offsets
    std::vector<int64_t> terminators;
    std::string fasta_headers;
    uint64_t max_bucket_bytes = 0;              // largest span between consecutive bucket starts
    uint32_t poly_key[8];                       // key bytes 32..63 as LE words
};

// Throws std::runtime_error with the reference's messages where it has one.
void parse_index_header(const uint8_t *file, size_t size, const uint8_t key64[64], HostImage &img);

}  // namespace e2fm
