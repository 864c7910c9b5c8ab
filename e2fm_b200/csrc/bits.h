// bits.h — bit-field helpers shared by the host loader and the device kernels.
// The E2FM file format is an MSB-first bit stream with big-endian multi-bit fields that may start at
// any bit offset (reference: MemoryBitReader::read, MemoryBitReader.cpp:126-174).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define E2FM_HD __host__ __device__ __forceinline__
#else
#define E2FM_HD inline
#endif

namespace e2fm {

// Utils::int_log2 (Utils.cpp:58-69): bits needed to represent u, with int_log2(0) == int_log2(1) == 1.
E2FM_HD uint32_t int_log2(uint64_t u) {
    uint32_t i = 1;
    uint64_t r = 1;
    while (i <= 64 && r < u) { r = 2 * r + 1; i++; }
    return i;
}

// n-bit (1..57) big-endian field at absolute bit offset `bit` of a byte buffer. The caller guarantees
// 8 readable bytes from bit/8 (the loader pads the file image, like MemoryBitReader::loadDataFromFile).
E2FM_HD uint64_t read_bits(const uint8_t *p, uint64_t bit, uint32_t n) {
    const uint8_t *q = p + (bit >> 3);
    uint64_t w = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) w = (w << 8) | q[i];
    uint32_t head = (uint32_t)(bit & 7);
    return (w >> (64 - head - n)) & ((1ull << n) - 1);
}

// Sequential reader over the same stream (host loader and the per-bucket header parse on the device).
struct BitCursor {
    const uint8_t *p;
    uint64_t bit;
    E2FM_HD uint64_t get(uint32_t n) { uint64_t v = read_bits(p, bit, n); bit += n; return v; }
    E2FM_HD uint64_t get64() { uint64_t hi = get(56); return (hi << 8) | get(8); }   // MemoryBitReader::getInt64 :51
    E2FM_HD void align() { bit = (bit + 7) & ~7ull; }                                // gotoNextByteStart :82
    // BitReader::integerDecode (BitReader.cpp:104-133)
    E2FM_HD uint64_t idec(uint32_t head_bits) {
        uint64_t i = get(head_bits);
        if (i < 2) return i;
        if (i == 2) return 2 + get(1);
        i--;
        return (1ull << i) + get((uint32_t)i);
    }
};

}  // namespace e2fm
