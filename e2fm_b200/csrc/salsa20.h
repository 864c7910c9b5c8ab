// salsa20.h — Salsa20/20 block function for host and device.
// Replaces the reference's salsa20.S (x86-64 SSE2, D. J. Bernstein's amd64-xmm6) on the search path.
// Standard state layout [c0,k0..k3,c1,n0,n1,ctr0,ctr1,c2,k4..k7,c3]; 256-bit key; the 64-bit nonce is the
// bucket number as a little-endian u64 (RandomGenerator.cpp:128), the 64-bit counter is the block index
// (RandomGenerator.cpp:175-179). Output words are little-endian, i.e. byte j of the block is
// (word[j/4] >> 8*(j%4)) & 0xFF.
#pragma once
#include <cstdint>
#include "bits.h"

namespace e2fm {

E2FM_HD uint32_t rotl32(uint32_t v, int c) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(v, v, c);   // one SHF.L.W on the ALU pipe
#else
    return (v << c) | (v >> (32 - c));
#endif
}

#define E2FM_QR(a, b, c, d)      \
    b ^= rotl32(a + d, 7);       \
    c ^= rotl32(b + a, 9);       \
    d ^= rotl32(c + b, 13);      \
    a ^= rotl32(d + c, 18);

// key words are the 32-byte key read as 8 little-endian u32.
E2FM_HD void salsa20_block(const uint32_t key[8], uint64_t nonce, uint64_t counter, uint32_t out[16]) {
    const uint32_t c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;  // "expand 32-byte k"
    uint32_t x0 = c0, x1 = key[0], x2 = key[1], x3 = key[2], x4 = key[3], x5 = c1;
    uint32_t x6 = (uint32_t)nonce, x7 = (uint32_t)(nonce >> 32), x8 = (uint32_t)counter, x9 = (uint32_t)(counter >> 32);
    uint32_t x10 = c2, x11 = key[4], x12 = key[5], x13 = key[6], x14 = key[7], x15 = c3;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        E2FM_QR(x0, x4, x8, x12) E2FM_QR(x5, x9, x13, x1) E2FM_QR(x10, x14, x2, x6) E2FM_QR(x15, x3, x7, x11)
        E2FM_QR(x0, x1, x2, x3) E2FM_QR(x5, x6, x7, x4) E2FM_QR(x10, x11, x8, x9) E2FM_QR(x15, x12, x13, x14)
    }
    out[0] = x0 + c0;        out[1] = x1 + key[0];  out[2] = x2 + key[1];   out[3] = x3 + key[2];
    out[4] = x4 + key[3];    out[5] = x5 + c1;      out[6] = x6 + (uint32_t)nonce;  out[7] = x7 + (uint32_t)(nonce >> 32);
    out[8] = x8 + (uint32_t)counter; out[9] = x9 + (uint32_t)(counter >> 32); out[10] = x10 + c2; out[11] = x11 + key[4];
    out[12] = x12 + key[5];  out[13] = x13 + key[6]; out[14] = x14 + key[7]; out[15] = x15 + c3;
}

}  // namespace e2fm
