// internal.h — what the translation units of libe2fm_b200.so share beyond the public C ABI (not exported in include/e2fm_b200.h).
#pragma once
#include <cstdint>

struct e2fm_index;
struct e2fm_result;

extern "C" {
// CUDA device ordinal of an index
int e2fm_internal_device(const e2fm_index *ix);
// wrap device arrays (allocated with cudaMallocAsync on the index's stream) as a result the public fetch / free calls understand;
// takes ownership of the arrays. csr / pos / seq / seqoff may be null when !located.
e2fm_result *e2fm_internal_make_result(e2fm_index *ix, uint64_t n, uint64_t total, bool located, unsigned long long *d_counts,
                                       uint64_t *d_csr, uint64_t *d_pos, uint32_t *d_seq, uint64_t *d_seqoff);
// sets the thread-local text e2fm_last_error() returns
void e2fm_internal_set_error(const char *msg);
}
