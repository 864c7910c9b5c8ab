// Minimal stand-in for sdsl::int_vector<0> (test infrastructure only): resize / width / operator[].
#pragma once
#include <vector>
#include <cstdint>
namespace sdsl {
template <unsigned char W = 0>
class int_vector {
    std::vector<uint64_t> v_;
    unsigned char w_;
public:
    int_vector() : w_(64) {}
    void resize(uint64_t n) { v_.resize(n, 0); }
    void width(unsigned char w) { w_ = w; }
    unsigned char width() const { return w_; }
    uint64_t size() const { return v_.size(); }
    uint64_t &operator[](uint64_t i) { return v_[i]; }
    const uint64_t &operator[](uint64_t i) const { return v_[i]; }
};
}  // namespace sdsl
