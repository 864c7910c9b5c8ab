// Minimal stand-in for boost::dynamic_bitset<> (test infrastructure only).
// Written for this repo so that the UNMODIFIED reference sources under /root/reference
// compile without Boost. Only the members the reference's search/build path uses.
#pragma once
#include <vector>
#include <cstdint>
#include <cstddef>
#include <ostream>
namespace boost {
template <typename Block = unsigned long>
class dynamic_bitset {
    std::vector<unsigned char> b_;
public:
    typedef std::size_t size_type;
    class reference {
        unsigned char &r_;
    public:
        explicit reference(unsigned char &r) : r_(r) {}
        reference &operator=(bool v) { r_ = v ? 1 : 0; return *this; }
        reference &operator=(const reference &o) { r_ = o.r_; return *this; }
        operator bool() const { return r_ != 0; }
        bool operator~() const { return r_ == 0; }
    };
    dynamic_bitset() {}
    explicit dynamic_bitset(size_type nbits, unsigned long value = 0) : b_(nbits, 0) {
        for (size_type i = 0; i < nbits && i < 8 * sizeof(unsigned long); i++) b_[i] = (value >> i) & 1ul;
    }
    void resize(size_type n, bool v = false) { b_.resize(n, v ? 1 : 0); }
    void clear() { b_.clear(); }
    size_type size() const { return b_.size(); }
    bool empty() const { return b_.empty(); }
    dynamic_bitset &set(size_type i, bool v = true) { b_[i] = v ? 1 : 0; return *this; }
    dynamic_bitset &reset(size_type i) { b_[i] = 0; return *this; }
    bool test(size_type i) const { return b_[i] != 0; }
    bool operator[](size_type i) const { return b_[i] != 0; }
    reference operator[](size_type i) { return reference(b_[i]); }
    size_type count() const { size_type c = 0; for (size_type i = 0; i < b_.size(); i++) c += b_[i]; return c; }
    unsigned long to_ulong() const {
        unsigned long v = 0;
        for (size_type i = 0; i < b_.size() && i < 8 * sizeof(unsigned long); i++) if (b_[i]) v |= (1ul << i);
        return v;
    }
};
template <typename B>
std::ostream &operator<<(std::ostream &os, const dynamic_bitset<B> &b) {
    for (std::size_t i = b.size(); i > 0; i--) os << (b[i - 1] ? '1' : '0');
    return os;
}
}  // namespace boost
