// Minimal stand-in for the SeqAn types the reference's FASTA loader touches (test infrastructure only).
// Sequences are plain upper-case ACGT in our synthetic inputs, so IUPAC normalisation is the identity.
#pragma once
#include <string>
#include <vector>
#include <fstream>
#include <ostream>
#include <stdexcept>
namespace seqan {
struct CharString : public std::string { CharString() {} CharString(const std::string &s) : std::string(s) {} };
struct IupacString : public std::string { IupacString() {} IupacString(const std::string &s) : std::string(s) {} };
template <typename T> struct StringSet : public std::vector<T> {};
template <typename T> inline std::size_t length(const StringSet<T> &s) { return s.size(); }
inline std::size_t length(const IupacString &s) { return s.size(); }
inline std::size_t length(const CharString &s) { return s.size(); }
inline const char *toCString(const CharString &s) { return s.c_str(); }
struct SeqFileIn {
    std::ifstream in;
    explicit SeqFileIn(const char *path) : in(path) { if (!in) throw std::runtime_error(std::string("cannot open FASTA ") + path); }
};
inline void readRecords(StringSet<CharString> &ids, StringSet<IupacString> &seqs, SeqFileIn &f) {
    std::string line, cur; bool have = false;
    while (std::getline(f.in, line)) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        if (!line.empty() && line[0] == '>') {
            if (have) seqs.push_back(IupacString(cur));
            ids.push_back(CharString(line.substr(1))); cur.clear(); have = true;
        } else if (have) cur += line;
    }
    if (have) seqs.push_back(IupacString(cur));
}
}  // namespace seqan
