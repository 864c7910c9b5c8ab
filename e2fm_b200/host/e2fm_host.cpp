// e2fm_host.cpp — see e2fm_host.h. Thin: file I/O, key handling, marshalling to / from the C ABI.
#include "e2fm_host.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <thread>

#include "../../include/e2fm_b200.h"

namespace e2fm {

namespace {
[[noreturn]] void raise_last(const char *what) { throw std::runtime_error(std::string(what) + ": " + e2fm_last_error()); }
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// fn(begin, end) over [0, n) on all host threads (the marshalling of 10^7 std::strings is the host's share of a batch)
template <class F>
void parallel_for(uint64_t n, F fn) {
    unsigned t = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32));
    if (n < 65536) t = 1;
    if (t == 1) { fn((uint64_t)0, n); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < t; i++) th.emplace_back([=]() { fn(n * i / t, n * (i + 1) / t); });
    for (auto &x : th) x.join();
}
}  // namespace

void *EFMIndex::Pinned::ensure(size_t need) {
    if (need > bytes) {
        release();
        const size_t want = need + need / 4 + 4096;
        p = e2fm_host_alloc(want);
        if (!p) throw std::runtime_error(std::string("e2fm_host_alloc: ") + e2fm_last_error());
        bytes = want;
    }
    return p;
}
void EFMIndex::Pinned::release() {
    if (p) e2fm_host_free(p);
    p = nullptr;
    bytes = 0;
}

EFMIndex::EFMIndex()
    : handle_(nullptr), device_(0), numberOfThreads_(1), haveKey_(false), infoBucketSize_(0), infoSuperBucketSize_(0), infoBuckets_(0),
      infoSuperBuckets_(0), infoN_(0), infoK_(0), infoMrp_(0) {
    memset(key_, 0, sizeof(key_));
}

EFMIndex::~EFMIndex() { close(); }

void EFMIndex::setEncryptionKey(u8 *encryptionKey) {
    memcpy(key_, encryptionKey, 64);
    haveKey_ = true;
}

// EncryptionManager::EncryptionManager(string&) (EncryptionManager.cpp:15-36): same checks, same messages
void EFMIndex::setEncryptionKeyFilePath(std::string encryptionKeyFilePath) {
    keyFilePath_ = encryptionKeyFilePath;
    std::ifstream is(encryptionKeyFilePath, std::ifstream::binary);
    if (!is) throw std::runtime_error("Error opening key file " + encryptionKeyFilePath);
    is.seekg(0, is.end);
    uint64_t length = (uint64_t)is.tellg();
    is.seekg(0, is.beg);
    if (length != 64) throw std::runtime_error("File " + encryptionKeyFilePath + "is not a valid key file");
    is.read((char *)key_, 64);
    haveKey_ = true;
}
std::string EFMIndex::getEncryptionKeyFilePath() { return keyFilePath_; }
void EFMIndex::setNumberOfThreads(uint32_t n) { numberOfThreads_ = n; }
uint32_t EFMIndex::getNumberOfThreads() { return numberOfThreads_; }
void EFMIndex::setDevice(int device) { device_ = device; }

void EFMIndex::load(std::string fileName) {
    close();
    if (!haveKey_) throw std::runtime_error("Encryption key not set");
    std::ifstream is(fileName, std::ifstream::binary);
    if (!is) throw std::runtime_error("Error opening index file " + fileName);
    is.seekg(0, is.end);
    size_t size = (size_t)is.tellg();
    is.seekg(0, is.beg);
    std::vector<uint8_t> file(size);   // MemoryBitReader::loadDataFromFile (MemoryBitReader.cpp:93): the whole file in RAM
    is.read((char *)file.data(), (std::streamsize)size);
    if (e2fm_index_open(file.data(), size, key_, device_, &handle_) != E2FM_OK) {
        handle_ = nullptr;
        throw std::runtime_error(e2fm_last_error());   // e.g. "File type is not Encrypted and Compressed Genomic Index"
    }
    e2fm_info info;
    e2fm_index_info(handle_, &info);
    infoBucketSize_ = info.bucket_size;
    infoSuperBucketSize_ = info.superbucket_size;
    infoBuckets_ = info.n_buckets;
    infoSuperBuckets_ = info.n_superbuckets;
    infoN_ = info.text_length;
    infoK_ = info.k;
    infoMrp_ = info.marked_rows_percentage;
    terminators_.resize(info.n_sequences);
    e2fm_index_terminators(handle_, terminators_.data(), terminators_.size());
    uint64_t hs = e2fm_index_fasta_headers(handle_, nullptr, 0);
    headersBuffer_.resize(hs);
    if (hs) e2fm_index_fasta_headers(handle_, &headersBuffer_[0], hs);
}

void EFMIndex::clearMatches() {
    for (Match *m : allSuperPatternMatches) delete m;
    allSuperPatternMatches.clear();
}

void EFMIndex::close() {
    clearMatches();
    if (handle_) e2fm_index_close(handle_);
    handle_ = nullptr;
    for (Pinned *b : {&stageIn_, &stageOffs_, &stageCounts_, &stageCsr_, &stageSeq_, &stageOff_}) b->release();
}

void EFMIndex::requireLoaded() const {
    if (!handle_) throw std::runtime_error("Index not loaded");
}

BatchResult EFMIndex::searchBatch(const std::vector<std::string> &patterns, Mode mode) {
    requireLoaded();
    const double t0 = now_ms();
    BatchResult out;
    const uint64_t n = patterns.size();
    std::vector<uint64_t> offs(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) offs[i + 1] = offs[i] + patterns[i].size();
    std::vector<uint8_t> bytes(offs[n] + 1);
    for (uint64_t i = 0; i < n; i++) memcpy(bytes.data() + offs[i], patterns[i].data(), patterns[i].size());
    e2fm_result *r = nullptr;
    if (e2fm_search_batch(handle_, bytes.data(), offs.data(), n, (int)mode, &r) != E2FM_OK) throw std::runtime_error(e2fm_last_error());
    const uint64_t total = e2fm_result_total_positions(r);
    out.counts.resize(n);
    out.offsets.assign(n + 1, 0);
    out.positions.resize(total);
    out.sequenceIndex.resize(total);
    out.sequenceOffset.resize(total);
    int rc = e2fm_result_fetch(r, out.counts.data(), out.offsets.data(), out.positions.data(), out.sequenceIndex.data(),
                               out.sequenceOffset.data());
    e2fm_result_free(r);
    if (rc != E2FM_OK) raise_last("e2fm_result_fetch");
    e2fm_last_batch_ms(handle_, out.deviceMs);
    out.elapsedMs = now_ms() - t0;
    return out;
}

CompactBatchResult EFMIndex::searchBatchCompact(const std::vector<std::string> &patterns, Mode mode) {
    requireLoaded();
    const double t0 = now_ms();
    CompactBatchResult out;
    const uint64_t n = patterns.size();
    out.n = n;
    e2fm_result *r = nullptr;
    const uint32_t len = n ? (uint32_t)patterns[0].size() : 0;
    bool packed = n > 0 && len > 0 && len <= 128;
    if (packed) {
        // one pass over the strings on all host threads: 2 bits per base into page-locked memory; a different length or a byte
        // outside A C G T sends the batch through the ASCII entry point instead. The strings live all over the heap: the data of
        // the ones a few places ahead is prefetched while the current one is packed.
        static const struct Lut { uint8_t v[256]; Lut() { memset(v, 4, sizeof(v)); v['A'] = 0; v['C'] = 1; v['G'] = 2; v['T'] = 3; } } lut;
        const uint32_t stride = e2fm_packed_stride(len);
        uint8_t *dst = (uint8_t *)stageIn_.ensure(n * stride);
        std::atomic<bool> ok(true);
        parallel_for(n, [&](uint64_t a, uint64_t b) {
            uint32_t bad = 0;
            for (uint64_t i = a; i < b; i++) {
                if (i + 8 < b) __builtin_prefetch(patterns[i + 8].data());
                const std::string &p = patterns[i];
                if (p.size() != len) { bad = 4; break; }
                const uint8_t *src = reinterpret_cast<const uint8_t *>(p.data());
                uint64_t *w = reinterpret_cast<uint64_t *>(dst + i * stride);
                for (uint32_t q = 0; q < stride / 8; q++) {
                    uint64_t cur = 0;
                    const uint32_t end = std::min<uint32_t>(len, (q + 1) * 32);
                    for (uint32_t j = q * 32, sh = 0; j < end; j++, sh += 2) {
                        const uint32_t c = lut.v[src[j]];
                        bad |= c;
                        cur |= (uint64_t)(c & 3u) << sh;
                    }
                    w[q] = cur;
                }
            }
            if (bad & 4u) ok.store(false);
        });
        packed = ok.load();
        if (packed) {
            out.packMs = now_ms() - t0;
            // host to host in one pipelined call (H2D, search and D2H overlap stage by stage)
            e2fm_hits h;
            memset(&h, 0, sizeof(h));
            uint64_t cap = mode == Mode::Locate ? std::max<uint64_t>(lastTotal_ + lastTotal_ / 8, n + n / 4) + 4096 : 0;
            for (int attempt = 0; attempt < 2; attempt++) {
                if (mode == Mode::Count) h.counts = (uint32_t *)stageCounts_.ensure((n + 1) * 4);
                else {
                    h.occ_offsets = (uint32_t *)stageCsr_.ensure((n + 1) * 4);
                    h.seq_index = (uint32_t *)stageSeq_.ensure((cap + 1) * 4);
                    h.seq_offset = (uint32_t *)stageOff_.ensure((cap + 1) * 4);
                    h.capacity = cap;
                }
                const int rc = e2fm_search_hits(handle_, dst, n, len, 1, (int)mode, &h);
                if (rc == E2FM_OK) {
                    out.packed = true;
                    out.total = h.total;
                    lastTotal_ = h.total;
                    out.counts = h.counts;
                    out.offsets = h.occ_offsets;
                    out.sequenceIndex = h.seq_index;
                    out.sequenceOffset = h.seq_offset;
                    e2fm_last_batch_ms(handle_, out.deviceMs);
                    out.elapsedMs = now_ms() - t0;
                    return out;
                }
                if (rc == E2FM_ERR_ARG && mode == Mode::Locate && h.total > cap) { cap = h.total + 1024; continue; }   // more occurrences than room
                break;            // this batch shape is not for the seed search (short patterns, no seed table): ASCII below
            }
            packed = false;
        }
    }
    if (!packed) {
        uint64_t *offs = (uint64_t *)stageOffs_.ensure((n + 1) * 8);
        offs[0] = 0;
        for (uint64_t i = 0; i < n; i++) offs[i + 1] = offs[i] + patterns[i].size();
        uint8_t *bytes = (uint8_t *)stageIn_.ensure(offs[n] + 1);
        parallel_for(n, [&](uint64_t a, uint64_t b) {
            for (uint64_t i = a; i < b; i++) memcpy(bytes + offs[i], patterns[i].data(), patterns[i].size());
        });
        out.packMs = now_ms() - t0;
        if (e2fm_search_batch(handle_, bytes, offs, n, (int)mode, &r) != E2FM_OK) throw std::runtime_error(e2fm_last_error());
    }
    out.packed = packed;
    out.total = e2fm_result_total_positions(r);
    int rc;
    if (mode == Mode::Count) {
        uint32_t *c = (uint32_t *)stageCounts_.ensure((n + 1) * 4);
        rc = e2fm_result_fetch_compact(r, c, nullptr, nullptr, nullptr);
        out.counts = c;
    } else {
        uint32_t *o = (uint32_t *)stageCsr_.ensure((n + 1) * 4);
        uint32_t *sq = (uint32_t *)stageSeq_.ensure((out.total + 1) * 4);
        uint32_t *so = (uint32_t *)stageOff_.ensure((out.total + 1) * 4);
        rc = e2fm_result_fetch_compact(r, nullptr, o, sq, so);
        out.offsets = o;
        out.sequenceIndex = sq;
        out.sequenceOffset = so;
    }
    e2fm_result_free(r);
    if (rc != E2FM_OK) raise_last("e2fm_result_fetch_compact");
    e2fm_last_batch_ms(handle_, out.deviceMs);
    out.elapsedMs = now_ms() - t0;
    return out;
}

// EFMIndex::exactMatch (EFMIndex.cpp:1694): one pattern; the outcome is parked in allSuperPatternMatches
void EFMIndex::exactMatch(std::string pattern) {
    clearMatches();
    BatchResult b = searchBatch(std::vector<std::string>(1, pattern), Mode::Locate);
    Match *m = new Match();
    m->count = b.counts[0];
    m->occurrences = std::move(b.positions);
    m->sequenceIndex = std::move(b.sequenceIndex);
    m->sequenceOffset = std::move(b.sequenceOffset);
    allSuperPatternMatches.push_back(m);
}

uint64_t EFMIndex::countOccurrences(std::vector<Match *> *matches) {
    uint64_t c = 0;
    for (Match *m : *matches) c += m->count;
    return c;
}

std::vector<uint64_t> *EFMIndex::locateOccurrences(std::vector<Match *> *matches) {
    std::vector<uint64_t> *occs = new std::vector<uint64_t>();
    for (Match *m : *matches) occs->insert(occs->end(), m->occurrences.begin(), m->occurrences.end());
    return occs;
}

std::string EFMIndex::extractSnippet(uint64_t from, uint64_t to) {
    requireLoaded();
    std::string out;
    out.resize(to >= from ? (size_t)(to - from + 1) + 64 : 64);
    uint64_t len = 0;
    if (e2fm_extract(handle_, from, to, &out[0], out.size(), &len) != E2FM_OK) throw std::runtime_error(e2fm_last_error());
    out.resize((size_t)len);
    return out;
}

uint64_t EFMIndex::getBucketSize() const { return infoBucketSize_; }
uint64_t EFMIndex::getSuperBucketSize() { return infoSuperBucketSize_; }
uint64_t EFMIndex::getBucketsNumber() { return infoBuckets_; }
uint64_t EFMIndex::getSuperBucketsNumber() { return infoSuperBuckets_; }
uint64_t EFMIndex::getNumberOfLoadedBuckets() { return handle_ ? infoBuckets_ : 0; }
uint64_t EFMIndex::getNumberOfLoadedSuperBuckets() { return handle_ ? infoSuperBuckets_ : 0; }
uint8_t EFMIndex::getMarkedRowsPercentage() { return (uint8_t)infoMrp_; }
uint8_t EFMIndex::getSuperAlphabetOrder() const { return (uint8_t)infoK_; }
uint64_t EFMIndex::getTextLength() const { return infoN_; }

// ------------------------------------------------------------------ EFMCollection

uint32_t EFMCollection::getSize() { return (uint32_t)terminators_.size(); }

std::string EFMCollection::getDescription(uint32_t sequenceIndex) {
    if (!headersSplit_) {   // boost::split(headersBuffer, "\n") (FMCollection.cpp:72)
        fastaHeaders_.clear();
        size_t a = 0;
        while (true) {
            size_t b = headersBuffer_.find('\n', a);
            fastaHeaders_.push_back(headersBuffer_.substr(a, b == std::string::npos ? std::string::npos : b - a));
            if (b == std::string::npos) break;
            a = b + 1;
        }
        headersSplit_ = true;
    }
    return fastaHeaders_.at(sequenceIndex);   // throws std::out_of_range like vector::at in the reference
}

std::vector<SearchResult> EFMCollection::searchBatch(const std::vector<std::string> &patterns, Mode mode, bool retrieveSequenceDescriptions) {
    const double t0 = now_ms();
    const CompactBatchResult b = searchBatchCompact(patterns, mode);
    std::vector<SearchResult> out(patterns.size());
    if (retrieveSequenceDescriptions && getSize()) getDescription(0);      // split the headers once, before the threads read them
    // 10^7 SearchResult objects are 2 x 10^7 heap allocations: built by all host threads, every vector reserved to its final size
    parallel_for(patterns.size(), [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; i++) {
            SearchResult &sr = out[i];
            sr.pattern = patterns[i];
            if (mode != Mode::Locate) continue;
            const uint32_t a = b.offsets[i], e = b.offsets[i + 1];
            sr.occurrences.reserve(e - a);
            for (uint32_t j = a; j < e; j++) {
                const uint32_t s = b.sequenceIndex[j];
                if (retrieveSequenceDescriptions && s != 0xFFFFFFFFu) sr.occurrences.emplace_back(s, fastaHeaders_.at(s), (uint64_t)b.sequenceOffset[j]);
                else sr.occurrences.emplace_back(s, std::string(), (uint64_t)b.sequenceOffset[j]);
            }
        }
    });
    const double per = patterns.empty() ? 0 : (now_ms() - t0) / (double)patterns.size();
    for (SearchResult &sr : out) sr.elapsedTime = per;
    return out;
}

// EFMCollection::getSequence (EFMCollection.cpp:43-56)
std::string EFMCollection::getSequence(uint32_t sequenceIndex) {
    uint64_t from = sequenceIndex > 0 ? (uint64_t)terminators_.at(sequenceIndex - 1) + getSuperAlphabetOrder() : 0;
    uint64_t to = (uint64_t)terminators_.at(sequenceIndex);
    // the reference keeps find('&') in a uint64_t and tests `p > -1` (EFMCollection.cpp:51-55): never true, so the padding and
    // the first terminator symbol stay in the returned string; mirrored
    return extractSnippet(from, to);
}

// EFMCollection::extract (EFMCollection.cpp:67-95): same range checks, same messages
std::string EFMCollection::extract(uint32_t sequenceIndex, uint64_t from, uint64_t to) {
    int64_t wholeFrom = sequenceIndex > 0 ? terminators_.at(sequenceIndex - 1) + (int64_t)getSuperAlphabetOrder() : 0;
    int64_t wholeTo = terminators_.at(sequenceIndex);
    int64_t maxLen = wholeTo - wholeFrom + 1;   // can exceed the real length by at most k-1 characters
    if (from > (uint64_t)(maxLen - 1)) throw std::runtime_error("start index out of range");
    if (to > (uint64_t)(maxLen - 1)) throw std::runtime_error("end index out of range");
    std::string snippet = extractSnippet((uint64_t)(wholeFrom + (int64_t)from), (uint64_t)(wholeFrom + (int64_t)to));
    size_t p = snippet.find('&');
    return p != std::string::npos ? snippet.substr(0, p) : snippet;
}

SearchResult *EFMCollection::search(std::string pattern, bool retrieveSequenceDescriptions) {
    std::vector<SearchResult> v = searchBatch(std::vector<std::string>(1, pattern), Mode::Locate, retrieveSequenceDescriptions);
    return new SearchResult(std::move(v[0]));
}

}  // namespace e2fm
