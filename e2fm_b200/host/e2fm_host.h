// e2fm_host.h — C++ host classes over the C ABI (include/e2fm_b200.h): the reference's load-and-search surface
// (EFMIndex.h:241-287, FMCollection.h:23-95, EFMCollection.h:22-36, SFMCollection.h:27-31) with the same method
// names, argument meaning, ownership (returned raw pointers belong to the caller) and error behaviour
// (std::runtime_error with the reference's messages), plus the batched searchBatch the GPU path is built for.
// The reference puts its classes in namespace std; these live in namespace e2fm.
//
// Everything that computes goes through libe2fm_b200.so (CUDA); there is no CPU search path here.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct e2fm_index;

namespace e2fm {

typedef uint8_t u8;

// FMCollection.h:23-44
struct Occurrence {
    Occurrence(uint32_t sequenceIndex, std::string sequenceDescription, uint64_t patternPosition)
        : sequenceIndex(sequenceIndex), sequenceDescription(std::move(sequenceDescription)), patternPosition(patternPosition) {}
    uint32_t sequenceIndex;
    std::string sequenceDescription;
    uint64_t patternPosition;
};

// FMCollection.h:47-61
struct SearchResult {
    SearchResult() : elapsedTime(0) {}
    std::string pattern;
    double elapsedTime;   // ms; for batched calls: batch time / number of patterns
    std::vector<Occurrence> occurrences;
};

// What exactMatch leaves behind for countOccurrences / locateOccurrences (reference: EFMIndex.h Match — a super-pattern
// with its suffix-array interval; here the device already resolved the intervals, so a Match carries their outcome).
struct Match {
    uint64_t count = 0;                  // sum of interval lengths (EFMIndex::countOccurrences, EFMIndex.cpp:2273)
    std::vector<uint64_t> occurrences;   // sorted, de-duplicated global text positions (EFMIndex.cpp:2281-2301)
    std::vector<uint32_t> sequenceIndex; // parallel to occurrences (EFMCollection.cpp:124-148)
    std::vector<uint64_t> sequenceOffset;
};

enum class Mode { Count = 0, Locate = 1 };

// result of one batch, CSR over the patterns
struct BatchResult {
    std::vector<uint64_t> counts;        // [n]
    std::vector<uint64_t> offsets;       // [n+1] (all zero in Count mode)
    std::vector<uint64_t> positions;     // global text positions
    std::vector<uint32_t> sequenceIndex;
    std::vector<uint64_t> sequenceOffset;
    double elapsedMs = 0;                // host wall clock of the whole call
    float deviceMs[8] = {0};             // e2fm_last_batch_ms
};

// The same CSR in 32 bits, in page-locked buffers OWNED BY THE INDEX OBJECT (valid until its next search or close): what a
// caller with millions of patterns reads instead of 10^7 SearchResult objects. counts is filled in Count mode, the other three
// in Locate mode (a pattern's occurrences are [offsets[i], offsets[i+1])).
struct CompactBatchResult {
    uint64_t n = 0, total = 0;
    const uint32_t *counts = nullptr;
    const uint32_t *offsets = nullptr;
    const uint32_t *sequenceIndex = nullptr;
    const uint32_t *sequenceOffset = nullptr;
    bool packed = false;                 // the batch went to the device as 2-bit codes (e2fm_search_batch_packed)
    double elapsedMs = 0, packMs = 0;    // whole call / its host-side marshalling part
    float deviceMs[8] = {0};
};

class EFMIndex {
public:
    EFMIndex();
    virtual ~EFMIndex();
    // --- the reference's setters that matter for load + search (EFMIndex.h:265-271)
    void setEncryptionKey(u8 *encryptionKey);                       // 64 bytes, copied
    void setEncryptionKeyFilePath(std::string encryptionKeyFilePath);
    std::string getEncryptionKeyFilePath();
    void setNumberOfThreads(uint32_t numberOfThreads);              // accepted for compatibility; the GPU path ignores it
    uint32_t getNumberOfThreads();
    void setDevice(int device);                                     // new: CUDA device ordinal (default 0)
    // --- load / close (EFMIndex.cpp:1536, :1602)
    void load(std::string fileName);
    void close();
    // --- search (EFMIndex.h:279-283): exactMatch fills allSuperPatternMatches
    void exactMatch(std::string pattern);
    std::vector<uint64_t> *locateOccurrences(std::vector<Match *> *matches);   // caller owns the result
    uint64_t countOccurrences(std::vector<Match *> *matches);
    std::string extractSnippet(uint64_t from, uint64_t to);         // EFMIndex.cpp:2325
    // --- getters (EFMIndex.h:245-262)
    uint64_t getBucketSize() const;
    uint64_t getSuperBucketSize();
    uint64_t getBucketsNumber();
    uint64_t getSuperBucketsNumber();
    uint64_t getNumberOfLoadedBuckets();        // every bucket is decoded at load
    uint64_t getNumberOfLoadedSuperBuckets();
    uint8_t getMarkedRowsPercentage();
    uint8_t getSuperAlphabetOrder() const;
    uint64_t getTextLength() const;             // super-text length n
    // --- batched entry points (new)
    BatchResult searchBatch(const std::vector<std::string> &patterns, Mode mode);
    // patterns of one length over A C G T are packed to 2 bits per base by all host threads straight into page-locked memory
    // and searched through e2fm_search_batch_packed; anything else goes through the ASCII entry point. Same answers.
    CompactBatchResult searchBatchCompact(const std::vector<std::string> &patterns, Mode mode);

    std::vector<Match *> allSuperPatternMatches;   // protected in the reference (EFMIndex.h:431); public here for drivers

protected:
    void requireLoaded() const;
    void clearMatches();
    ::e2fm_index *handle_;
    int device_;
    uint32_t numberOfThreads_;
    bool haveKey_;
    u8 key_[64];
    std::string keyFilePath_;
    std::vector<int64_t> terminators_;
    std::string headersBuffer_;
    uint64_t infoBucketSize_, infoSuperBucketSize_, infoBuckets_, infoSuperBuckets_, infoN_;
    uint32_t infoK_, infoMrp_;
    // grow-only page-locked staging (e2fm_host_alloc): patterns in, 32-bit results out
    struct Pinned {
        void *p = nullptr;
        size_t bytes = 0;
        void *ensure(size_t need);
        void release();
    } stageIn_, stageOffs_, stageCounts_, stageCsr_, stageSeq_, stageOff_;
    uint64_t lastTotal_ = 0;                    // occurrences of the previous batch: sizes the next one's result buffers
};

// FMCollection.h:65-95 (the part that load + search use)
class FMCollection {
public:
    virtual ~FMCollection() {}
    virtual uint32_t getSize() = 0;
    virtual std::string getDescription(uint32_t sequenceIndex) = 0;
    virtual SearchResult *search(std::string pattern, bool retrieveSequenceDescriptions) = 0;   // caller owns the result
    virtual void load(std::string fileName) = 0;
};

class EFMCollection : public EFMIndex, public FMCollection {
public:
    EFMCollection() {}
    virtual ~EFMCollection() {}
    virtual void load(std::string fileName) { EFMIndex::load(fileName); }
    virtual uint32_t getSize();
    virtual std::string getDescription(uint32_t sequenceIndex);   // FMCollection.cpp:68-74
    virtual SearchResult *search(std::string pattern, bool retrieveSequenceDescriptions);   // EFMCollection.cpp:115-153
    virtual std::string getSequence(uint32_t sequenceIndex);                                // EFMCollection.cpp:43-56
    virtual std::string extract(uint32_t sequenceIndex, uint64_t from, uint64_t to);        // EFMCollection.cpp:67-95
    // batched: one device call for all patterns (results owned by the caller's vector)
    std::vector<SearchResult> searchBatch(const std::vector<std::string> &patterns, Mode mode, bool retrieveSequenceDescriptions);
private:
    std::vector<std::string> fastaHeaders_;
    bool headersSplit_ = false;
};

// The reference's SFMCollection is the plain sdsl FM-index baseline (SFMCollection.h:40); its arithmetic lives in
// sdsl-lite, which is not vendored. This class keeps its load/search/close surface (SFMCollection.h:27-31) and serves
// it from the same E2FM index + kernels; occurrences are the same (sequence, offset) pairs.
class SFMCollection : public FMCollection {
public:
    SFMCollection() {}
    virtual ~SFMCollection() {}
    void setEncryptionKeyFilePath(std::string p) { impl_.setEncryptionKeyFilePath(p); }
    void setEncryptionKey(u8 *k) { impl_.setEncryptionKey(k); }
    void setDevice(int d) { impl_.setDevice(d); }
    virtual void load(std::string fileName) { impl_.load(fileName); }
    virtual SearchResult *search(std::string pattern, bool retrieveSequenceDescriptions) { return impl_.search(pattern, retrieveSequenceDescriptions); }
    virtual uint32_t getSize() { return impl_.getSize(); }
    virtual std::string getDescription(uint32_t i) { return impl_.getDescription(i); }
    std::vector<SearchResult> searchBatch(const std::vector<std::string> &patterns, Mode mode, bool descr) {
        return impl_.searchBatch(patterns, mode, descr);
    }
    void close() { impl_.close(); }
private:
    EFMCollection impl_;
};

}  // namespace e2fm
