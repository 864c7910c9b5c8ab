// ref_cli — driver that links against the UNMODIFIED reference sources (compiled from where they
// lie under /root/reference by oracle/Makefile) and exposes what the tests and the CPU baseline need:
//
//   e2fm_ref build  <fasta> <index> <keyfile> <k> <bucketSize> <mrp> [threads]
//   e2fm_ref search <index> <keyfile> <patterns.txt> <out.bin> <mode:0=count,1=count+locate> [threads] [passes]
//   e2fm_ref dump   <index> <keyfile> <out.bin>          (decoded internals after readText())
//   e2fm_ref extract <index> <keyfile> <queries.txt> <out.txt>   (EFMCollection::extract / getSequence, EFMIndex::extractSnippet)
//   e2fm_ref kat                                          (known-answer vectors, text)
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product path (e2fm_b200/, include/) links or runs this.
// This file is ours; it contains no reference code. The reference classes it drives:
//   EFMCollection::build/load/search   EFMCollection.cpp:115,327,333
//   EFMIndex::exactMatch/countOccurrences/locateOccurrences   EFMIndex.cpp:1694,2273,2281
//   EFMIndex::readText   EFMIndex.cpp:1348
#define private public
#define protected public
#include "EFMCollection.h"
#include "EncryptionManager.h"
#include "RandomGenerator.h"
#include "BitSetRLEEncoder.h"
#include "Utils.h"
#undef private
#undef protected

#include <cstdio>
#include <cstdlib>
#include <unistd.h>

using namespace std;

static void die(const string &m) { fprintf(stderr, "e2fm_ref: %s\n", m.c_str()); _Exit(2); }

static void w64(FILE *f, uint64_t v) { fwrite(&v, 8, 1, f); }
static void w32(FILE *f, uint32_t v) { fwrite(&v, 4, 1, f); }

static vector<string> readLines(const string &path) {
    ifstream in(path);
    if (!in) die("cannot open " + path);
    vector<string> v; string l;
    while (getline(in, l)) {
        if (!l.empty() && l[l.size() - 1] == '\r') l.erase(l.size() - 1);
        for (size_t i = 0; i < l.size(); i++) l[i] = toupper(l[i]);   // Main.cpp:1056-1059
        v.push_back(l);
    }
    return v;
}

static int cmdBuild(int argc, char **argv) {
    if (argc < 8) die("build <fasta> <index> <keyfile> <k> <bucketSize> <mrp> [threads]");
    string fasta = argv[2], index = argv[3], key = argv[4];
    int k = atoi(argv[5]); long bs = atol(argv[6]); int mrp = atoi(argv[7]);
    // same sequence of setters as the reference CLI's build(), Main.cpp:735-752
    EFMCollection *efm = new EFMCollection();
    efm->setSuperAlphabetOrder(k);
    efm->setBucketSize(bs);
    efm->setSuperBucketSize(bs * 16);
    efm->setMarkedRowsPercentage((uint8_t)mrp);
    efm->setEncryptionKeyFilePath(key);
    efm->setComputeStatistics(mrp > 0);   // writeHeader dereferences an unset markedRows pointer when mrp==0 (EFMIndex.cpp:1048)
    efm->maximumBucketAlphaSize = 0;   // uninitialised in the reference (EFMIndex.h:395); pin it so files are reproducible
    if (argc > 8) efm->setNumberOfThreads(atoi(argv[8]));
    auto t0 = chrono::steady_clock::now();
    efm->build(fasta, index);
    double ms = chrono::duration<double, milli>(chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "build_ms=%.1f n=%llu buckets=%llu superbuckets=%llu compactAlphabet=%u\n", ms,
            (unsigned long long)efm->textLength, (unsigned long long)efm->bucketsNumber,
            (unsigned long long)efm->superBucketsNumber, efm->compactAlphabetSize);
    fflush(stdout); fflush(stderr);
    _Exit(0);
}

// out.bin layout (little endian):
//   u64 magic 'E2FMRES1'; u64 nPatterns; u64 mode; u64 totalOcc
//   u64 counts[n]; u64 nocc[n]; u64 gpos[totalOcc]; u32 seq[totalOcc]; u64 off[totalOcc]
// stdout: one JSON line with timings of every pass (wall clock around the pattern loop only, Main.cpp:1062-1064)
static int cmdSearch(int argc, char **argv) {
    if (argc < 7) die("search <index> <keyfile> <patterns.txt> <out.bin> <mode> [threads] [passes]");
    string index = argv[2], key = argv[3], pats = argv[4], out = argv[5];
    int mode = atoi(argv[6]);
    int threads = argc > 7 ? atoi(argv[7]) : 1;
    int passes = argc > 8 ? atoi(argv[8]) : 1;
    vector<string> patterns = readLines(pats);
    EFMCollection *efm = new EFMCollection();
    efm->setEncryptionKeyFilePath(key);
    efm->setNumberOfThreads(threads);
    auto l0 = chrono::steady_clock::now();
    efm->load(index);
    double loadMs = chrono::duration<double, milli>(chrono::steady_clock::now() - l0).count();
    size_t n = patterns.size();
    vector<uint64_t> counts(n), nocc(n);
    vector<uint64_t> gpos, off; vector<uint32_t> seq;
    vector<double> passMs;
    for (int pass = 0; pass < passes; pass++) {
        gpos.clear(); off.clear(); seq.clear();
        auto t0 = chrono::steady_clock::now();
        for (size_t i = 0; i < n; i++) {
            if (mode == 0) {
                efm->exactMatch(patterns[i]);
                counts[i] = efm->countOccurrences(&efm->allSuperPatternMatches);
                nocc[i] = 0;
            } else {
                // EFMCollection::search does exactMatch + locateOccurrences + (seq,offset) mapping
                SearchResult *r = efm->search(patterns[i], false);
                counts[i] = efm->countOccurrences(&efm->allSuperPatternMatches);
                nocc[i] = r->occurrences.size();
                for (size_t j = 0; j < r->occurrences.size(); j++) {
                    const Occurrence &o = r->occurrences[j];
                    uint64_t start = o.sequenceIndex > 0 && o.sequenceIndex != 0xFFFFFFFFu
                                         ? (uint64_t)efm->terminatorsPositions->at(o.sequenceIndex - 1) + efm->superAlphabetOrder
                                         : 0;
                    gpos.push_back(o.patternPosition + start);
                    seq.push_back(o.sequenceIndex);
                    off.push_back(o.patternPosition);
                }
                delete r;
            }
        }
        passMs.push_back(chrono::duration<double, milli>(chrono::steady_clock::now() - t0).count());
    }
    FILE *f = fopen(out.c_str(), "wb");
    if (!f) die("cannot write " + out);
    w64(f, 0x3153455246324545ull /* "EE2FRES1" */); w64(f, n); w64(f, mode); w64(f, gpos.size());
    fwrite(counts.data(), 8, n, f); fwrite(nocc.data(), 8, n, f);
    fwrite(gpos.data(), 8, gpos.size(), f); fwrite(seq.data(), 4, seq.size(), f); fwrite(off.data(), 8, off.size(), f);
    fclose(f);
    printf("{\"patterns\": %zu, \"mode\": %d, \"threads\": %d, \"load_ms\": %.3f, \"pass_ms\": [", n, mode, threads, loadMs);
    for (size_t i = 0; i < passMs.size(); i++) printf("%s%.3f", i ? ", " : "", passMs[i]);
    printf("], \"occurrences\": %zu, \"loaded_buckets\": %llu}\n", gpos.size(),
           (unsigned long long)efm->getNumberOfLoadedBuckets());
    fflush(stdout);
    _Exit(0);   // EFMIndex::close() frees uninitialised pointers with 1 matcher thread (EFMIndex.cpp:1661-1664)
}

// dump layout (little endian u64 unless noted):
//   magic 'E2FMDMP1', n, k, nSym, sym bytes (nSym x u64), initialN, finalN, lastCharPos, SBsize, Bsize, mrp,
//   A, C[A+1], superAlphabetSize, charactersRemappings-or-(2^64-1)[size], permutation[size],
//   mrn, markedRowsBuckets[mrn], S, B,
//   per SB: alphaSize, bits[A], occ[A]
//   per bucket: length, alpha, isOdd, isFirst, clen, hasTable, table[sbAlpha] (if hasTable), bits[sbAlpha],
//               marked[length], compact bwt code[length] (logical order), nMarked, (offset,value) pairs
//   nTerm, terminators[nTerm]
static int cmdDump(int argc, char **argv) {
    if (argc < 5) die("dump <index> <keyfile> <out.bin>");
    EFMCollection *efm = new EFMCollection();
    efm->setEncryptionKeyFilePath(argv[3]);
    efm->setNumberOfThreads(1);
    efm->load(argv[2]);
    // same loop as EFMIndex::readText (EFMIndex.cpp:1348-1356), minus readMarkedRows when mrp==0
    // (the reference dereferences a NULL marked bitmap there)
    bool marks = efm->markedRowsPercentage > 0;
    for (uint64_t b = 0; b < efm->bucketsNumber; b++) {
        efm->checkBucketLoaded(efm->bitReader, b);
        efm->buckets[b]->readText(efm->bitReader, efm->bucketEncryptionRandomGenerator);
        if (marks) efm->buckets[b]->readMarkedRows(efm->bitReader);
    }
    FILE *f = fopen(argv[4], "wb");
    if (!f) die("cannot write dump");
    uint32_t A = efm->compactAlphabetSize;
    w64(f, 0x31504d444d463245ull); w64(f, efm->textLength); w64(f, efm->superAlphabetOrder);
    w64(f, efm->originalSymbols.size());
    for (char c : efm->originalSymbols) w64(f, (uint8_t)c);
    w64(f, efm->initialNCharacters); w64(f, efm->finalNCharacters); w64(f, efm->lastCharacterPosition);
    w64(f, efm->superBucketSize); w64(f, efm->bucketSize); w64(f, efm->markedRowsPercentage);
    w64(f, A);
    for (uint32_t i = 0; i <= A; i++) w64(f, efm->precedingCharactersOccurrences[i]);
    uint32_t sz = efm->superAlphabet->getSize();
    w64(f, sz);
    for (uint32_t i = 0; i < sz; i++) w64(f, (*efm->charactersBitSet)[i] ? (uint64_t)efm->charactersRemappings[i] : ~0ull);
    for (uint32_t i = 0; i < sz; i++) w64(f, (*efm->superAlphabet->permutation)[i]);
    uint64_t mrn = marks ? efm->markedRowsNumber : 0;
    w64(f, mrn);
    for (uint64_t i = 0; i < mrn; i++) w64(f, efm->markedRowsBuckets[i]);
    w64(f, efm->superBucketsNumber); w64(f, efm->bucketsNumber);
    for (uint64_t s = 0; s < efm->superBucketsNumber; s++) {
        SuperBucket *sb = efm->superBuckets[s];
        w64(f, sb->alphaSize);
        for (uint32_t i = 0; i < A; i++) w64(f, (*sb->charactersBitSet)[i] ? 1 : 0);
        for (uint32_t i = 0; i < A; i++) w64(f, sb->previousSuperbucketsOccurrences[i]);
    }
    for (uint64_t b = 0; b < efm->bucketsNumber; b++) {
        Bucket *bk = efm->buckets[b];
        SuperBucket *sb = bk->superBucket;
        w64(f, bk->length); w64(f, bk->bucketAlphaSize); w64(f, bk->isOdd); w64(f, bk->isFirstInSuperBucket);
        w64(f, bk->bucketAlphaSize > 1 ? bk->compressedLength : 0);
        bool hasTable = bk->previousBucketsOccurrences != NULL;
        w64(f, hasTable);
        if (hasTable) for (uint32_t i = 0; i < bk->superBucketAlphaSize; i++) w64(f, bk->previousBucketsOccurrences[i]);
        for (uint32_t i = 0; i < bk->superBucketAlphaSize; i++) w64(f, (*bk->charactersBitSet)[i] ? 1 : 0);
        for (uint64_t i = 0; i < bk->length; i++) w64(f, marks && (*bk->markedRowsBitmap)[i] ? 1 : 0);
        for (uint64_t i = 0; i < bk->length; i++) w64(f, sb->decode_map[bk->decodeMap[bk->bucketText[i]]]);
        w64(f, bk->markedRowsCache.size());
        for (auto &kv : bk->markedRowsCache) { w64(f, kv.first); w64(f, kv.second); }
    }
    w64(f, efm->terminatorsPositions->size());
    for (int64_t t : *efm->terminatorsPositions) w64(f, t);
    fclose(f);
    fflush(stdout);
    _Exit(0);
}

// extract <index> <keyfile> <queries.txt> <out.txt>: one query per line, "S <seq> <from> <to>" = EFMCollection::extract,
// "G <seq>" = EFMCollection::getSequence, "X <from> <to>" = EFMIndex::extractSnippet; one result line each ("!" + what() on throw)
static int cmdExtract(int argc, char **argv) {
    if (argc < 6) die("extract <index> <keyfile> <queries.txt> <out.txt>");
    EFMCollection *efm = new EFMCollection();
    efm->setEncryptionKeyFilePath(argv[3]);
    efm->setNumberOfThreads(2);
    efm->load(argv[2]);
    ifstream in(argv[4]);
    FILE *f = fopen(argv[5], "w");
    if (!in || !f) die("cannot open query / output file");
    string line;
    while (getline(in, line)) {
        if (line.empty()) continue;
        istringstream is(line);
        char kind; is >> kind;
        string r;
        try {
            if (kind == 'S') { uint32_t q; uint64_t a, b; is >> q >> a >> b; r = efm->extract(q, a, b); }
            else if (kind == 'G') { uint32_t q; is >> q; r = efm->getSequence(q); }
            else { uint64_t a, b; is >> a >> b; r = efm->extractSnippet(a, b); }
        } catch (std::exception &e) { r = string("!") + e.what(); }
        fprintf(f, "%s\n", r.c_str());
    }
    fclose(f);
    _Exit(0);
}

static int cmdKat(int, char **) {
    // Salsa20/20 ECRYPT vector: key = 80 00.., iv = 0 (drives the reference's salsa20.S through its C ABI)
    u8 key[32]; memset(key, 0, 32); key[0] = 0x80;
    u8 iv[8]; memset(iv, 0, 8);
    ECRYPT_ctx ctx; ECRYPT_init(); ECRYPT_keysetup(&ctx, key, 256, 64); ECRYPT_ivsetup(&ctx, iv);
    u8 ks[128]; ECRYPT_keystream_bytes(&ctx, ks, 128);
    printf("salsa20_block0="); for (int i = 0; i < 64; i++) printf("%02X", ks[i]); printf("\n");
    printf("salsa20_block1="); for (int i = 64; i < 128; i++) printf("%02X", ks[i]); printf("\n");
    u8 k64[64]; for (int i = 0; i < 64; i++) k64[i] = (7 * i + 3) & 0xFF;
    EncryptionManager em(k64);
    {
        RandomGenerator rg(131072, em.polySeed);
        rg.populateKeyStreamBuffer(5, 256, 0, 4095);
        printf("rg_nonce5_max256="); for (int i = 0; i < 16; i++) printf("%u ", rg.nextInt()); printf("\n");
        rg.populateKeyStreamBuffer(5, 256, 100, 103);
        printf("rg_nonce5_max256_start100="); for (int i = 0; i < 4; i++) printf("%u ", rg.nextInt()); printf("\n");
        rg.populateKeyStreamBuffer(77, 255, 0, 4095);
        printf("rg_nonce77_max255="); for (int i = 0; i < 16; i++) printf("%u ", rg.nextInt()); printf("\n");
        rg.populateKeyStreamBuffer(123456, 1, 0, 999);
        printf("rg_nonce123456_max1="); for (int i = 0; i < 32; i++) printf("%u", rg.nextInt()); printf("\n");
    }
    {
        vector<uint32_t> *sk = em.computeScramblingKey(1296);
        printf("scramble1296="); for (int i = 0; i < 1296; i++) printf("%u ", (*sk)[i]); printf("\n");
        vector<uint32_t> *sk2 = em.computeScramblingKey(625);
        printf("scramble625="); for (int i = 0; i < 625; i++) printf("%u ", (*sk2)[i]); printf("\n");
        vector<uint32_t> *sk3 = em.computeScramblingKey(7776);   // exercises the 16384-byte refill-restart (quirk 6)
        printf("scramble7776="); for (int i = 0; i < 7776; i++) printf("%u ", (*sk3)[i]); printf("\n");
    }
    {
        uint32_t *p = em.computeFrequenciesPermutation(259, (1ull << 22) - 1);
        printf("freqperm259="); for (int i = 0; i < 259; i++) printf("%u ", p[i]); printf("\n");
        uint64_t *ks2 = em.computeCumulativeFreqsKeyStream(259, (1ull << 22) - 1);
        printf("freqks259_22="); for (int i = 0; i < 259; i++) printf("%llu ", (unsigned long long)ks2[i]); printf("\n");
        uint64_t *ks3 = em.computeCumulativeFreqsKeyStream(259, (1ull << 30) - 1);   // 32-bit bit-buffer overflow (quirk 5)
        printf("freqks259_30="); for (int i = 0; i < 259; i++) printf("%llu ", (unsigned long long)ks3[i]); printf("\n");
    }
    printf("int_log2=");
    uint64_t xs[] = {0, 1, 2, 3, 4, 7, 8, 255, 256, 257, 4095, 4096, 65535, 65536, (1ull << 32) - 1, 1ull << 32};
    for (uint64_t x : xs) printf("%llu:%u ", (unsigned long long)x, (unsigned)Utils::int_log2(x));
    printf("\n");
    {
        boost::dynamic_bitset<> bs(40);
        int on[] = {0, 1, 9, 10, 30};
        for (int i : on) bs.set(i);
        boost::dynamic_bitset<> *c = BitSetRLEEncoder::compress(&bs);
        printf("rle_compress="); for (size_t i = 0; i < c->size(); i++) printf("%d", (int)(*c)[i]); printf("\n");
        boost::dynamic_bitset<> *d = BitSetRLEEncoder::decompress(c);
        printf("rle_decompress="); for (size_t i = 0; i < d->size(); i++) printf("%d", (int)(*d)[i]); printf("\n");
    }
    fflush(stdout);
    _Exit(0);
}

int main(int argc, char **argv) {
    if (argc < 2) die("usage: e2fm_ref build|search|dump|kat ...");
    string cmd = argv[1];
    try {
        if (cmd == "build") return cmdBuild(argc, argv);
        if (cmd == "search") return cmdSearch(argc, argv);
        if (cmd == "dump") return cmdDump(argc, argv);
        if (cmd == "extract") return cmdExtract(argc, argv);
        if (cmd == "kat") return cmdKat(argc, argv);
    } catch (const exception &e) {
        die(string("exception: ") + e.what());
    }
    die("unknown command " + cmd);
    return 2;
}
