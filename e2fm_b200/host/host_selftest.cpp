// host_selftest — drives the C++ host classes the way the reference's own drivers do (Main.cpp:1022 locate,
// Test.cpp:113 executePatternSearchInCollection) and prints what they return, for tests/test_gpu_parity.py:
//   host_selftest <index> <keyfile> <patterns>
//   host_selftest <index> <keyfile> <patterns> [extract-queries]
// per pattern one line:  P <i> <count> <n> <pos...> | <seq:off ...> | <sfm seq:off ...>
// per extract query ("S seq from to" | "G seq" | "X from to", as e2fm_ref extract takes them) one line:  E <result or !what()>
#include <fstream>
#include <iostream>
#include <sstream>

#include "e2fm_host.h"

using namespace e2fm;

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    try {
        EFMCollection efm;
        efm.setEncryptionKeyFilePath(argv[2]);
        efm.setNumberOfThreads(1);
        efm.load(argv[1]);
        SFMCollection sfm;
        sfm.setEncryptionKeyFilePath(argv[2]);
        sfm.load(argv[1]);
        std::cout << "INFO " << efm.getSize() << " " << (int)efm.getSuperAlphabetOrder() << " " << efm.getBucketSize() << " " << efm.getSuperBucketSize()
                  << " " << efm.getBucketsNumber() << " " << efm.getSuperBucketsNumber() << " " << (int)efm.getMarkedRowsPercentage() << " "
                  << efm.getNumberOfLoadedBuckets() << "\n";
        for (uint32_t i = 0; i < efm.getSize(); i++) std::cout << "DESC " << i << " " << efm.getDescription(i) << "\n";
        std::ifstream pf(argv[3]);
        std::string p;
        int i = 0;
        while (std::getline(pf, p)) {
            efm.exactMatch(p);                                                       // EFMIndex.h:279
            uint64_t c = efm.countOccurrences(&efm.allSuperPatternMatches);          // EFMIndex.h:283
            std::vector<uint64_t> *occ = efm.locateOccurrences(&efm.allSuperPatternMatches);   // EFMIndex.h:282, caller owns
            std::cout << "P " << i << " " << c << " " << occ->size();
            for (uint64_t o : *occ) std::cout << " " << o;
            delete occ;
            SearchResult *r = efm.search(p, true);                                   // EFMCollection.cpp:115, caller owns
            std::cout << " |";
            for (const Occurrence &o : r->occurrences) std::cout << " " << o.sequenceIndex << ":" << o.patternPosition << ":" << o.sequenceDescription;
            delete r;
            r = sfm.search(p, false);                                                // SFMCollection.cpp:59
            std::cout << " |";
            for (const Occurrence &o : r->occurrences) std::cout << " " << o.sequenceIndex << ":" << o.patternPosition;
            delete r;
            std::cout << "\n";
            i++;
        }
        if (argc > 4) {
            std::ifstream qf(argv[4]);
            std::string line;
            while (std::getline(qf, line)) {
                if (line.empty()) continue;
                std::istringstream is(line);
                char kind;
                is >> kind;
                std::string r;
                try {
                    if (kind == 'S') { uint32_t q; uint64_t a, b; is >> q >> a >> b; r = efm.extract(q, a, b); }        // EFMCollection.cpp:67
                    else if (kind == 'G') { uint32_t q; is >> q; r = efm.getSequence(q); }                               // EFMCollection.cpp:43
                    else { uint64_t a, b; is >> a >> b; r = efm.extractSnippet(a, b); }                                 // EFMIndex.cpp:2325
                } catch (std::exception &e) { r = std::string("!") + e.what(); }
                std::cout << "E " << r << "\n";
            }
        }
        // error behaviour: bad key file, bad magic
        try { EFMCollection bad; bad.setEncryptionKeyFilePath(argv[1]); std::cout << "ERR none\n"; }
        catch (std::exception &e) { std::cout << "ERR " << e.what() << "\n"; }
        try { EFMCollection bad; bad.setEncryptionKeyFilePath(argv[2]); bad.load(argv[2]); std::cout << "ERR none\n"; }
        catch (std::exception &e) { std::cout << "ERR " << e.what() << "\n"; }
        sfm.close();
        efm.close();
    } catch (std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
