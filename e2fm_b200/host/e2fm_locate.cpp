// e2fm_locate — the reference CLI's `locate` command (Main.cpp:1022-1151, dispatched at :1580-1608) on the GPU path:
//   e2fm_locate <index.efm> <keyfile> <patterns.txt> <results.tsv> [statistics.csv] [retrieveDescriptions 0|1] [device] [count]
// Same inputs and outputs: patterns one per line, upper-cased (Main.cpp:1056-1059); results TSV with the header
// "Pattern sequenceIndex sequenceDescription patternPosition" and one line per occurrence in ascending global position;
// a statistics CSV line with the reference's columns. The difference is the loop: all patterns go to the device as ONE
// batch, so per-pattern times are the batch time divided by the number of patterns (stddev and outlier counts are 0).
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>

#include "e2fm_host.h"

using namespace e2fm;

int main(int argc, char **argv) {
    if (argc < 5) {
        std::cerr << "usage: e2fm_locate <index> <keyfile> <patterns> <results.tsv> [statistics.csv] [retrieveDescriptions 0|1] [device] [count]\n";
        return 2;
    }
    try {
        const std::string indexPath = argv[1], keyPath = argv[2], patternsPath = argv[3], resultsPath = argv[4];
        const std::string statsPath = argc > 5 ? argv[5] : "";
        const bool descr = argc > 6 && atoi(argv[6]) != 0;
        const int device = argc > 7 ? atoi(argv[7]) : 0;
        const bool countOnly = argc > 8 && std::string(argv[8]) == "count";
        std::ifstream patternsFile(patternsPath);
        if (!patternsFile.is_open()) throw std::runtime_error("Unable to open patterns file " + patternsPath);   // Main.cpp:1150
        EFMCollection index;
        index.setEncryptionKeyFilePath(keyPath);
        index.setDevice(device);
        auto t0 = std::chrono::steady_clock::now();
        index.load(indexPath);
        double openMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "Index opening time: " << openMs << "ms" << std::endl;
        std::vector<std::string> patterns;
        std::string line;
        uint64_t patternsLength = 0;
        while (std::getline(patternsFile, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            for (char &c : line) c = (char)toupper((unsigned char)c);
            patternsLength += line.size();
            patterns.push_back(line);
        }
        t0 = std::chrono::steady_clock::now();
        std::vector<SearchResult> results = index.searchBatch(patterns, countOnly ? Mode::Count : Mode::Locate, descr);
        double batchMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::ofstream out(resultsPath);
        out << "Pattern" << "\t" << "sequenceIndex" << "\t" << "sequenceDescription" << "\t" << "patternPosition" << std::endl;
        uint64_t occs = 0;
        for (const SearchResult &r : results)
            for (const Occurrence &o : r.occurrences) {
                out << r.pattern << "\t" << o.sequenceIndex << "\t" << o.sequenceDescription << "\t" << o.patternPosition << "\n";
                occs++;
            }
        out.close();
        const double per = patterns.empty() ? 0 : batchMs / (double)patterns.size();
        std::cout << patterns.size() << " patterns, " << occs << " occurrences, batch time " << batchMs << "ms ("
                  << (batchMs > 0 ? 1000.0 * patterns.size() / batchMs : 0) << " patterns/s)" << std::endl;
        std::cout << "Median of search times: " << per << std::endl;
        if (!statsPath.empty()) {
            std::ofstream st(statsPath, std::ofstream::out | std::ofstream::app);
            if (st.tellp() == 0)
                st << "indexFileName,k,bucketSize,markedRowsPercentage,numberOfPatterns,patternsLength,median,mean,stdDev,mean_no,stddev_no,"
                      "numberOfOutliersInTime,medianPerOccurrence,meanPerOccurrence,stdDevPerOccurrence,meanPerOccurrence_no,"
                      "stdDevperOccurrrence_no,numberOfOutliersInTimesPerOccurrence,patternsFilePath"
                   << std::endl;
            const double perOcc = occs ? batchMs / (double)occs : -1;
            st << indexPath << "," << (int)index.getSuperAlphabetOrder() << "," << index.getBucketSize() << "," << (int)index.getMarkedRowsPercentage()
               << "," << patterns.size() << "," << (patterns.empty() ? 0 : patternsLength / patterns.size()) << "," << per << "," << per << ",0," << per
               << ",0,0," << perOcc << "," << perOcc << ",0," << perOcc << ",0,0," << patternsPath << std::endl;
        }
        index.close();
    } catch (std::exception &e) {
        std::cerr << e.what() << std::endl;   // Main.cpp:1682-1684 prints what() and goes on
        return 1;
    }
    return 0;
}
