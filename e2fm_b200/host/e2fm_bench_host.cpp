// e2fm_bench_host — end-to-end throughput of the C++ host classes (e2fm_host.h) on one batch, for bench.py's `e2e_cpp`:
//   e2fm_bench_host <index.efm> <keyfile> <patterns.raw> <n> <length> <mode: count|locate> <reps> [device] [skip]
// patterns.raw = n x length ASCII bytes after `skip` header bytes (a .npy file's header, say). Timed, host wall clock, std::vector<std::string> in:
//   compact   EFMIndex::searchBatchCompact  -> 32-bit CSR in page-locked memory
//   objects   EFMCollection::searchBatch    -> std::vector<SearchResult>, the reference's result type (FMCollection.h:47-61)
// One JSON line on stdout.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>

#include "e2fm_host.h"

using namespace e2fm;

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    if (argc < 8) {
        std::cerr << "usage: e2fm_bench_host <index> <keyfile> <patterns.raw> <n> <length> <count|locate> <reps> [device]\n";
        return 2;
    }
    try {
        const uint64_t n = strtoull(argv[4], nullptr, 10);
        const uint32_t len = (uint32_t)atoi(argv[5]);
        const Mode mode = std::string(argv[6]) == "count" ? Mode::Count : Mode::Locate;
        const int reps = std::max(1, atoi(argv[7]));
        EFMCollection index;
        index.setEncryptionKeyFilePath(argv[2]);
        if (argc > 8) index.setDevice(atoi(argv[8]));
        double t0 = now_ms();
        index.load(argv[1]);
        const double loadMs = now_ms() - t0;
        std::vector<std::string> patterns(n);
        {
            std::ifstream f(argv[3], std::ifstream::binary);
            if (!f) throw std::runtime_error("cannot open the patterns file");
            if (argc > 9) f.seekg((std::streamoff)strtoull(argv[9], nullptr, 10));
            std::string buf((size_t)len, '\0');
            for (uint64_t i = 0; i < n; i++) {
                f.read(&buf[0], len);
                if (!f) throw std::runtime_error("patterns file too short");
                patterns[i] = buf;
            }
        }
        index.searchBatchCompact(patterns, mode);                  // warm-up: work buffers, learned densities
        double best = 1e300, sum = 0, pack = 0, dev = 0;
        uint64_t total = 0, checksum = 0;
        bool packed = false;
        for (int r = 0; r < reps; r++) {
            t0 = now_ms();
            CompactBatchResult c = index.searchBatchCompact(patterns, mode);
            const double ms = now_ms() - t0;
            best = std::min(best, ms);
            sum += ms;
            pack += c.packMs;
            dev += c.deviceMs[6];
            total = c.total;
            packed = c.packed;
            checksum = 0;
            if (mode == Mode::Locate) for (uint64_t i = 0; i < c.total; i += 997) checksum += c.sequenceOffset[i] + c.sequenceIndex[i];
            else for (uint64_t i = 0; i < n; i += 997) checksum += c.counts[i];
        }
        const int reps_obj = std::min(reps, 2);
        double obj_best = 1e300, obj_sum = 0;
        uint64_t obj_occ = 0;
        for (int r = 0; r < reps_obj; r++) {
            t0 = now_ms();
            std::vector<SearchResult> v = index.searchBatch(patterns, mode, false);
            const double ms = now_ms() - t0;
            obj_best = std::min(obj_best, ms);
            obj_sum += ms;
            obj_occ = 0;
            for (uint64_t i = 0; i < v.size(); i += 101) obj_occ += v[i].occurrences.size();
        }
        printf("{\"patterns\": %llu, \"pattern_len\": %u, \"mode\": \"%s\", \"occurrences\": %llu, \"packed\": %s, \"reps\": %d, "
               "\"compact\": {\"api\": \"e2fm::EFMIndex::searchBatchCompact(vector<string>) -> 32-bit CSR\", \"ms_mean\": %.3f, \"ms_best\": %.3f, "
               "\"patterns_per_s\": %.1f, \"host_marshalling_ms\": %.3f, \"device_ms\": %.3f}, "
               "\"objects\": {\"api\": \"e2fm::EFMCollection::searchBatch(vector<string>) -> vector<SearchResult>\", \"ms_mean\": %.3f, "
               "\"ms_best\": %.3f, \"patterns_per_s\": %.1f, \"reps\": %d}, \"index_load_ms\": %.1f, \"checksum\": %llu, \"sampled_occurrences\": %llu}\n",
               (unsigned long long)n, len, mode == Mode::Count ? "count" : "locate", (unsigned long long)total, packed ? "true" : "false", reps,
               sum / reps, best, 1000.0 * (double)n / (sum / reps), pack / reps, dev / reps, obj_sum / reps_obj, obj_best,
               1000.0 * (double)n / (obj_sum / reps_obj), reps_obj, loadMs, (unsigned long long)checksum, (unsigned long long)obj_occ);
        index.close();
    } catch (std::exception &e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
