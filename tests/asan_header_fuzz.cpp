// Host header parse (e2fm_b200/csrc/host_format.cpp) under AddressSanitizer: every truncation of an index file, and
// random byte corruptions of its header region, must end in a clean exception or a clean parse — never in a read outside
// the caller's (unpadded, exactly-sized heap) buffer. Built and run by tests/test_capi_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../e2fm_b200/csrc/host_format.h"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + got);
    fclose(f);
    uint8_t key[64];
    f = fopen(argv[2], "rb");
    if (!f || fread(key, 1, 64, f) != 64) return 2;
    fclose(f);
    unsigned ok = 0, thrown = 0;
    auto run = [&](const uint8_t *src, size_t len) {
        uint8_t *exact = (uint8_t *)malloc(len ? len : 1);     // exactly sized: ASan traps a read one byte past it
        memcpy(exact, src, len);
        try {
            e2fm::HostImage img;
            e2fm::parse_index_header(exact, len, key, img);
            ok++;
        } catch (const std::exception &) { thrown++; }
        free(exact);
    };
    run(file.data(), file.size());
    if (ok != 1) { fprintf(stderr, "the intact file does not parse\n"); return 1; }
    // truncations: every length in the header region, then a sweep over the rest, then the last bytes
    for (size_t len = 0; len < file.size() && len < 400; len++) run(file.data(), len);
    for (size_t len = 400; len < file.size(); len += 97) run(file.data(), len);
    for (size_t back = 1; back < 64 && back < file.size(); back++) run(file.data(), file.size() - back);
    // corruptions of the fixed header, the bookmarks and the start tables
    uint64_t x = 88172645463325252ull;
    std::vector<uint8_t> m = file;
    for (int it = 0; it < 3000; it++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const size_t region = it % 3 == 0 ? 120 : file.size();
        const size_t at = (size_t)(x % region);
        const uint8_t old = m[at];
        m[at] = (uint8_t)(x >> 32);
        run(m.data(), m.size());
        if (it % 4) m[at] = old;                               // some corruptions accumulate
    }
    printf("parsed=%u rejected=%u\n", ok, thrown);
    return 0;
}
