"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/e2fm_b200.h declares, and
fails loudly (no fallback) when there is no CUDA device. No compute calls here."""
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from e2fm_b200 import capi


def declared_symbols():
    with open(os.path.join(ROOT, "include", "e2fm_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(e2fm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(L, name), f"{name} is declared in include/e2fm_b200.h but not exported"
    assert sorted(capi.EXPORTS) == decl
    assert b"sm_100a" in L.e2fm_version()


def test_info_struct_layout_matches_header():
    # 8 u64 + 12 u32 + 8 float
    import ctypes
    assert ctypes.sizeof(capi.Info) == 8 * 8 + 12 * 4 + 8 * 4


def test_no_cpu_fallback(golden, key64):
    """Without a GPU every compute entry point must fail with E2FM_ERR_CUDA, never compute on the host."""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.E2FMError) as e:
        capi.DeviceIndex(golden["col3_k4"]["index"], key64)
    assert e.value.code == -1 and "no CPU search path" in str(e.value)
    with pytest.raises(capi.E2FMError):
        capi.debug_keystream(key64, 5, 256, 0, 16)
    with pytest.raises(capi.E2FMError):
        capi.debug_salsa20(key64[:32], 0, 0, 1)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under e2fm_b200/ or include/ may reference it."""
    bad = []
    for base in ("e2fm_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    with open(os.path.join(dp, fn), errors="replace") as f:
                        txt = f.read()
                    if re.search(r"liboracle|e2fm_oracle|from oracle|import oracle|oracle/", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_key_must_be_64_bytes(golden):
    with pytest.raises(ValueError):
        capi.DeviceIndex(golden["col3_k4"]["index"], b"short")


@pytest.mark.parametrize("name", ["col3_k4", "one_k3", "col2_k2", "col2_k4_nomarks"])
def test_host_header_parse_matches_oracle(golden, key64, oracle_mod, name):
    """the host half of e2fm_index_open (magic, geometry, scrambling permutation, C[] decryption, terminators) runs without a
    GPU and must agree with the oracle, which is pinned to the reference's dump"""
    info, Cc, term = capi.parse_header(golden[name]["index"], key64)
    orc = oracle_mod.OracleIndex(golden[name]["index"], key64)
    o = orc.info
    assert (info.text_length, info.k, info.n_symbols, info.marked_rows_percentage) == (o.n, o.k, o.n_sym, o.mrp)
    assert (info.bucket_size, info.superbucket_size, info.n_buckets, info.n_superbuckets) == (
        o.bucket_size, o.superbucket_size, o.n_buckets, o.n_superbuckets)
    assert info.compact_alphabet == o.compact_alphabet and info.n_sequences == o.n_sequences and info.initial_n == o.initial_n
    assert info.quirks == 0
    assert np.array_equal(Cc, orc.header_C())
    assert np.array_equal(term, orc.header_terminators())


def test_host_header_parse_errors(golden, key64):
    idx = golden["col3_k4"]["index"]
    with pytest.raises(capi.E2FMError) as e:
        capi.parse_header(b"XY" + idx[2:], key64)
    assert e.value.code == -2 and "File type is not Encrypted and Compressed Genomic Index" in str(e.value)   # EFMIndex.cpp:1155
    with pytest.raises(capi.E2FMError) as e:
        capi.parse_header(idx, bytes((b + 1) & 0xFF for b in key64))
    assert e.value.code == -2 and "wrong encryption key" in str(e.value)
    for cut in (10, 60, 200, len(idx) // 3):
        with pytest.raises(capi.E2FMError):
            capi.parse_header(idx[:cut], key64)


def test_header_parse_never_reads_outside_the_buffer(tmp_path):
    """truncated and corrupted index files under AddressSanitizer: the host parse works on the caller's unpadded buffer and must
    reject (or parse) every one of them without touching memory outside it"""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "asan_header_fuzz")
    src = [os.path.join(ROOT, "tests", "asan_header_fuzz.cpp"), os.path.join(ROOT, "e2fm_b200", "csrc", "host_format.cpp")]
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe] + src,
                       capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("AddressSanitizer runtime not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("col3_k4", "col2_k2"):
        r = subprocess.run([exe, os.path.join(GOLDEN, name + ".efm"), os.path.join(GOLDEN, "key.bin")], capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        assert "parsed=" in r.stdout


def test_tight_packing_is_the_prefix_of_the_word_packing():
    """e2fm_pack_patterns_tight writes the first ceil(L / 4) bytes of what e2fm_pack_patterns writes (base i at bits 2i), and both
    count the patterns that hold a byte outside ACGT"""
    rng = np.random.default_rng(7)
    for L in (1, 3, 4, 5, 12, 31, 32, 33, 50, 64, 65, 130):
        n = 37
        pats = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))].copy()
        words, bad_w = capi.pack_patterns(pats)
        tight, bad_t = capi.pack_patterns_tight(pats)
        tb = capi.packed_tight_stride(L)
        assert tb == (L + 3) // 4 and bad_w == 0 and bad_t == 0
        assert np.array_equal(words.reshape(n, capi.packed_stride(L))[:, :tb], tight.reshape(n, tb)), L
        assert not words.reshape(n, capi.packed_stride(L))[:, tb:].any()
        pats[5, L // 2] = ord("N")
        pats[9, L - 1] = ord("$")
        assert capi.pack_patterns_tight(pats)[1] == 2 and capi.pack_patterns(pats)[1] == 2


def test_hits_stage_plan_covers_every_batch():
    """the stages e2fm_search_hits cuts a batch into: contiguous, non-empty, covering; automatic mode gives a small batch (a GPU's
    share of a partitioned batch) several stages and a long one a short head; an explicit stage size is kept"""
    for n in (1, 7, 100_000, 131_072, 500_000, 1_250_000, 2_500_000, 4_000_000, 10_000_000, 33_554_433):
        for opt in (0, 7, 64, 8192, 1 << 20):
            if opt and n // opt > 100_000:
                continue
            b = capi.hits_stages(n, opt)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b.astype(np.int64)) > 0), (n, opt, b[:8])
            sizes = np.diff(b.astype(np.int64))
            if opt:
                assert sizes.max() <= max(opt, 1) and len(sizes) >= -(-n // opt)
            else:
                assert sizes.max() <= 1 << 20
                if n >= 1_000_000:
                    assert len(sizes) >= 4, (n, sizes)              # copies and search of a small batch overlap
                if n >= 8 << 20:
                    assert sizes[0] == (1 << 20) // 4              # the search starts after a quarter stage
                assert np.all(sizes[:-1] >= min(n, 65_536)), (n, sizes)       # (only the remainder may be tiny)
