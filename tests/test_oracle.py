"""The oracle (oracle/e2fm_oracle.c) pinned against the reference: known-answer vectors printed by the
reference's own classes, sha256 of its dump of every decoded structure, its search results, and naive search."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_SEARCH_CASES
from e2fm_b200 import synth


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "kat.txt")) as f:
        return dict(l.rstrip("\n").split("=", 1) for l in f if "=" in l)


def ints(s):
    return [int(x) for x in s.split()]


def test_salsa20_ecrypt_vector(oracle_mod, kat):
    key = bytes([0x80] + [0] * 31)   # ECRYPT Set 1 vector 0
    assert oracle_mod.salsa20_block(key, 0, 0).hex().upper() == kat["salsa20_block0"]
    assert oracle_mod.salsa20_block(key, 0, 1).hex().upper() == kat["salsa20_block1"]


def test_keystream_integers(oracle_mod, kat, key64):
    poly = key64[32:]
    assert list(oracle_mod.keystream_ints(poly, 5, 256, 0, 16)) == ints(kat["rg_nonce5_max256"])
    assert list(oracle_mod.keystream_ints(poly, 5, 256, 100, 4)) == ints(kat["rg_nonce5_max256_start100"])
    assert list(oracle_mod.keystream_ints(poly, 77, 255, 0, 16)) == ints(kat["rg_nonce77_max255"])
    got = "".join(str(x) for x in oracle_mod.keystream_ints(poly, 123456, 1, 0, len(kat["rg_nonce123456_max1"])))
    assert got == kat["rg_nonce123456_max1"]


@pytest.mark.parametrize("card", [1296, 625, 7776])
def test_scrambling_key(oracle_mod, kat, key64, card):
    assert list(oracle_mod.scrambling_key(key64, card)) == ints(kat[f"scramble{card}"])


def test_frequency_secrets(oracle_mod, kat, key64):
    assert list(oracle_mod.freq_permutation(key64, 259)) == ints(kat["freqperm259"])
    assert list(oracle_mod.freq_keystream(key64, 259, 2**22 - 1)) == ints(kat["freqks259_22"])
    assert list(oracle_mod.freq_keystream(key64, 259, 2**30 - 1)) == ints(kat["freqks259_30"])


def test_int_log2(oracle_mod, kat):
    for pair in kat["int_log2"].split():
        u, v = pair.split(":")
        assert oracle_mod.int_log2(int(u)) == int(v)


def test_bitset_rle(oracle_mod, kat):
    comp = [int(c) for c in kat["rle_compress"]]
    want = [int(c) for c in kat["rle_decompress"]]
    assert list(oracle_mod.rle_decompress(comp)) == want


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
def test_dump_matches_reference(oracle_mod, golden, key64, name, tmp_path):
    ix = oracle_mod.OracleIndex(golden[name]["index"], key64)
    out = str(tmp_path / "dump.bin")
    ix.dump(out)
    with open(out, "rb") as f:
        got = hashlib.sha256(f.read()).hexdigest()
    with open(golden[name]["base"] + ".dump.sha256") as f:
        assert got == f.read().strip()


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
def test_search_matches_reference(oracle_mod, golden, key64, name):
    g = golden[name]
    ix = oracle_mod.OracleIndex(g["index"], key64)
    res = g["res"]
    for i, p in enumerate(g["patterns"]):
        cnt, pos = ix.search(p, True)
        a, b = int(res["pos_offsets"][i]), int(res["pos_offsets"][i + 1])
        assert cnt == int(res["counts"][i]), (name, i, p)
        assert np.array_equal(pos, res["positions"][a:b]), (name, i, p)
        for j in range(a, b):
            assert ix.map_position(int(res["positions"][j])) == (int(res["seq"][j]), int(res["off"][j]))


def test_search_matches_naive(oracle_mod, golden, key64):
    """the authors' own intended ground truth (naiveSearch, Main.cpp:1502-1510) on the '&'-joined text"""
    g = golden["col3_k4"]
    seqs = []
    with open(g["base"] + ".fa", "rb") as f:
        for line in f.read().split(b"\n"):
            if line and not line.startswith(b">"):
                seqs.append(np.frombuffer(line, dtype=np.uint8))
    text = synth.collection_text(seqs, 4)
    ix = oracle_mod.OracleIndex(g["index"], key64)
    checked = 0
    for p in g["patterns"]:
        if len(p) < 8 or any(c not in b"ACGT" for c in p):
            continue   # shorter than 2k: the reference (and so the oracle) loses occurrences, SURVEY.md quirk 1
        cnt, pos = ix.search(p, True)
        assert [int(x) for x in pos] == synth.naive_find_all(text, p)
        assert cnt == len(pos)
        checked += 1
    assert checked > 200


def test_oracle_rejects_bad_input(oracle_mod, golden, key64):
    with pytest.raises(RuntimeError):
        oracle_mod.OracleIndex(b"XX" + golden["col3_k4"]["index"][2:], key64)
    ix = oracle_mod.OracleIndex(golden["col2_k4_nomarks"]["index"], key64)
    with pytest.raises(RuntimeError):
        ix.search(b"ACGTACGTACGT", True)


@pytest.mark.skipif(not os.path.exists("/root/reference"), reason="the reference sources are only in the authoring container")
def test_reference_still_agrees_with_committed_goldens(oracle_mod, golden, key64, tmp_path):
    """re-run the compiled reference on one fixture: the committed .res must be reproducible"""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref not built")
    g = golden["one_k3"]
    keyfile = os.path.join(GOLDEN, "key.bin")
    res, _ = oracle_mod.ref_search(g["base"] + ".efm", keyfile, g["base"] + ".pat", str(tmp_path / "r.res"), True, 2)
    assert np.array_equal(res["counts"], g["res"]["counts"])
    assert np.array_equal(res["positions"], g["res"]["positions"])


def _collection_extract(ix, terms, k, line):
    """EFMCollection::extract / getSequence (EFMCollection.cpp:43-95) on top of extractSnippet, as the host classes do it"""
    f = line.split()
    def cut(b):
        p = b.find(b"&")
        return b[:p] if p >= 0 else b
    if f[0] == b"X":
        return ix.extract_snippet(int(f[1]), int(f[2]))
    q = int(f[1])
    whole_from = terms[q - 1] + k if q > 0 else 0
    whole_to = terms[q]
    if f[0] == b"G":
        # getSequence holds find()'s result in a uint64_t and tests `p > -1`, which is never true: the '&' padding stays
        return ix.extract_snippet(whole_from, whole_to)
    a, b = int(f[2]), int(f[3])
    max_len = whole_to - whole_from + 1
    if a > max_len - 1:
        return b"!start index out of range"
    if b > max_len - 1:
        return b"!end index out of range"
    return cut(ix.extract_snippet(whole_from + a, whole_from + b))


@pytest.mark.parametrize("name,k,terms", [("col3_k4", 4, [30012, 50020, 59028]), ("one_k3", 3, None)])
def test_extract_matches_reference(oracle_mod, golden, key64, name, k, terms):
    """EFMIndex::extractSnippet / EFMCollection::extract / getSequence against the reference's own output"""
    ix = oracle_mod.OracleIndex(golden[name]["index"], key64)
    base = golden[name]["base"]
    with open(base + ".extract.q", "rb") as f:
        qs = f.read().split(b"\n")[:-1]
    with open(base + ".extract.out", "rb") as f:
        want = f.read().split(b"\n")[:-1]
    if terms is None:   # single sequence: its terminator is the padded length
        with open(base + ".fa", "rb") as f:
            L = len(f.read().split(b"\n")[1])
        terms = [(L + k - 1) // k * k]
    assert len(qs) == len(want) > 40
    for q, w in zip(qs, want):
        assert _collection_extract(ix, terms, k, q) == w, q
