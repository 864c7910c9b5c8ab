"""SURVEY.md 8(f) row 4, the bucket text encoder (Bucket::encodeText, Bucket.cpp:442-536 + the symbol loop of Bucket::writeText,
:232-246). The bar: BYTE-IDENTICAL bucket texts against what the unmodified reference builder wrote into the .efm files.
CPU: the oracle's restatement against the golden files. GPU: e2fm_encode_buckets against the same bytes and the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from e2fm_b200 import capi, synth

CASES = ["col3_k4", "one_k3", "col2_k2", "two_k5", "col2_k4_nomarks"]


def bucket_table(oracle_mod, raw, key64):
    """every bucket of a reference-built index: what the builder saw (bucket codes, alphabet size) and what it wrote"""
    ix = oracle_mod.OracleIndex(raw, key64)
    out = []
    for bn in range(ix.info.n_buckets):
        bi = ix.bucket_info(bn)
        bi["codes"] = ix.bucket_codes(bn)
        nbytes = (bi["clen"] * bi["tbits"] + 7) // 8
        assert bi["alpha"] <= 1 or bi["text_bit"] % 8 == 0
        bi["text"] = raw[bi["text_bit"] // 8: bi["text_bit"] // 8 + nbytes] if bi["alpha"] > 1 else b""
        out.append(bi)
    per = ix.info.superbucket_size // ix.info.bucket_size
    meta = {"bucket_size": int(ix.info.bucket_size), "per": int(per), "B": int(ix.info.n_buckets)}
    ix.close()
    return out, meta


@pytest.mark.parametrize("name", CASES)
def test_oracle_encoder_reproduces_the_reference_builder(name, golden, key64, oracle_mod):
    """orc_encode_bucket + orc_pack_sequence == the bytes of every bucket text of the reference-built goldens (k = 2..5, bucket sizes
    128..8192, reversed buckets, a short last bucket)"""
    buckets, meta = bucket_table(oracle_mod, golden[name]["index"], key64)
    assert any(b["is_odd"] for b in buckets) and buckets[-1]["length"] <= meta["bucket_size"]
    for bn, b in enumerate(buckets):
        if b["alpha"] <= 1:
            continue
        seq = oracle_mod.encode_bucket(key64, bn, b["codes"], b["alpha"], bool(b["is_odd"]))
        assert len(seq) == b["clen"], (name, bn)
        assert oracle_mod.pack_sequence(seq, b["tbits"]) == b["text"], (name, bn)


def homopolymer_index(oracle_mod, tmp_path):
    """two sequences with a homopolymer stretch in the middle (at an end the unmodified reference's suffix sorter crashes), small
    buckets: some buckets hold ONE symbol, others two with long runs of rank 0"""
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [np.frombuffer(acgt[rng.integers(0, 4, 3000)].tobytes() + b"A" * 3000 + acgt[rng.integers(0, 4, 3000)].tobytes(), dtype=np.uint8),
            np.frombuffer(acgt[rng.integers(0, 4, 5000)].tobytes() + b"T" * 2400 + acgt[rng.integers(0, 4, 500)].tobytes(), dtype=np.uint8)]
    fa, efm, keyfile = str(tmp_path / "h.fa"), str(tmp_path / "h.efm"), os.path.join(GOLDEN, "key.bin")
    synth.write_fasta(fa, seqs)
    oracle_mod.ref_build(fa, efm, keyfile, k=4, bucket_size=128, mrp=4, threads=2)
    with open(efm, "rb") as f:
        return f.read()


def test_oracle_encoder_single_character_buckets(key64, oracle_mod, tmp_path):
    """the same pin on an index built here by the compiled reference: single-character buckets (no text) and long runs"""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/e2fm_ref not built")
    buckets, meta = bucket_table(oracle_mod, homopolymer_index(oracle_mod, tmp_path), key64)
    assert sum(b["alpha"] == 1 for b in buckets) >= 3
    for bn, b in enumerate(buckets):
        if b["alpha"] > 1:
            seq = oracle_mod.encode_bucket(key64, bn, b["codes"], b["alpha"], bool(b["is_odd"]))
            assert len(seq) == b["clen"] and oracle_mod.pack_sequence(seq, b["tbits"]) == b["text"], bn


def _gpu_check(oracle_mod, raw, key64, tag, splits=(1, 3)):
    buckets, meta = bucket_table(oracle_mod, raw, key64)
    B, bs = meta["B"], meta["bucket_size"]
    for parts in splits:                                  # the whole index in one call, then in runs of buckets
        bounds = [B * i // parts for i in range(parts + 1)]
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            if lo == hi:
                continue
            sel = buckets[lo:hi]
            codes = np.concatenate([b["codes"] for b in sel])
            alpha = np.array([b["alpha"] for b in sel], dtype=np.uint32)
            r = capi.encode_buckets(key64, codes, bs, lo, meta["per"], B, alpha, want_sequence=True)
            row = 0
            for i, b in enumerate(sel):
                what = (tag, lo + i)
                assert int(r["clen"][i]) == b["clen"], what
                off = int(r["offsets"][i])
                assert off % 8 == 0
                assert r["payload"][off:off + len(b["text"])].tobytes() == b["text"], what
                if b["alpha"] > 1:
                    want = oracle_mod.encode_bucket(key64, lo + i, b["codes"], b["alpha"], bool(b["is_odd"]))
                    assert np.array_equal(r["sequence"][row:row + b["clen"]], want), what
                row += b["length"]
    return buckets, meta


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_encoder_reproduces_the_reference_builder(name, golden, key64, oracle_mod):
    """e2fm_encode_buckets: compressed lengths, packed texts (byte for byte what the reference builder wrote) and sequences of every
    bucket of the goldens, whole index and runs of buckets (first_bucket > 0: nonces and reversed buckets follow the absolute number)"""
    if capi.device_count() == 0:
        pytest.skip("no CUDA device")
    _gpu_check(oracle_mod, golden[name]["index"], key64, name)


@pytest.mark.gpu
def test_gpu_encoder_single_character_buckets_and_errors(key64, oracle_mod, tmp_path):
    """a collection with long homopolymer stretches gives buckets with ONE symbol (no text: Bucket.cpp:166-170) and long RLE0 runs;
    bad arguments are refused"""
    if capi.device_count() == 0:
        pytest.skip("no CUDA device")
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/e2fm_ref not built")
    raw = homopolymer_index(oracle_mod, tmp_path)
    buckets, meta = _gpu_check(oracle_mod, raw, key64, "homopolymers", splits=(1, 2))
    assert any(b["alpha"] == 1 for b in buckets), "the fixture was meant to hold single-character buckets"
    assert max(b["length"] / max(b["clen"], 1) for b in buckets if b["alpha"] > 1) > 2, "and long runs"
    # errors
    b0 = max(buckets, key=lambda b: b["clen"])
    codes = b0["codes"].copy()
    with pytest.raises(capi.E2FMError):                   # a code outside the bucket's alphabet
        bad = codes.copy()
        bad[len(bad) // 2] = b0["alpha"]
        capi.encode_buckets(key64, bad, meta["bucket_size"], 0, meta["per"], meta["B"], np.array([b0["alpha"]], dtype=np.uint32))
    with pytest.raises(capi.E2FMError) as e:              # payload buffer too small: the size needed is reported
        capi.encode_buckets(key64, codes, meta["bucket_size"], 0, meta["per"], meta["B"], np.array([b0["alpha"]], dtype=np.uint32), payload_cap=8)
    assert e.value.payload_bytes >= len(b0["text"])
    with pytest.raises(capi.E2FMError):                   # rows do not fit the buckets
        capi.encode_buckets(key64, np.zeros(3 * meta["bucket_size"], dtype=np.uint32), meta["bucket_size"], 0, meta["per"], meta["B"],
                            np.array([2], dtype=np.uint32))


@pytest.mark.gpu
def test_gpu_encoder_config1_full_size(key64, oracle_mod, tmp_path):
    """BASELINE configs[0] (10 Mbp, k = 4, bucket 4096, 611 buckets): every bucket text byte-identical with the reference builder's"""
    if capi.device_count() == 0:
        pytest.skip("no CUDA device")
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/e2fm_ref not built")
    lens = synth.safe_lengths([10_000_000], 4, 50)
    seqs = synth.random_sequences(lens, 42)
    fa, efm, keyfile = str(tmp_path / "c1.fa"), str(tmp_path / "c1.efm"), os.path.join(GOLDEN, "key.bin")
    synth.write_fasta(fa, seqs)
    oracle_mod.ref_build(fa, efm, keyfile, k=4, bucket_size=4096, mrp=2, threads=min(16, os.cpu_count() or 2))
    with open(efm, "rb") as f:
        raw = f.read()
    buckets, meta = _gpu_check(oracle_mod, raw, key64, "config1", splits=(1,))
    assert meta["B"] > 600


def random_bucket(rng, length, alpha, runs):
    """bucket codes that use every symbol of the alphabet when the bucket is long enough; `runs`: long stretches of one symbol"""
    if runs:
        codes = np.repeat(rng.integers(0, alpha, size=max(1, length // 7)), rng.integers(1, 40, size=max(1, length // 7)))[:length]
        codes = np.concatenate([codes, rng.integers(0, alpha, size=length - len(codes))]) if len(codes) < length else codes
    else:
        codes = rng.integers(0, alpha, size=length)
    return np.ascontiguousarray(codes, dtype=np.uint32)


@pytest.mark.parametrize("alpha,length", [(2, 1), (2, 31), (2, 4096), (3, 33), (5, 128), (255, 4096), (256, 4095), (257, 4097), (1026, 8192), (4099, 8192)])
@pytest.mark.parametrize("runs", [False, True])
def test_oracle_encode_decode_round_trip(alpha, length, runs, key64, oracle_mod):
    """size-independent property of the restatement: readTextUntilTo(encodeText(bucket)) == bucket, for even and reversed buckets, over
    alphabets and bucket lengths the goldens do not reach"""
    rng = np.random.default_rng(alpha * 100003 + length + (7 if runs else 0))
    codes = random_bucket(rng, length, alpha, runs)
    for bn, odd in ((0, False), (12, True), (2 ** 33 + 5, False)):
        seq = oracle_mod.encode_bucket(key64, bn, codes, alpha, odd)
        assert len(seq) <= length and (len(seq) == 0 or int(seq.max()) <= alpha)
        back = oracle_mod.decode_sequence(key64, bn, seq, length, alpha, odd)
        assert np.array_equal(back, codes), (alpha, length, runs, bn)


@pytest.mark.gpu
def test_gpu_encoder_random_buckets_against_the_oracle(key64, oracle_mod):
    """e2fm_encode_buckets on synthetic buckets (alphabets 1..4099, long runs, a short last bucket, bucket numbers that start far from
    0): sequences and packed texts equal the oracle's, and the oracle decodes them back to the rows"""
    if capi.device_count() == 0:
        pytest.skip("no CUDA device")
    rng = np.random.default_rng(99)
    for bs, alphas, per, first, total in ((4096, [2, 3, 1, 17, 255, 256, 257, 1, 258, 100, 2, 258], 4, 6, 40),
                                          (8192, [1026, 4099, 1000, 2, 4099], 2, 1000, 1005),
                                          (100, [5, 1, 2, 7, 9, 3, 4], 3, 0, 7)):
        nb = len(alphas)
        lens = [bs] * (nb - 1) + [bs - 37 if first + nb == total else bs]
        parts = [random_bucket(rng, L, a, runs=(i % 2 == 1)) for i, (L, a) in enumerate(zip(lens, alphas))]
        r = capi.encode_buckets(key64, np.concatenate(parts), bs, first, per, total, np.array(alphas, dtype=np.uint32), want_sequence=True)
        row = 0
        for i, (codes, a) in enumerate(zip(parts, alphas)):
            bn = first + i
            odd = (bn % 2 == 0) and (bn % per != 0) and (bn != total - 1)          # EFMIndex.cpp:898-901
            if a == 1:
                assert int(r["clen"][i]) == 0
            else:
                want = oracle_mod.encode_bucket(key64, bn, codes, a, odd)
                got = r["sequence"][row:row + int(r["clen"][i])]
                assert np.array_equal(got, want), (bs, bn, a)
                tb = oracle_mod.int_log2(a)
                off = int(r["offsets"][i])
                packed = oracle_mod.pack_sequence(want, tb)
                assert r["payload"][off:off + len(packed)].tobytes() == packed, (bs, bn, a)
                assert np.array_equal(oracle_mod.decode_sequence(key64, bn, got, len(codes), a, odd), codes)
            row += len(codes)
