"""GPU parity: every kernel of the CUDA path, through the C ABI, against the oracle (oracle/e2fm_oracle.c)
and the committed reference results (tests/golden/*.res). Integer work: the bar is bit-exact."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_SEARCH_CASES, pack
from e2fm_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    assert capi.device_count() > 0, "GPU tests need a CUDA device"
    return 0


# ---------------------------------------------------------------- K1

def test_salsa20_blocks(gpu, oracle_mod):
    key = bytes([0x80] + [0] * 31)
    got = capi.debug_salsa20(key, 0, 0, 2)
    assert got[:64] == oracle_mod.salsa20_block(key, 0, 0)
    assert got[64:] == oracle_mod.salsa20_block(key, 0, 1)
    rng = np.random.default_rng(1)
    for _ in range(4):
        key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
        nonce = int(rng.integers(0, 2**62))
        first = int(rng.integers(0, 2**40))
        got = capi.debug_salsa20(key, nonce, first, 5)
        for b in range(5):
            assert got[64 * b:64 * b + 64] == oracle_mod.salsa20_block(key, nonce, first + b)


@pytest.mark.parametrize("max_value", [1, 2, 3, 255, 256, 257, 1023, 70000])
def test_keystream_integers(gpu, oracle_mod, key64, max_value):
    for nonce, start, count in [(0, 0, 1), (5, 0, 4200), (77, 100, 333), (2**40 + 3, 12345, 5000)]:
        got = capi.debug_keystream(key64, nonce, max_value, start, count)
        want = oracle_mod.keystream_ints(key64[32:], nonce, max_value, start, count)
        assert np.array_equal(got, want), (max_value, nonce, start)


# ---------------------------------------------------------------- K2 + rank structure

@pytest.fixture(scope="module")
def opened(gpu, golden, key64, oracle_mod):
    out = {}
    for name in GOLDEN_SEARCH_CASES:
        out[name] = (capi.DeviceIndex(golden[name]["index"], key64), oracle_mod.OracleIndex(golden[name]["index"], key64))
    yield out
    for d, _ in out.values():
        d.close()


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
def test_decoded_bwt_and_marks(opened, name):
    dev, orc = opened[name]
    n = int(dev.info.text_length)
    assert n == orc.info.n and dev.info.k == orc.info.k and dev.info.compact_alphabet == orc.info.compact_alphabet
    assert dev.info.n_buckets == orc.info.n_buckets and dev.info.n_sequences == orc.info.n_sequences
    assert dev.info.quirks == 0
    assert np.array_equal(dev.debug_bwt(0, n + 3), orc.bwt(0, n + 3))
    assert np.array_equal(dev.debug_marks(0, n + 3), orc.marks(0, n + 3))


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
def test_lf_and_rank(opened, name):
    dev, orc = opened[name]
    n = int(dev.info.text_length)
    rows = np.arange(n, dtype=np.uint64)
    assert np.array_equal(dev.debug_lf(rows), orc.lf_rows(rows))
    rng = np.random.default_rng(3)
    A = int(dev.info.compact_alphabet)
    codes = rng.integers(0, A, 200000, dtype=np.uint32)
    rr = rng.integers(0, n, 200000, dtype=np.uint64)
    # block boundaries and the ends
    blk = int(dev.info.rank_block)
    edge = np.array([0, 1, blk - 1, blk, blk + 1, n - 1, n - 2], dtype=np.uint64) % n
    rr[:len(edge)] = edge
    assert np.array_equal(dev.debug_rank(codes, rr), orc.rank_rows(codes, rr))


# ---------------------------------------------------------------- K3 + K4 + result pipeline

def check_against_res(dev, patterns, res, locate):
    data, offs = pack(patterns)
    r = dev.search(data, offs, locate)
    out = r.fetch()
    assert np.array_equal(out["counts"], res["counts"])
    if locate:
        assert np.array_equal(out["pos_offsets"], res["pos_offsets"])
        assert np.array_equal(out["positions"], res["positions"])
        assert np.array_equal(out["seq"], res["seq"])
        assert np.array_equal(out["off"], res["off"])
    r.free()


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
@pytest.mark.parametrize("locate", [False, True])
@pytest.mark.parametrize("dense", [0, 1])
def test_search_matches_reference_goldens(opened, golden, name, locate, dense):
    """dense=1: positions from the sa[]/text[] tables built at load; dense=0: per-query walks from the marked rows"""
    dev, _ = opened[name]
    assert dev.info.dense == 1
    dev.set_option(capi.OPT_DENSE, dense)
    try:
        check_against_res(dev, golden[name]["patterns"], golden[name]["res"], locate)
    finally:
        dev.set_option(capi.OPT_DENSE, 1)


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
def test_search_small_chunks_and_windows(opened, golden, name):
    """same results when the batch is cut into many chunks and task windows"""
    dev, _ = opened[name]
    dev.set_option(capi.OPT_CHUNK, 7)
    dev.set_option(capi.OPT_WINDOW, 5)
    try:
        for dense in (0, 1):
            dev.set_option(capi.OPT_DENSE, dense)
            check_against_res(dev, golden[name]["patterns"], golden[name]["res"], True)
    finally:
        dev.set_option(capi.OPT_CHUNK, 0)
        dev.set_option(capi.OPT_WINDOW, 0)
        dev.set_option(capi.OPT_DENSE, 1)


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
def test_fast_path_overflow_is_redone_exactly(opened, golden, name):
    """the fast path sizes its task and result buffers from learned densities; when they are too small the kernels write
    nothing past the end and the host redoes the chunk (tasks) or the batch (results) through the windowed path"""
    dev, _ = opened[name]
    g = golden[name]
    worst = 0
    for density in (0, 1, 100):          # per-mille: 0 -> only the fixed slack (1024 tasks per chunk, 4096 results per batch)
        for chunk in (0, 64, 16):        # 16: every chunk's tasks fit, the whole batch's results may not -> whole-batch redo
            dev.set_option(capi.OPT_DENSITY, density)
            dev.set_option(capi.OPT_CHUNK, chunk)
            try:
                check_against_res(dev, g["patterns"], g["res"], True)
                worst = max(worst, dev.last_batch_counters()["chunks_redone"])
                dev.set_option(capi.OPT_DENSITY, density)
                check_against_res(dev, g["patterns"], g["res"], False)
            finally:
                dev.set_option(capi.OPT_CHUNK, 0)
    assert worst >= 1                    # the redo path really ran
    # and once the densities have been learned the same batch goes through without a redo
    check_against_res(dev, g["patterns"], g["res"], True)
    check_against_res(dev, g["patterns"], g["res"], True)
    assert dev.last_batch_counters()["chunks_redone"] == 0


def test_op_counters_match_oracle(opened, golden):
    """rank / lf / mark operation counts equal the oracle's for the same batch"""
    dev, orc = opened["col3_k4"]
    pats = [p for p in golden["col3_k4"]["patterns"] if len(p) >= 8]
    orc.reset_counters()
    for p in pats:
        orc.search(p, True)
    want = orc.counters()
    data, offs = pack(pats)
    dev.set_option(capi.OPT_DENSE, 0)      # the per-query walks execute the reference's operations one for one
    try:
        r = dev.search(data, offs, True)
        c = dev.last_batch_counters()
        r.free()
    finally:
        dev.set_option(capi.OPT_DENSE, 1)
    assert (c["rank"], c["lf"], c["mark"]) == want


def test_edge_batches(opened, oracle_mod):
    dev, orc = opened["col3_k4"]
    # empty batch
    r = dev.search(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64), True)
    assert r.n == 0 and r.total == 0
    r.free()
    # ragged lengths incl. empty pattern, foreign symbols, lower case (not upper-cased by the library), '&' and '$'
    pats = [b"", b"A", b"ACG", b"ACGT", b"ACGTA", b"acgtacgt", b"ACGTNNNNACGT", b"&&&&", b"ACGT&&&&", b"$$$$", b"TTTTTTTT" * 40,
            b"ACGTACGTACGTACGT", b"GATTACA", b"CATCATCAT"]
    data, offs = pack(pats)
    for locate in (False, True):
        r = dev.search(data, offs, locate)
        out = r.fetch()
        for i, p in enumerate(pats):
            cnt, pos = orc.search(p, True)
            assert int(out["counts"][i]) == cnt, (p, locate)
            if locate:
                a, b = int(out["pos_offsets"][i]), int(out["pos_offsets"][i + 1])
                assert np.array_equal(out["positions"][a:b], pos), p
                for j in range(a, b):
                    assert orc.map_position(int(out["positions"][j])) == (int(out["seq"][j]), int(out["off"][j]))
        r.free()


def test_heavy_hitters_sort_paths(opened):
    """short patterns with thousands of occurrences exercise the CTA / global-memory segment sorts"""
    dev, orc = opened["col2_k2"]
    pats = [b"AC", b"ACG", b"TTT", b"GA", b"CCCC", b"A" * 5]
    data, offs = pack(pats)
    r = dev.search(data, offs, True)
    out = r.fetch()
    for i, p in enumerate(pats):
        cnt, pos = orc.search(p, True)
        a, b = int(out["pos_offsets"][i]), int(out["pos_offsets"][i + 1])
        assert int(out["counts"][i]) == cnt
        assert np.array_equal(out["positions"][a:b], pos), p
    r.free()


def test_fixed_length_entry_point_and_pinned_buffers(opened, golden):
    dev, orc = opened["col3_k4"]
    pats = [p for p in golden["col3_k4"]["patterns"] if len(p) == 20]
    pin = capi.PinnedArray((len(pats), 20), np.uint8)
    pin.array[:] = np.frombuffer(b"".join(pats), dtype=np.uint8).reshape(len(pats), 20)
    r = dev.search_fixed(pin.array, True)
    out = r.fetch()
    for i, p in enumerate(pats):
        cnt, pos = orc.search(p, True)
        a, b = int(out["pos_offsets"][i]), int(out["pos_offsets"][i + 1])
        assert int(out["counts"][i]) == cnt and np.array_equal(out["positions"][a:b], pos)
    r.free()
    pin.free()


@pytest.mark.parametrize("pipe", [3, 64, 0])
def test_pipelined_host_batches(opened, golden, pipe):
    """host batches are copied in and searched chunk by chunk (H2D / kernels / D2H of counts overlap): same answers
    for any pipeline stage size, ragged or fixed-length"""
    dev, orc = opened["col3_k4"]
    g = golden["col3_k4"]
    dev.set_option(capi.OPT_PIPE, pipe)
    try:
        check_against_res(dev, g["patterns"], g["res"], True)          # ragged, through e2fm_search_batch
        pats = [p for p in g["patterns"] if len(p) == 13]
        arr = np.frombuffer(b"".join(pats), dtype=np.uint8).reshape(len(pats), 13).copy()
        for locate in (False, True):
            pin = capi.PinnedArray(len(pats), np.uint64)
            pin.array[:] = 12345
            r = dev.search_stream(arr, locate, pin.array)
            out = r.fetch()
            assert np.array_equal(pin.array, out["counts"])
            for i, p in enumerate(pats):
                cnt, pos = orc.search(p, True)
                assert int(pin.array[i]) == cnt
                if locate:
                    a, b = int(out["pos_offsets"][i]), int(out["pos_offsets"][i + 1])
                    assert np.array_equal(out["positions"][a:b], pos)
            r.free()
            pin.free()
    finally:
        dev.set_option(capi.OPT_PIPE, 0)


def test_error_behaviour(gpu, golden, key64):
    idx = golden["col3_k4"]["index"]
    with pytest.raises(capi.E2FMError) as e:       # bad magic: the reference's message (EFMIndex.cpp:1155)
        capi.DeviceIndex(b"XY" + idx[2:], key64)
    assert e.value.code == -2 and "not Encrypted and Compressed Genomic Index" in str(e.value)
    wrong = bytes((b + 1) & 0xFF for b in key64)
    with pytest.raises(capi.E2FMError) as e:       # wrong key: C[] does not decrypt
        capi.DeviceIndex(idx, wrong)
    assert e.value.code == -2
    with pytest.raises(capi.E2FMError):            # truncated file
        capi.DeviceIndex(idx[:len(idx) // 2], key64)
    nm = capi.DeviceIndex(golden["col2_k4_nomarks"]["index"], key64)
    with pytest.raises(capi.E2FMError) as e:       # EFMIndex.cpp:2305-2306
        nm.search_list([b"ACGTACGT"], True)
    assert e.value.code == -4 and "marked rows" in str(e.value)
    nm.close()


# ---------------------------------------------------------------- BASELINE config 1 at full size

@pytest.fixture(scope="module")
def config1(gpu, key64, oracle_mod, tmp_path_factory):
    """10 Mbp random ACGT, k=4, bucket 4096, mrp=2, built by the compiled reference (oracle/_ref travels to the box)."""
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/e2fm_ref not built")
    d = tmp_path_factory.mktemp("cfg1")
    lens = synth.safe_lengths([10_000_000], 4, 50)
    seqs = synth.random_sequences(lens, 42)
    fa, efm, keyfile = str(d / "c1.fa"), str(d / "c1.efm"), os.path.join(GOLDEN, "key.bin")
    synth.write_fasta(fa, seqs)
    oracle_mod.ref_build(fa, efm, keyfile, k=4, bucket_size=4096, mrp=2, threads=min(16, os.cpu_count() or 2))
    with open(efm, "rb") as f:
        raw = f.read()
    return {"seqs": seqs, "index": raw, "efm": efm, "keyfile": keyfile, "dir": d}


def test_config1_count_locate_vs_reference(config1, key64, oracle_mod):
    """BASELINE.json configs[0]: 10k patterns of length 20, count + locate, against the reference's own CPU search"""
    pats = synth.sample_patterns(config1["seqs"], 10_000, 20, 43)
    pfile = str(config1["dir"] / "p.txt")
    synth.write_patterns(pfile, pats)
    res, _ = oracle_mod.ref_search(config1["efm"], config1["keyfile"], pfile, str(config1["dir"] / "r.res"), True, 2)
    dev = capi.DeviceIndex(config1["index"], key64)
    for dense in (1, 0):
        dev.set_option(capi.OPT_DENSE, dense)
        for locate in (False, True):
            r = dev.search_fixed(np.ascontiguousarray(pats), locate)
            out = r.fetch()
            assert np.array_equal(out["counts"], res["counts"])
            if locate:
                assert np.array_equal(out["pos_offsets"], res["pos_offsets"])
                assert np.array_equal(out["positions"], res["positions"])
                assert np.array_equal(out["seq"], res["seq"]) and np.array_equal(out["off"], res["off"])
            r.free()
    dev.close()


def test_config1_whole_index_decode_and_properties(config1, key64, oracle_mod):
    """every decoded row == the oracle's; LF is a permutation; 1M sampled patterns all found at their sample position"""
    dev = capi.DeviceIndex(config1["index"], key64)
    orc = oracle_mod.OracleIndex(config1["index"], key64)
    n = int(dev.info.text_length)
    assert np.array_equal(dev.debug_bwt(0, n), orc.bwt(0, n))
    assert np.array_equal(dev.debug_marks(0, n), orc.marks(0, n))
    lf = dev.debug_lf(np.arange(n, dtype=np.uint64))
    assert np.array_equal(np.sort(lf), np.arange(n, dtype=np.uint64))      # LF-mapping is a bijection on rows
    # size-independent property at a large batch: a pattern cut from the text at position p is located at p
    seq = config1["seqs"][0]
    rng = np.random.default_rng(5)
    npat, L = 1_000_000, 32
    starts = rng.integers(0, len(seq) - L, npat)
    pats = seq[starts[:, None] + np.arange(L)[None, :]]
    r = dev.search_fixed(np.ascontiguousarray(pats), True)
    out = r.fetch()
    offs = out["pos_offsets"]
    assert np.array_equal(out["counts"], np.diff(offs))
    assert (out["counts"] >= 1).all()
    first = offs[:-1].astype(np.int64)
    # every segment sorted ascending and contains the sampling position
    pos = out["positions"]
    seg = np.repeat(np.arange(npat), np.diff(offs).astype(np.int64))
    assert (np.diff(pos.astype(np.int64))[np.diff(seg) == 0] > 0).all()
    hit = np.zeros(npat, dtype=bool)
    hit[seg[pos == starts[seg].astype(np.uint64)]] = True
    assert hit.all()
    # count-only agrees with locate, and the per-query walks agree with the dense tables
    r2 = dev.search_fixed(np.ascontiguousarray(pats), False)
    assert np.array_equal(r2.fetch()["counts"], out["counts"])
    dev.set_option(capi.OPT_DENSE, 0)
    r3 = dev.search_fixed(np.ascontiguousarray(pats), True)
    out3 = r3.fetch()
    assert np.array_equal(out3["positions"], pos) and np.array_equal(out3["pos_offsets"], offs)
    r3.free()
    dev.set_option(capi.OPT_DENSE, 1)
    # spot-check 300 of them against the oracle
    for i in rng.integers(0, npat, 300):
        cnt, p = orc.search(pats[i].tobytes(), True)
        assert cnt == int(out["counts"][i]) and np.array_equal(p, pos[int(offs[i]):int(offs[i + 1])])
    r.free(); r2.free()
    dev.close()


# ---------------------------------------------------------------- 2 GPUs: shard by sequence, merge over NCCL

def test_sharded_two_gpus(gpu):
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import socket
    import sys
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sharded_gpu_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), worker], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


# ---------------------------------------------------------------- C++ host classes and the locate CLI

HOST_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "e2fm_b200")


def test_cpp_host_classes_match_reference(gpu, golden):
    """EFMIndex::exactMatch / countOccurrences / locateOccurrences, EFMCollection::search, SFMCollection::search,
    driven from C++ exactly as the reference's drivers call them, against the reference's results"""
    exe = os.path.join(HOST_DIR, "e2fm_host_selftest")
    assert os.path.exists(exe), "build e2fm_b200/host (python -c 'import __graft_entry__ as g; g.build()')"
    g = golden["col3_k4"]
    r = subprocess.run([exe, g["base"] + ".efm", os.path.join(GOLDEN, "key.bin"), g["base"] + ".pat"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    info = [l for l in lines if l.startswith("INFO")][0].split()
    assert info[1:4] == ["3", "4", "256"] and info[7] == "2"
    descs = [l.split(" ", 2)[2] for l in lines if l.startswith("DESC")]
    assert descs == ["seq0", "seq1", "seq2"]
    res = g["res"]
    plines = [l for l in lines if l.startswith("P ")]
    assert len(plines) == len(g["patterns"])
    for i, l in enumerate(plines):
        head, efm_part, sfm_part = l.split("|")
        f = head.split()
        a, b = int(res["pos_offsets"][i]), int(res["pos_offsets"][i + 1])
        assert int(f[2]) == int(res["counts"][i]) and int(f[3]) == b - a
        assert [int(x) for x in f[4:]] == [int(x) for x in res["positions"][a:b]]
        want = [f"{int(res['seq'][j])}:{int(res['off'][j])}" for j in range(a, b)]
        assert [":".join(x.split(":")[:2]) for x in efm_part.split()] == want
        assert all(x.split(":")[2] == f"seq{x.split(':')[0]}" for x in efm_part.split())
        assert sfm_part.split() == want
    errs = [l for l in lines if l.startswith("ERR")]
    assert "is not a valid key file" in errs[0]                                    # EncryptionManager.cpp:25
    assert "File type is not Encrypted and Compressed Genomic Index" in errs[1]    # EFMIndex.cpp:1155


def test_locate_cli_formats(gpu, golden, tmp_path):
    """e2fm_locate writes the reference CLI's results TSV (Main.cpp:1052,1074) and statistics CSV (Main.cpp:1138-1146)"""
    exe = os.path.join(HOST_DIR, "e2fm_locate")
    g = golden["col3_k4"]
    pat = tmp_path / "p.txt"
    pat.write_bytes(b"\n".join(p.lower() for p in g["patterns"]) + b"\n")    # the CLI upper-cases (Main.cpp:1056-1059)
    out, st = tmp_path / "r.tsv", tmp_path / "s.csv"
    r = subprocess.run([exe, g["base"] + ".efm", os.path.join(GOLDEN, "key.bin"), str(pat), str(out), str(st), "1"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = out.read_text().splitlines()
    assert rows[0] == "Pattern\tsequenceIndex\tsequenceDescription\tpatternPosition"
    res = g["res"]
    want = []
    for i, p in enumerate(g["patterns"]):
        for j in range(int(res["pos_offsets"][i]), int(res["pos_offsets"][i + 1])):
            want.append(f"{p.decode()}\t{int(res['seq'][j])}\tseq{int(res['seq'][j])}\t{int(res['off'][j])}")
    assert rows[1:] == want
    s = st.read_text().splitlines()
    assert s[0].startswith("indexFileName,k,bucketSize,markedRowsPercentage,numberOfPatterns,patternsLength,median")
    assert s[1].split(",")[1:5] == ["4", "256", "2", str(len(g["patterns"]))]


def test_config1_short_patterns_take_the_wildcard_scan(config1, key64, oracle_mod):
    """len 9-13 patterns on the 10 Mbp index: the shifted super-patterns reach their '?'-prefixed first character with
    tens to hundreds of rows left, so the long intervals go through firstcheck_kernel (row-order scan of bwt16) and the
    short ones through the expanded row tasks; dense tables and per-query walks must both equal the oracle"""
    dev = capi.DeviceIndex(config1["index"], key64)
    orc = oracle_mod.OracleIndex(config1["index"], key64)
    rng = np.random.default_rng(11)
    seq = config1["seqs"][0]
    pats = []
    for L in (9, 10, 11, 12, 13):
        st = rng.integers(0, len(seq) - L, 120)
        pats += [seq[s:s + L].tobytes() for s in st]
    # one k-mer (tens of thousands of occurrences: the global-memory segment sort) and two (hundreds: the CTA sort)
    pats += [b"ACGTACGTACGT", b"AAAAAAAAAAAA", b"ACGTNACGTACG", b"ACGT", b"TTTT", b"GATTACAG", b"ACGTAC"]
    data, offs = pack(pats)
    want = [orc.search(p, True) for p in pats]
    assert max(w[0] for w in want) > 4096
    scanned = 0
    for dense in (1, 0):
        dev.set_option(capi.OPT_DENSE, dense)
        for locate in (True, False):
            r = dev.search(data, offs, locate)
            out = r.fetch()
            if dense:
                scanned += dev.last_batch_counters()["scan_rows"]
            for i, (cnt, pos) in enumerate(want):
                assert int(out["counts"][i]) == cnt, (pats[i], dense, locate)
                if locate:
                    a, b = int(out["pos_offsets"][i]), int(out["pos_offsets"][i + 1])
                    assert np.array_equal(out["positions"][a:b], pos), (pats[i], dense)
            r.free()
    assert scanned > 1000          # the row-order scan really ran
    dev.close()


@pytest.mark.gpu
def test_config1_short_fixed_length_batches_through_the_seed_tables(config1, key64, oracle_mod):
    """fixed-length batches of SHORT patterns (BASELINE configs[4]'s shape: fewer fixed super-characters than the index's own seed
    length) are searched with a shorter seed table, the '?'-suffixed last super-character of a shift enumerated where it is the
    seed's last one: same counts and positions as the oracle, through the two-call form and through e2fm_search_hits"""
    dev = capi.DeviceIndex(config1["index"], key64)
    orc = oracle_mod.OracleIndex(config1["index"], key64)
    try:
        assert dev.info.seed_chars >= 3 and dev.info.vrec_mb > 0
        rng = np.random.default_rng(5)
        seq = config1["seqs"][0]
        seeded_lengths = []
        for L in (8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 23):
            st = rng.integers(0, len(seq) - L, 150)
            pats = [seq[s:s + L].tobytes() for s in st] + [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), L)) for _ in range(30)]
            pats += [b"ACGTACGTACGTACGTACGTACGTACGT"[:L], b"A" * L]
            arr = np.frombuffer(b"".join(pats), dtype=np.uint8).reshape(len(pats), L).copy()
            want = [orc.search(p, True) for p in pats]
            for locate in (True, False):
                r = dev.search_fixed(arr, locate)
                c = dev.last_batch_counters()
                out = r.fetch()
                for i, (cnt, pos) in enumerate(want):
                    assert int(out["counts"][i]) == cnt, (L, pats[i], locate, c["seeded"])
                    if locate:
                        a, b = int(out["pos_offsets"][i]), int(out["pos_offsets"][i + 1])
                        assert np.array_equal(out["positions"][a:b], pos), (L, pats[i], c["seeded"])
                r.free()
            if c["seeded"]:
                seeded_lengths.append(L)
                total = sum(w[0] for w in want)
                packed, bad = capi.pack_patterns(arr)
                assert bad == 0
                o32 = np.zeros(len(pats) + 1, dtype=np.uint32)
                s32 = np.zeros(total + 4, dtype=np.uint32)
                f32 = np.zeros(total + 4, dtype=np.uint32)
                got = dev.search_hits(packed, len(pats), L, True, True, None, o32, s32, f32)
                assert got == total and np.array_equal(np.diff(o32.astype(np.int64)), [w[0] for w in want]), L
                for i in (0, 17, 151, len(pats) - 1):
                    a, b = int(o32[i]), int(o32[i + 1])
                    assert [orc.map_position(int(x)) for x in want[i][1]] == list(zip(s32[a:b].tolist(), f32[a:b].tolist())), (L, i)
        assert 12 in seeded_lengths and 16 in seeded_lengths, seeded_lengths
    finally:
        dev.close()
        orc.close()


# ---------------------------------------------------------------- extract

@pytest.mark.parametrize("name", ["col3_k4", "one_k3"])
def test_extract_snippet_matches_reference_and_oracle(opened, golden, name):
    """e2fm_extract = EFMIndex::extractSnippet (EFMIndex.cpp:2325), including its one-k-mer substr quirk"""
    dev, orc = opened[name]
    base = golden[name]["base"]
    with open(base + ".extract.q", "rb") as f:
        qs = f.read().split(b"\n")[:-1]
    with open(base + ".extract.out", "rb") as f:
        want = f.read().split(b"\n")[:-1]
    n_x = 0
    for q, w in zip(qs, want):
        f_ = q.split()
        if f_[0] != b"X":
            continue
        a, b = int(f_[1]), int(f_[2])
        assert dev.extract_snippet(a, b) == w, q
        assert orc.extract_snippet(a, b) == w, q
        n_x += 1
    assert n_x >= 20
    rng = np.random.default_rng(5)
    total = int(dev.info.text_length) * int(dev.info.k)
    for _ in range(300):
        a = int(rng.integers(0, total))
        b = a + int(rng.integers(0, 200))
        assert dev.extract_snippet(a, b) == orc.extract_snippet(a, b), (a, b)


def test_cpp_host_extract_matches_reference(gpu, golden):
    """EFMCollection::extract / getSequence / EFMIndex::extractSnippet through the C++ host classes, reference goldens"""
    exe = os.path.join(HOST_DIR, "e2fm_host_selftest")
    g = golden["col3_k4"]
    empty = os.path.join(GOLDEN, "empty.pat")
    open(empty, "wb").close()
    try:
        r = subprocess.run([exe, g["base"] + ".efm", os.path.join(GOLDEN, "key.bin"), empty, g["base"] + ".extract.q"], capture_output=True,
                           text=True, timeout=600)
    finally:
        os.remove(empty)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [l[2:] for l in r.stdout.splitlines() if l.startswith("E ") or l == "E"]
    with open(g["base"] + ".extract.out") as f:
        want = f.read().split("\n")[:-1]
    assert got == want


# ---------------------------------------------------------------- seed table + seed search (fixed-length batches, packed patterns)

def _subset(res, idxs, locate=True):
    """the reference's answers for the patterns idxs, re-packed as a CSR of their own"""
    counts = res["counts"][idxs]
    segs = [(int(res["pos_offsets"][i]), int(res["pos_offsets"][i + 1])) for i in idxs]
    offs = np.zeros(len(idxs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([b - a for a, b in segs])
    cat = lambda a: np.concatenate([a[x:y] for x, y in segs]) if segs else a[:0]  # noqa: E731
    return {"counts": counts, "pos_offsets": offs, "positions": cat(res["positions"]), "seq": cat(res["seq"]), "off": cat(res["off"])}


def _by_length(patterns):
    groups = {}
    for i, p in enumerate(patterns):
        groups.setdefault(len(p), []).append(i)
    return groups


def _check_out(out, want, locate, what):
    assert np.array_equal(out["counts"], want["counts"]), what
    if locate:
        assert np.array_equal(out["pos_offsets"], want["pos_offsets"]), what
        assert np.array_equal(out["positions"], want["positions"]), what
        assert np.array_equal(out["seq"], want["seq"]) and np.array_equal(out["off"], want["off"]), what


@pytest.mark.parametrize("name,seed_chars", [(n, j) for n in GOLDEN_SEARCH_CASES for j in (None, 2, 3)] + [("col3_k4", 4)])
def test_seed_search_matches_reference(gpu, golden, key64, name, seed_chars):
    """fixed-length batches go through seed_search_kernel when every shift holds a whole seed: same counts, positions and
    (sequence, offset) as the reference, for the automatic seed length and forced ones, from ASCII and from packed patterns,
    and the same again with the seed search switched off"""
    g = golden[name]
    if seed_chars is None:
        os.environ.pop("E2FM_SEED_CHARS", None)
    else:
        os.environ["E2FM_SEED_CHARS"] = str(seed_chars)
    try:
        dev = capi.DeviceIndex(g["index"], key64)
    finally:
        os.environ.pop("E2FM_SEED_CHARS", None)
    try:
        j = int(dev.info.seed_chars)
        assert j >= 2 and (seed_chars is None or j == seed_chars)
        # the table against the rank structure: random seeds + every seed of a few rows' own first characters
        rng = np.random.default_rng(11)
        A, n = int(dev.info.compact_alphabet), int(dev.info.text_length)
        codes = rng.integers(0, A, (4000, j), dtype=np.uint32)
        bwt = dev.debug_bwt(0, n)
        _, C, _ = capi.parse_header(g["index"], key64)
        C = C.astype(np.int64)
        # 1000 seeds that do occur: j consecutive super-characters of the text, read off the BWT (F[r] preceded by bwt[r],
        # preceded by bwt[LF(r)], ...)
        rr = rng.integers(0, n, 1000).astype(np.uint64)
        codes[:1000, j - 1] = np.searchsorted(C, rr.astype(np.int64), side="right") - 1
        for q in range(j - 2, -1, -1):
            codes[:1000, q] = bwt[rr.astype(np.int64)]
            rr = dev.debug_lf(rr)
        got = dev.debug_seed(codes)
        c = codes[:, j - 1].astype(np.int64)
        s, e = C[c].copy(), C[c + 1].copy()
        for q in range(j - 2, -1, -1):                               # EFMIndex::backwardStep (:1966) through the rank hook
            c = codes[:, q].astype(np.uint32)
            live = e > s
            lo_rows = np.where(s > 0, s - 1, 0).astype(np.uint64)
            ns = np.where(s > 0, dev.debug_rank(c, lo_rows).astype(np.int64), C[c.astype(np.int64)])
            ne = dev.debug_rank(c, np.where(live, e - 1, 0).astype(np.uint64)).astype(np.int64)
            s, e = np.where(live, ns, s), np.where(live, ne, e)
        rows = np.maximum(e - s, 0)
        exact = got[:, 2] == 1
        assert np.array_equal(got[exact, 1].astype(np.int64), rows[exact])
        hit = exact & (rows > 0)
        assert np.array_equal(got[hit, 0].astype(np.int64), s[hit])
        assert np.all(rows[:1000] > 0)
        if A ** j >= 4 * n:                                           # about one row per seed or fewer: counters hardly ever saturate
            assert hit.sum() >= 900 and exact.sum() > len(codes) * 0.9
        seeded_any = False
        for L, idxs in sorted(_by_length(g["patterns"]).items()):
            if L == 0:
                continue
            arr = np.frombuffer(b"".join(g["patterns"][i] for i in idxs), dtype=np.uint8).reshape(len(idxs), L).copy()
            want = _subset(g["res"], idxs)
            for locate in (False, True):
                try:
                    r = dev.search_fixed(arr, locate)
                except capi.E2FMError as e:
                    raise AssertionError(f"search_fixed failed for length {L}, {len(idxs)} patterns, locate={locate}: {e}")
                seeded = dev.last_batch_counters()["seeded"]
                _check_out(r.fetch(), want, locate, (name, L, "seed" if seeded else "generic"))
                r.free()
                if seeded:
                    seeded_any = True
                    dev.set_option(capi.OPT_SEED, 0)
                    r = dev.search_fixed(arr, locate)
                    assert dev.last_batch_counters()["seeded"] == 0
                    _check_out(r.fetch(), want, locate, (name, L, "seed off"))
                    r.free()
                    dev.set_option(capi.OPT_SEED, 1)
                    # packed entry point: only batches of pure ACGT patterns
                    keep = [q for q, i in enumerate(idxs) if all(ch in b"ACGT" for ch in g["patterns"][i])]
                    if not keep:
                        continue
                    sub = np.ascontiguousarray(arr[keep])
                    packed, bad = capi.pack_patterns(sub)
                    assert bad == 0
                    cnt = np.zeros(len(keep), dtype=np.uint64)
                    r = dev.search_packed(packed, len(keep), L, locate, cnt)
                    wsub = _subset(g["res"], [idxs[q] for q in keep])
                    out = r.fetch()
                    _check_out(out, wsub, locate, (name, L, "packed"))
                    assert np.array_equal(cnt, wsub["counts"])
                    if locate:
                        c32 = np.zeros(len(keep), dtype=np.uint32)
                        o32 = np.zeros(len(keep) + 1, dtype=np.uint32)
                        s32 = np.zeros(r.total, dtype=np.uint32)
                        f32 = np.zeros(r.total, dtype=np.uint32)
                        r.fetch_compact(c32, o32, s32, f32)
                        assert np.array_equal(c32, wsub["counts"]) and np.array_equal(o32, wsub["pos_offsets"])
                        assert np.array_equal(s32, wsub["seq"]) and np.array_equal(f32, wsub["off"])
                    r.free()
        assert seeded_any, "no batch of this golden went through the seed search"
    finally:
        dev.close()


def test_seed_search_repeats_and_odd_symbols(gpu, golden, key64, oracle_mod):
    """a pattern that occurs more than once or meets a saturated seed counter is finished by the seed search itself (backward
    steps in place of the table entry); one that holds an index symbol outside ACGT is answered by the generic path; a foreign
    symbol means no match at all (EFMIndex.cpp:1702-1717)"""
    g = golden["col3_k4"]
    dev = capi.DeviceIndex(g["index"], key64)
    orc = oracle_mod.OracleIndex(g["index"], key64)
    try:
        seqs = synth.random_sequences(synth.safe_lengths([30011, 20003, 9001], 4, 50), 7)   # tests/golden/make_golden.py
        text = synth.collection_text(seqs, 4)
        L = 12
        pats = [text[100:100 + L], text[30009:30009 + L], b"ACGTACGTACGT", b"A" * L, b"ACGTNNNNACGT", b"ACGT&&&&ACGT",
                text[30000:30000 + L], b"acgtacgtacgt", text[59020:59020 + L], text[5:5 + L]]
        # (a pattern of '$' only is left out: the unmodified reference itself dies with SIGSEGV on it, and so does the oracle)
        # 12-mers that occur twice in the text
        seen, twice = {}, []
        whole = seqs[0].tobytes()
        for i in range(len(whole) - L):
            w = whole[i:i + L]
            if w in seen and len(twice) < 20:
                twice.append(w)
            seen[w] = i
        pats += twice
        arr = np.frombuffer(b"".join(pats), dtype=np.uint8).reshape(len(pats), L).copy()
        for locate in (False, True):
            r = dev.search_fixed(arr, locate)
            c = dev.last_batch_counters()
            assert c["seeded"] == 1
            out = r.fetch()
            for i, p in enumerate(pats):
                cnt, pos = orc.search(p, True)
                assert int(out["counts"][i]) == cnt, (p, locate)
                if locate:
                    a, b = int(out["pos_offsets"][i]), int(out["pos_offsets"][i + 1])
                    assert np.array_equal(out["positions"][a:b], pos), p
            assert c["complex_patterns"] >= 1       # the '&' pattern
            r.free()
        # tiny result store: the seed search flags the overflow and is run again with room
        dev.set_option(capi.OPT_DENSITY, 1)
        big = np.ascontiguousarray(np.tile(arr[:1], (5000, 1)))
        r = dev.search_fixed(big, True)
        out = r.fetch()
        assert r.total == 5000 and np.all(out["counts"] == 1) and np.all(out["positions"] == out["positions"][0])
        assert dev.last_batch_counters()["chunks_redone"] >= 1
        r.free()
    finally:
        dev.close()
        orc.close()


# ---------------------------------------------------------------- e2fm_search_hits: host to host, pipelined

def _hits_call(dev, arr, L, packed, locate, cap):
    n = len(arr)
    pats = arr
    if packed:
        pats, bad = capi.pack_patterns_tight(arr) if packed == 2 else capi.pack_patterns(arr)
        assert bad == 0
    c32 = np.full(n, 0xDEADBEEF, dtype=np.uint32)
    o32 = np.full(n + 1, 0xDEADBEEF, dtype=np.uint32) if locate else None
    s32 = np.full(cap, 0xDEADBEEF, dtype=np.uint32) if locate else None
    f32 = np.full(cap, 0xDEADBEEF, dtype=np.uint32) if locate else None
    total = dev.search_hits(np.ascontiguousarray(pats).reshape(-1), n, L, packed, locate, c32, o32, s32, f32)
    return total, c32, o32, s32, f32


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
@pytest.mark.parametrize("stage,vrec", [(0, 1), (64, 1), (7, 1), (64, 0)])
def test_search_hits_matches_reference(gpu, golden, key64, name, stage, vrec):
    """e2fm_search_hits (pipeline stages of 2^20, 64 and 7 patterns; candidate rows verified through the 32-byte verification
    records and, with E2FM_VREC=0, through sa[] / text[]): counts, occurrence offsets, sequence index and offset in the sequence of
    every occurrence equal the reference's, from ASCII, from packed (16-byte words) and from tightly packed patterns (ceil(L / 4)
    bytes each, widened on the device), count and locate"""
    g = golden[name]
    os.environ["E2FM_VREC"] = str(vrec)
    try:
        dev = capi.DeviceIndex(g["index"], key64)
    finally:
        os.environ.pop("E2FM_VREC", None)
    assert (dev.info.vrec_mb > 0 or dev.info.text_length * 32 < (1 << 20)) if vrec else dev.info.vrec_mb == 0
    try:
        dev.set_option(capi.OPT_HITS_STAGE, stage)
        piped_any = False
        for L, idxs in sorted(_by_length(g["patterns"]).items()):
            if L == 0:
                continue
            arr = np.frombuffer(b"".join(g["patterns"][i] for i in idxs), dtype=np.uint8).reshape(len(idxs), L).copy()
            want = _subset(g["res"], idxs)
            acgt = all(all(ch in b"ACGT" for ch in g["patterns"][i]) for i in idxs)
            for packed in ((0, 1, 2) if acgt else (0,)):
                for locate in (False, True):
                    try:
                        total, c32, o32, s32, f32 = _hits_call(dev, arr, L, packed, locate, len(want["seq"]) + 3)
                    except capi.E2FMError:
                        assert packed, "only the packed form may be refused (batch shapes the seed search does not take)"
                        continue
                    what = (name, L, packed, locate, stage)
                    assert np.array_equal(c32, want["counts"].astype(np.uint32)), what
                    stages = dev.last_batch_counters()["hits_stages"]
                    piped_any |= stages > 0
                    if stage and stages:
                        assert stages >= -(-len(idxs) // stage)      # full stages, then a tapered tail
                    if locate:
                        assert total == len(want["seq"]), what
                        assert np.array_equal(o32, want["pos_offsets"].astype(np.uint32)), what
                        assert np.array_equal(s32[:total], want["seq"]) and np.array_equal(f32[:total], want["off"].astype(np.uint32)), what
                        assert np.all(s32[total:] == 0xDEADBEEF) and np.all(f32[total:] == 0xDEADBEEF), what
        assert piped_any, "no batch of this golden went through the pipeline"
    finally:
        dev.close()


@pytest.mark.gpu
def test_search_hits_undersized_stores_and_capacity(gpu, golden, key64):
    """stores sized from earlier batches that turn out too small make the call fall back to the exact-size path (same answers);
    a caller's capacity that is too small is an error that reports the capacity needed"""
    g = golden["col3_k4"]
    dev = capi.DeviceIndex(g["index"], key64)
    try:
        groups = _by_length(g["patterns"])
        L = 24                                            # long enough for a whole seed in every shift
        idxs = [i for i in groups[L] if all(ch in b"ACGT" for ch in g["patterns"][i])]   # (others go to the generic path anyway)
        arr = np.frombuffer(b"".join(g["patterns"][i] for i in idxs), dtype=np.uint8).reshape(len(idxs), L).copy()
        reps = -(-40000 // len(idxs))
        big = np.ascontiguousarray(np.tile(arr, (reps, 1)))
        want = _subset(g["res"], idxs * reps)
        assert len(want["seq"]) > 30000
        dev.set_option(capi.OPT_HITS_STAGE, 8192)
        dev.set_option(capi.OPT_DENSITY, 1)          # one task / result per 1000 patterns: every stage overflows
        total, c32, o32, s32, f32 = _hits_call(dev, big, L, False, True, len(want["seq"]) + 8)
        assert dev.last_batch_counters()["hits_stages"] == 0, "the call must have fallen back"
        assert total == len(want["seq"]) and np.array_equal(c32, want["counts"].astype(np.uint32))
        assert np.array_equal(o32, want["pos_offsets"].astype(np.uint32)) and np.array_equal(s32[:total], want["seq"])
        # the same through the tight packing: its fallback widens on the host
        dev.set_option(capi.OPT_DENSITY, 1)
        total, c32, o32, s32, f32 = _hits_call(dev, big, L, 2, True, len(want["seq"]) + 8)
        assert dev.last_batch_counters()["hits_stages"] == 0
        assert total == len(want["seq"]) and np.array_equal(c32, want["counts"].astype(np.uint32))
        assert np.array_equal(o32, want["pos_offsets"].astype(np.uint32)) and np.array_equal(f32[:total], want["off"].astype(np.uint32))
        # now the densities are learned: the same call is pipelined
        total, c32, o32, s32, f32 = _hits_call(dev, big, L, False, True, len(want["seq"]) + 8)
        c = dev.last_batch_counters()
        assert c["hits_stages"] > 1, ("gave up", c["hits_gave_up"], "tasks", c["tasks"], "occ", c["occurrences"], "complex", c["complex_patterns"])
        assert total == len(want["seq"]) and np.array_equal(f32[:total], want["off"].astype(np.uint32))
        assert np.array_equal(o32, want["pos_offsets"].astype(np.uint32)) and np.array_equal(s32[:total], want["seq"])
        with pytest.raises(capi.E2FMError):
            _hits_call(dev, big, L, False, True, len(want["seq"]) - 1)
        assert dev.last_hits_total == len(want["seq"])
        # empty batch
        e = np.zeros((0, L), dtype=np.uint8)
        total, c32, o32, s32, f32 = _hits_call(dev, e, L, False, True, 4)
        assert total == 0 and o32[0] == 0
    finally:
        dev.close()
