"""Regenerates tests/golden/ from the UNMODIFIED reference (oracle/_ref/e2fm_ref, compiled from
/root/reference by oracle/Makefile). Run in the authoring container only:

    python tests/golden/make_golden.py

Fixtures (all small, committed):
  kat.txt                 known-answer vectors printed by the reference's own classes (Salsa20 blocks,
                          RandomGenerator streams, scrambling / frequency permutations, int_log2, bitset RLE)
  <name>.fa/.efm/.pat/.res   synthetic collection, the index the reference built from it, the patterns and
                          the reference's search results (binary, see oracle.read_result_file)
  <name>.dump.sha256      sha256 of `e2fm_ref dump` (every decoded internal structure) for the oracle to match
  key.bin                 the 64-byte test key
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from e2fm_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# name, sequence lengths, k, bucket size, mrp, seed, pattern lengths
CASES = [
    ("col3_k4", [30011, 20003, 9001], 4, 256, 2, 7, [4, 5, 7, 8, 9, 11, 12, 13, 16, 20, 24, 31, 32, 41]),
    ("one_k3", [24000], 3, 128, 5, 8, [3, 4, 6, 7, 10, 15, 20]),
    ("col2_k4_nomarks", [12000, 8000], 4, 256, 0, 9, [8, 12, 20]),
    ("col2_k2", [9000, 5000], 2, 512, 4, 10, [2, 3, 5, 8, 13]),
    # the authors' second regime (Main.cpp:623-626): k = 5, bucketSize 8192 (super-buckets of 131072 rows). 19 buckets: a
    # super-bucket boundary at bucket 16, and an even LAST bucket, which the odd-bucket rule excludes (EFMIndex.cpp:898-901)
    ("two_k5", [520003, 230002], 5, 8192, 2, 11, [10, 14, 15, 19, 25, 40, 50]),
]
NO_FASTA = {"two_k5"}     # too big to commit: synth.random_sequences(lens, seed) regenerates it


def main():
    oracle.build(quiet=True)
    assert oracle.have_ref(), "the reference must be compiled (make -C oracle ref)"
    with open(os.path.join(OUT, "kat.txt"), "w") as f:
        f.write(subprocess.run([oracle.REF_BIN, "kat"], capture_output=True, text=True, check=True).stdout)
    key = synth.test_key()
    keyfile = os.path.join(OUT, "key.bin")
    with open(keyfile, "wb") as f:
        f.write(key)
    only = set(sys.argv[1:])
    for name, lens, k, bs, mrp, seed, plens in CASES:
        if only and name not in only:
            continue
        rate = 100 // mrp if mrp else 0
        lens = synth.safe_lengths(lens, k, rate)
        seqs = synth.random_sequences(lens, seed)
        base = os.path.join(OUT, name)
        synth.write_fasta(base + ".fa", seqs)
        oracle.ref_build(base + ".fa", base + ".efm", keyfile, k=k, bucket_size=bs, mrp=mrp, threads=2)
        pats = []
        for L in plens:
            p = synth.sample_patterns(seqs, 24, L, seed * 100 + L, random_frac=0.25)
            pats += [bytes(r) for r in p]
        pats += [b"ACGTNACGT", b"ACGU", b"A" * 9, b"ACGT" * 5, b"T"]   # foreign symbols, repeats, shorter than k
        with open(base + ".pat", "wb") as f:
            f.write(b"\n".join(pats) + b"\n")
        if mrp == 0:
            # the reference itself crashes when it searches an index without marked rows (hat verification calls
            # getRowPosition, EFMIndex.cpp:1862-1899); the fixture only pins the load path + the error behaviour
            print(f"{name}: index only ({os.path.getsize(base + '.efm')} B)")
            continue
        locate = True
        res, _ = oracle.ref_search(base + ".efm", keyfile, base + ".pat", base + ".res", locate=locate, threads=2)
        oracle.ref_dump(base + ".efm", keyfile, base + ".dump")
        with open(base + ".dump", "rb") as f:
            h = hashlib.sha256(f.read()).hexdigest()
        os.remove(base + ".dump")
        with open(base + ".dump.sha256", "w") as f:
            f.write(h + "\n")
        # sanity: the reference agrees with naive search on the '&'-padded text (Main.cpp:1502 naiveSearch)
        text = synth.collection_text(seqs, k)
        bad = 0
        if locate:
            for i, p in enumerate(pats):
                if len(p) < 2 * k:
                    continue   # the reference loses occurrences of patterns shorter than ~2k
                want = synth.naive_find_all(text, p) if all(c in b"ACGT" for c in p) else []
                got = list(res["positions"][int(res["pos_offsets"][i]):int(res["pos_offsets"][i + 1])])
                if [int(x) for x in got] != want:
                    bad += 1
        print(f"{name}: n_patterns={len(pats)} occurrences={len(res['positions'])} index={os.path.getsize(base + '.efm')} B "
              f"reference-vs-naive mismatches={bad}")
        if name in NO_FASTA:
            os.remove(base + ".fa")
        if name in ("col3_k4", "one_k3"):
            # extract / getSequence / extractSnippet (EFMCollection.cpp:43,67; EFMIndex.cpp:2325) from the reference
            rng = np.random.default_rng(seed + 77)
            qs = []
            for si, L_ in enumerate(lens):
                qs.append(f"G {si}")
                for _ in range(12):
                    a = int(rng.integers(0, L_))
                    b = int(min(L_ + k, a + rng.integers(0, 80)))
                    qs.append(f"S {si} {a} {b}")
                qs += [f"S {si} 0 {L_ - 1}", f"S {si} {L_ - 1} {L_ + 2}", f"S {si} 5 5", f"S {si} 1 2", f"S {si} 9 3", f"S {si} {L_ + 3 * k} {L_ + 4 * k}"]
            total = sum((L_ + k - 1) // k * k + k for L_ in lens)
            for _ in range(20):
                a = int(rng.integers(0, total))
                qs.append(f"X {a} {a + int(rng.integers(0, 60))}")
            qs += ["X 0 0", "X 3 3", f"X {total - 3} {total + 5}", f"X {total + 10} {total + 20}", "X 7 2"]
            with open(base + ".extract.q", "w") as f:
                f.write("\n".join(qs) + "\n")
            outs = oracle.ref_extract(base + ".efm", keyfile, base + ".extract.q", base + ".extract.out")
            print(f"  extract goldens: {len(outs)} queries, {sum(o.startswith(b'!') for o in outs)} that throw")
        if name == "col3_k4":
            # the same collection cut into 2 shards by sequence (e2fm_b200/sharded.py): one reference index per shard;
            # the merged shard results must equal the unsharded col3_k4.res
            from e2fm_b200.sharded import plan_shards
            for r, (first, cnt, tbase) in enumerate(plan_shards(lens, 2, k)):
                sl = lens[first:first + cnt]
                n_sh = synth.supertext_length(sl, k)
                q = n_sh // rate
                assert n_sh % rate != 0 and (q & (q - 1)) != 0, "shard hits a reference load-time quirk: change the lengths"
                sb = os.path.join(OUT, f"{name}_shard{r}of2")
                synth.write_fasta(sb + ".fa", seqs[first:first + cnt])
                oracle.ref_build(sb + ".fa", sb + ".efm", keyfile, k=k, bucket_size=bs, mrp=mrp, threads=2)
                os.remove(sb + ".fa")
                print(f"  shard {r}: sequences {first}..{first + cnt - 1}, text base {tbase}, n={n_sh}, index={os.path.getsize(sb + '.efm')} B")


if __name__ == "__main__":
    main()
