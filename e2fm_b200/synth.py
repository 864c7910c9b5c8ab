"""Synthetic inputs of the shapes BASELINE.json names: uniform random ACGT
multi-FASTA collections, the fixed 64-byte test key, and pattern sets sampled from the text.

Pure host-side data generation (numpy); used by tests/, bench.py and the golden-fixture script.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def test_key() -> bytes:
    """The 64-byte key every fixture uses: key[i] = (7*i+3) & 0xFF."""
    return bytes((7 * i + 3) & 0xFF for i in range(64))


def supertext_length(seq_lens, k: int) -> int:
    """n = paddedTotal/k + 1 with every sequence padded to a multiple of k and followed by a
    k-long terminator (reference: EFMCollection.cpp:187-233, EFMIndex.cpp:563)."""
    total = 0
    for L in seq_lens:
        total += (L + k - 1) // k * k + k
    return total // k + 1


def safe_lengths(seq_lens, k: int, rate: int):
    """Lengthen the last sequence by k bases until the super-text length n avoids the two
    reference load-time bugs: n % rate == 0, and floor(n/rate)
    a power of two (which makes the reference mis-size the marked-row values)."""
    lens = list(int(x) for x in seq_lens)
    if rate <= 0:
        return lens
    while True:
        n = supertext_length(lens, k)
        q = n // rate
        if n % rate != 0 and (q & (q - 1)) != 0 and q > 1:
            return lens
        lens[-1] += k


def random_sequences(seq_lens, seed: int):
    """One uint8 ASCII array per sequence, i.i.d. uniform over ACGT."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [ACGT[rng.integers(0, 4, size=int(L), dtype=np.uint8)] for L in seq_lens]


def write_fasta(path: str, seqs, names=None) -> None:
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            name = names[i] if names else f"seq{i}"
            f.write(b">" + name.encode() + b"\n")
            f.write(s.tobytes())
            f.write(b"\n")


def sample_patterns(seqs, n_patterns: int, length: int, seed: int, random_frac: float = 0.05):
    """`n_patterns` substrings of `length` starting at uniform positions inside sequences (never
    spanning a boundary) plus `random_frac` uniformly random strings (mostly non-occurring).
    Returns a uint8 array of shape (n_total, length)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    ok = lens >= length
    if not ok.any():
        raise ValueError("no sequence is long enough for the requested pattern length")
    weights = np.where(ok, lens - length + 1, 0).astype(np.float64)
    weights /= weights.sum()
    which = rng.choice(len(seqs), size=n_patterns, p=weights)
    out = np.empty((n_patterns, length), dtype=np.uint8)
    idx = np.arange(length, dtype=np.int64)
    for si in np.unique(which):
        sel = np.nonzero(which == si)[0]
        starts = rng.integers(0, lens[si] - length + 1, size=len(sel))
        out[sel] = seqs[si][starts[:, None] + idx[None, :]]
    n_rand = int(round(n_patterns * random_frac))
    if n_rand:
        rnd = ACGT[rng.integers(0, 4, size=(n_rand, length), dtype=np.uint8)]
        out = np.concatenate([out, rnd], axis=0)
    return out


def write_patterns(path: str, pats: np.ndarray) -> None:
    """One pattern per line (the reference CLI's patterns-file format, Main.cpp:1055)."""
    n, L = pats.shape
    buf = np.empty((n, L + 1), dtype=np.uint8)
    buf[:, :L] = pats
    buf[:, L] = ord("\n")
    with open(path, "wb") as f:
        f.write(buf.tobytes())


def pack_patterns(pats: np.ndarray):
    """Flat byte buffer + uint32 offsets (n+1) — the batch layout the C ABI takes."""
    n, L = pats.shape
    offs = (np.arange(n + 1, dtype=np.uint64) * L).astype(np.uint32 if n * L < 2**32 else np.uint64)
    return np.ascontiguousarray(pats).reshape(-1), offs


def naive_find_all(text: bytes, pat: bytes):
    """Ground truth: every start of `pat` in `text` (overlaps included) — the authors' own
    intended check, Main.cpp:1502-1510 (naiveSearch)."""
    out = []
    i = text.find(pat)
    while i != -1:
        out.append(i)
        i = text.find(pat, i + 1)
    return out


def collection_text(seqs, k: int) -> bytes:
    """The concatenated, '&'-padded text the reference indexes (EFMCollection.cpp:187-233)."""
    parts = []
    for s in seqs:
        b = s.tobytes()
        pad = (-len(b)) % k
        parts.append(b + b"&" * pad + b"&" * k)
    return b"".join(parts)
