"""A numpy model of the DEVICE search algorithm, checked against the oracle on the CPU.

TEST INFRASTRUCTURE (like oracle/): it is not shipped and nothing in e2fm_b200/ imports it. It restates, per
(pattern, shift), what the kernels of e2fm_b200/csrc/search.cu do over the flattened layout of
csrc/device_index.cuh — dense sa[] / text[] tables, one rank sector per step — so that

  * the algorithm the kernels implement has a readable, CPU-checkable specification (`search`, plan "layout"), and
  * changes to that algorithm can be proven bit-exact BEFORE any GPU time is spent on them: plan "verify" is the
    suffix-verification shortcut and plan "seed" adds the alignment seed filter of DESIGN.md section 7b. Both must
    return exactly what plan "layout" returns; the model also counts the 32-byte sectors each plan gathers.

Reference counterparts: computeSuperPatterns (EFMIndex.cpp:2055-2207), start interval (:1728-1748), backwardStep
(:1966-2001), backwardStepIfMatch (:2011-2041), findWholePatternMatches (:1862-1899), getRowPosition (:2468-2497),
locateOccurrences (:2281-2323).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Gathers:
    """32-byte sectors gathered, by array (the pair table is L2-resident and listed apart)"""
    pair: int = 0
    occ: int = 0
    rec: int = 0
    sa: int = 0
    text: int = 0
    seed: int = 0

    def hbm(self) -> int:
        return self.occ + self.rec + self.sa + self.text + self.seed

    def add(self, o: "Gathers") -> None:
        for f in ("pair", "occ", "rec", "sa", "text", "seed"):
            setattr(self, f, getattr(self, f) + getattr(o, f))


@dataclass
class DeviceImage:
    """what the load pipeline leaves in HBM, rebuilt here from the oracle's per-row views"""
    n: int
    k: int
    rate: int
    initial_n: int
    A: int
    C: np.ndarray                 # C[c] for compact codes, A+1 entries
    bwt: np.ndarray               # compact code per row
    sa: np.ndarray                # super-text position per row            (dense_build_kernel)
    text: np.ndarray              # compact code per super-text position   (dense_build_kernel)
    kmer: list                    # compact code -> k symbols (bytes), None when the code never shows in [0, n-2]
    code_of: dict                 # k symbols -> compact code
    rows_of: list = field(default_factory=list)   # per code: ascending rows whose BWT symbol it is
    blk_shift: int = 11
    seeds: set | None = None      # every SEED_LEN-gram of text[], cyclic

    SEED_LEN = 4

    @classmethod
    def from_oracle(cls, ix) -> "DeviceImage":
        inf = ix.info
        n, k, rate = int(inf.n), int(inf.k), int(inf.rate)
        A = int(inf.compact_alphabet)
        bwt = ix.bwt(0, n).astype(np.int64)
        lf = ix.lf_rows(np.arange(n, dtype=np.uint64)).astype(np.int64)
        marks = ix.marks(0, n)
        marked = marks != np.uint64(0xFFFFFFFFFFFFFFFF)
        # EFMIndex::getRowPosition for every row: walk LF to the next marked row, position = (value * rate + steps) mod n
        sa = np.full(n, -1, dtype=np.int64)
        cur = np.arange(n, dtype=np.int64)
        todo = np.arange(n, dtype=np.int64)
        for t in range(rate + 2):
            hit = marked[cur]
            sa[todo[hit]] = (marks[cur[hit]].astype(np.int64) * rate + t) % n
            todo, cur = todo[~hit], lf[cur[~hit]]
            if len(todo) == 0:
                break
        assert len(todo) == 0, "a row did not reach a marked row within `rate` steps"
        assert len(np.unique(sa)) == n
        text = np.empty(n, dtype=np.int64)
        text[(sa - 1) % n] = bwt
        # k-mer of every compact code, read off the extracted text: super-character p covers symbols [p*k, (p+1)*k)
        first = int(inf.initial_n)
        plain = ix.extract_snippet(first, first + (n - 1) * k - 1)
        assert len(plain) == (n - 1) * k
        kmer = [None] * A
        code_of = {}
        for p in range(n - 1):
            c = int(text[p])
            if kmer[c] is None:
                kmer[c] = plain[p * k:(p + 1) * k]
                code_of[kmer[c]] = c
        Cc = ix.header_C().astype(np.int64)
        order = np.argsort(bwt, kind="stable")
        bounds = np.searchsorted(bwt[order], np.arange(A + 1))
        rows_of = [order[bounds[c]:bounds[c + 1]] for c in range(A)]
        blk = min(16384, max(64, 1 << ((8 * A).bit_length() - 1)))       # api.cu: rank block = 8*A rounded down to 2^j
        img = cls(n, k, rate, first, A, Cc, bwt, sa, text, kmer, code_of, rows_of, blk.bit_length() - 1)
        return img

    def build_seeds(self) -> None:
        L = self.SEED_LEN
        t = self.text
        self.seeds = set(zip(*[np.roll(t, -j).tolist() for j in range(L)]))

    # C[c] + occ(c, row): rows <= `row` whose BWT symbol is c (sector_rank over one 32-byte sector on the device)
    def rank(self, c: int, row: int) -> int:
        return int(self.C[c]) + int(np.searchsorted(self.rows_of[c], row, side="right"))


@dataclass
class ShiftOutcome:
    count: int = 0
    positions: list = field(default_factory=list)
    rejected_by_seed: bool = False


def _final_row(img: DeviceImage, pat: bytes, sa_shift: int, L: int, hat: bool, locate: bool, g: Gathers,
               bwt_code: int, pos_of_row: int, out: ShiftOutcome, fresh_gathers: bool) -> None:
    """resolve_kernel for one row task. `bwt_code` is the row's BWT symbol, `pos_of_row` its sa[] value.
    fresh_gathers: the plan reads rec / sa / text for this row (False when a text run already covers them)."""
    k, n, m = img.k, img.n, len(pat)
    if sa_shift > 0:
        # backwardStepIfMatch: the '?'-prefixed first super-character fixes its LAST f symbols
        f = k - sa_shift
        if fresh_gathers:
            g.rec += 1
        km = img.kmer[bwt_code]
        if km is None or km[k - f:] != pat[:f]:
            return
    need_tp = hat or locate
    tp = pos_of_row
    if need_tp and fresh_gathers:
        g.sa += 1
    if sa_shift > 0:
        tp = n - 1 if tp == 0 else tp - 1          # the match stands for the suffix one super-character earlier
    if hat:
        # findWholePatternMatches / extractSuperCharacter: the '?'-suffixed last super-character fixes its FIRST f symbols
        frm = (L - 1) * k - sa_shift
        f = m - frm
        p = tp + L - 1
        if p >= n:
            return
        if fresh_gathers:
            g.text += 1
        km = img.kmer[int(img.text[p])]
        if km is None or km[:f] != pat[frm:]:
            return
    out.count += 1
    if locate:
        out.positions.append(tp * k + sa_shift + img.initial_n)


def search_shift(img: DeviceImage, pat: bytes, sa_shift: int, locate: bool, plan: str, g: Gathers) -> ShiftOutcome:
    """one (pattern, shift) item: backward_search_kernel + firstcheck/resolve. plan: 'layout' | 'verify' | 'seed'"""
    out = ShiftOutcome()
    k, n, m = img.k, img.n, len(pat)
    if m == 0:
        return out
    L = (m + sa_shift + k - 1) // k
    last_var = (m + sa_shift) % k != 0
    l = -1
    if not last_var:
        if not (L == 1 and sa_shift > 0):
            l = L
    elif L > 1 and (L > 2 or sa_shift == 0):
        l = L - 1
    if l <= 0:
        return out                                  # quirk 1: the shift silently yields nothing
    lo = 1 if sa_shift > 0 else 0

    def code(c):                                    # compact code of super-character c of this shift (all k symbols fixed)
        return img.code_of.get(pat[c * k - sa_shift:(c + 1) * k - sa_shift])

    codes = {c: code(c) for c in range(lo, l)}
    if codes[l - 1] is None or (l - 2 >= lo and codes[l - 2] is None):
        return out
    if plan == "seed" and l - lo >= img.SEED_LEN:
        g.seed += 1
        gram = tuple(codes[c] for c in range(l - img.SEED_LEN, l))
        if None in gram or gram not in img.seeds:
            out.rejected_by_seed = True             # (the tests check that the plain search comes back empty too)
            return out
    cc = codes[l - 1]
    s, e = int(img.C[cc]), int(img.C[cc + 1])       # start interval [s, e)
    i = l - 2
    if i >= lo:                                     # pair table: start interval + first step in one (L2-resident) gather
        c1 = codes[i]
        ns = int(img.C[c1]) if s == 0 else img.rank(c1, s - 1)
        s, e = ns, img.rank(c1, e - 1)
        g.pair += 1
        i -= 1
    known_pos = None                                # plan verify: super-text position of the final row, when it was derived
    while e > s and i >= lo:
        if plan in ("verify", "seed") and e - s == 1 and i - lo + 1 >= 2:
            # suffix verification: the remaining super-characters must precede the row's suffix in the text
            p = int(img.sa[s])
            g.sa += 1
            need = i - lo + 1
            span_lo = p - need - (1 if sa_shift > 0 else 0)          # ... plus the wildcard character before them
            span_hi = p - 1
            g.text += (span_hi // 16) - (span_lo // 16) + 1           # 16 u16 codes per 32-byte sector
            ok = all(codes[i - j] is not None and int(img.text[(p - 1 - j) % n]) == codes[i - j] for j in range(need))
            if not ok:
                s = e = 0
                break
            known_pos = (p - need) % n
            i = lo - 1
            break
        cc = codes[i]
        if cc is None:
            s = e = 0
            break
        re_ = e - 1
        if s == 0:
            ns = int(img.C[cc])
            g.occ += 1
        else:
            rs = s - 1
            g.occ += 1 if (rs >> img.blk_shift) == (re_ >> img.blk_shift) else 2
            ns = img.rank(cc, rs)
        s, e = ns, img.rank(cc, re_)
        i -= 1
    if e <= s and known_pos is None:
        return out
    first_var, hat = sa_shift > 0, last_var
    if known_pos is not None:
        # one row whose position is known: its BWT symbol is the text symbol before it; the hat symbol lies L-1 further on
        bw = int(img.text[(known_pos - 1) % n])
        before = Gathers()
        _final_row(img, pat, sa_shift, L, hat, locate, before, bw, known_pos, out, fresh_gathers=False)
        if hat:
            g.text += 1                             # the hat symbol is ~L positions ahead: one more sector
        return out
    if not locate and not first_var and not hat:
        out.count = e - s                           # countOccurrences: the interval length
        return out
    for row in range(s, e):
        _final_row(img, pat, sa_shift, L, hat, locate, g, int(img.bwt[row]), int(img.sa[row]), out, fresh_gathers=True)
    return out


def search(img: DeviceImage, pat: bytes, locate: bool = True, plan: str = "layout", g: Gathers | None = None):
    """-> (count, sorted de-duplicated positions) exactly as EFMIndex::countOccurrences / locateOccurrences report them"""
    g = g if g is not None else Gathers()
    count, pos = 0, []
    for sa_shift in range(img.k):
        o = search_shift(img, pat, sa_shift, locate, plan, g)
        count += o.count
        pos.extend(o.positions)
    return count, np.unique(np.asarray(pos, dtype=np.uint64))


def main():
    """gather statistics on a mid-size index: python tests/search_model.py [Mbp] [patterns] [pattern length]"""
    import os
    import sys
    import tempfile
    import time

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from e2fm_b200 import synth
    from oracle import oracle

    mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    npat = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    plen = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    k, rate = 4, 50
    with tempfile.TemporaryDirectory() as td:
        lens = synth.safe_lengths([int(mbp * 1e6)], k, rate)
        seqs = synth.random_sequences(lens, 4242)
        fa, efm, keyf = os.path.join(td, "t.fa"), os.path.join(td, "t.efm"), os.path.join(td, "key.bin")
        synth.write_fasta(fa, seqs)
        with open(keyf, "wb") as f:
            f.write(synth.test_key())
        oracle.ref_build(fa, efm, keyf, k=k, bucket_size=4096, mrp=2)
        ix = oracle.OracleIndex.from_file(efm, synth.test_key())
        t0 = time.time()
        img = DeviceImage.from_oracle(ix)
        img.build_seeds()
        print(f"n = {img.n} rows, A = {img.A}, rank block {1 << img.blk_shift} rows, image in {time.time() - t0:.1f}s; "
              f"{len(img.seeds)} distinct {img.SEED_LEN}-grams = {len(img.seeds) / img.A ** img.SEED_LEN:.4%} of A^{img.SEED_LEN}")
        pats = synth.sample_patterns(seqs, npat, plen, 777)
        for locate in (False, True):
            tot = {p: Gathers() for p in ("layout", "verify", "seed")}
            for row in pats:
                pat = row.tobytes()
                want = search(img, pat, locate, "layout", tot["layout"])
                for plan in ("verify", "seed"):
                    got = search(img, pat, locate, plan, tot[plan])
                    assert got[0] == want[0] and np.array_equal(got[1], want[1]), (plan, pat)
            print(f"{'count+locate' if locate else 'count'}: {npat} patterns of length {plen} — 32-byte sectors per pattern")
            for plan, gg in tot.items():
                print(f"  {plan:7s} hbm {gg.hbm() / npat:6.2f}  (occ {gg.occ / npat:5.2f}  rec {gg.rec / npat:5.2f}  sa {gg.sa / npat:5.2f}  "
                      f"text {gg.text / npat:5.2f}  seed {gg.seed / npat:5.2f})   pair table {gg.pair / npat:5.2f}")


if __name__ == "__main__":
    main()
