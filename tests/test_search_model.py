"""tests/search_model.py (a numpy model of the device search algorithm) against the oracle: the model of what the kernels
do today, and the two gather-saving plans of DESIGN.md section 7b, must all return the oracle's counts and positions."""
import numpy as np
import pytest

from conftest import GOLDEN_SEARCH_CASES
import search_model as sm


def _extra_patterns(img, rng, n_each=40):
    """substrings of the indexed text at every alignment and many lengths (incl. the short ones that lose shifts, and ones
    that run over padding / terminators), plus strings that occur nowhere"""
    plain = b"".join(img.kmer[int(c)] for c in img.text[: img.n - 1])
    out = []
    for length in (1, 2, 3, 4, 5, 7, 8, 9, 11, 12, 13, 16, 17, 21, 24, 31, 32, 40):
        for _ in range(n_each if length > 4 else 8):
            a = int(rng.integers(0, max(1, len(plain) - length)))
            out.append(plain[a:a + length])
    syms = sorted(set(plain))
    for length in (6, 12, 20):
        for _ in range(10):
            out.append(bytes(rng.choice(syms, size=length).tolist()))
    out.append(b"")
    out.append(b"ACGTN#")          # foreign symbols
    return out


@pytest.mark.parametrize("name", GOLDEN_SEARCH_CASES)
def test_model_and_plans_match_oracle(oracle_mod, golden, key64, name):
    ix = oracle_mod.OracleIndex(golden[name]["index"], key64)
    img = sm.DeviceImage.from_oracle(ix)
    img.build_seeds()
    rng = np.random.default_rng(2026)
    pats = list(golden[name]["patterns"]) + _extra_patterns(img, rng)
    n_hits = n_seed_rejects = 0
    for pat in pats:
        want_count, want_pos = ix.search(pat, locate=True)
        n_hits += want_count > 0
        for plan in ("layout", "verify", "seed"):
            got_count, got_pos = sm.search(img, pat, True, plan)
            assert got_count == want_count and np.array_equal(got_pos, want_pos), (name, plan, pat)
            c_only, _ = sm.search(img, pat, False, plan)
            assert c_only == want_count, (name, plan, pat, "count mode")
        # a shift the seed filter rejects must be one the plain search finds nothing for
        for sh in range(img.k):
            o = sm.search_shift(img, pat, sh, True, "seed", sm.Gathers())
            if o.rejected_by_seed:
                n_seed_rejects += 1
                assert sm.search_shift(img, pat, sh, True, "layout", sm.Gathers()).count == 0
    assert n_hits > len(pats) // 2
    ix.close()


def test_plans_gather_fewer_sectors(oracle_mod, golden, key64):
    """the point of the plans: fewer HBM sectors than the layout's step-by-step search on patterns that occur"""
    ix = oracle_mod.OracleIndex(golden["col3_k4"]["index"], key64)
    img = sm.DeviceImage.from_oracle(ix)
    img.build_seeds()
    plain = b"".join(img.kmer[int(c)] for c in img.text[: img.n - 1])
    rng = np.random.default_rng(7)
    tot = {p: sm.Gathers() for p in ("layout", "verify", "seed")}
    for _ in range(200):
        a = int(rng.integers(0, len(plain) - 40))
        for plan, g in tot.items():
            sm.search(img, plain[a:a + 40], True, plan, g)
    assert tot["verify"].hbm() < tot["layout"].hbm()
    assert tot["seed"].hbm() <= tot["verify"].hbm() + tot["seed"].seed
    ix.close()
