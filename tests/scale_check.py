"""Full-size parity check for a bench workload (run by hand on the GPU box after `bench.py --workload W` has cached the index):

    python tests/scale_check.py human

Size-independent properties at BASELINE.json's full sizes: every pattern cut from the text is located at its own
position (collection-wide position AND (sequence, offset)), positions ascending per pattern, count == number of positions,
count-only == locate counts, dense tables == per-query walks; plus a sample against the reference's CPU search when the
reference can open the index."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from e2fm_b200 import capi, synth  # noqa: E402
from e2fm_b200.sharded import padded_length  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "collection25"
    npat = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    prep = bench.prepare(wl, 0)
    w = prep["w"]
    L = w["pat_len"]
    seqs = prep["seqs"]
    with open(prep["efm"], "rb") as f:
        raw = f.read()
    t0 = time.time()
    dev = capi.DeviceIndex(raw, synth.test_key(), 0)
    print(f"[scale] {wl}: n={dev.info.text_length} sequences={dev.info.n_sequences} device={dev.info.device_bytes / 1e9:.2f} GB "
          f"load {time.time() - t0:.2f}s dense={dev.info.dense}", flush=True)
    rng = np.random.default_rng(123)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    which = rng.choice(len(seqs), size=npat, p=(lens - L + 1) / (lens - L + 1).sum())
    starts = np.empty(npat, dtype=np.int64)
    pats = np.empty((npat, L), dtype=np.uint8)
    idx = np.arange(L)
    for si in np.unique(which):
        sel = np.nonzero(which == si)[0]
        st = rng.integers(0, lens[si] - L + 1, size=len(sel))
        starts[sel] = st
        pats[sel] = seqs[si][st[:, None] + idx[None, :]]
    base = np.concatenate([[0], np.cumsum([padded_length(int(x), bench.K) for x in lens])])[:-1]
    gpos = (base[which] + starts).astype(np.uint64)
    r = dev.search_fixed(pats, True)
    out = r.fetch()
    offs = out["pos_offsets"].astype(np.int64)
    cnt = np.diff(offs)
    assert np.array_equal(out["counts"].astype(np.int64), cnt)
    assert (cnt >= 1).all()
    pos = out["positions"]
    seg = np.repeat(np.arange(npat), cnt)
    d = np.diff(pos.astype(np.int64))
    assert (d[np.diff(seg) == 0] > 0).all(), "positions not ascending inside a pattern"
    hit = np.zeros(npat, dtype=bool)
    m = pos == gpos[seg]
    hit[seg[m]] = True
    assert hit.all(), "a pattern was not located at its sampling position"
    assert np.array_equal(out["seq"][m].astype(np.int64), which[seg[m]]) and np.array_equal(out["off"][m].astype(np.int64), starts[seg[m]])
    r2 = dev.search_fixed(pats, False)
    assert np.array_equal(r2.fetch()["counts"], out["counts"])
    dev.set_option(capi.OPT_DENSE, 0)
    sub = pats[:200_000]
    r3 = dev.search_fixed(np.ascontiguousarray(sub), True)
    o3 = r3.fetch()
    assert np.array_equal(o3["positions"], pos[:offs[200_000]]) and np.array_equal(o3["counts"], out["counts"][:200_000])
    dev.set_option(capi.OPT_DENSE, 1)
    print(f"[scale] properties ok: {npat} patterns, {len(pos)} occurrences; walks == dense tables on 200k", flush=True)
    # the reference's CPU search on a sample
    status = "reference not built"
    if oracle.have_ref():
        with tempfile.TemporaryDirectory() as td:
            pf = os.path.join(td, "p.txt")
            k = 2000
            synth.write_patterns(pf, pats[:k])
            try:
                res, tj = oracle.ref_search(prep["efm"], prep["keyfile"], pf, os.path.join(td, "r.res"), True, 2)
                a = int(offs[k])
                assert np.array_equal(res["counts"], out["counts"][:k])
                assert np.array_equal(res["positions"], pos[:a]) and np.array_equal(res["seq"], out["seq"][:a]) and np.array_equal(res["off"], out["off"][:a])
                status = f"bit-exact with the reference's CPU search on {k} patterns ({tj['pass_ms'][0]:.0f} ms on the CPU)"
            except RuntimeError as e:
                status = "reference could not search this index: " + str(e)[-200:]
    print("[scale] " + status, flush=True)
    print(json.dumps({"workload": wl, "patterns": npat, "occurrences": int(len(pos)), "reference_sample": status}))
    dev.close()


if __name__ == "__main__":
    main()
