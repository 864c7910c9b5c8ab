"""Where does the end-to-end time of e2fm_search_hits go? Stage-size sweep, count-only against locate, and the raw PCIe copy times
of the same byte counts (one direction, both directions at once). Usage: python profiles/exp/e2e_stages.py [workload]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from e2fm_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "mid50"
prep = bench.prepare(wl, 0)
w = prep["w"]
n, L = w["n_patterns"], w["pat_len"]
ix = capi.DeviceIndex(open(prep["efm"], "rb").read(), synth.test_key(), 0)
pats = np.ascontiguousarray(bench.load_patterns(prep)[:n])
stride = capi.packed_stride(L)
h_in = capi.PinnedArray((n, stride), np.uint8)
capi.pack_patterns(pats, h_in.array.reshape(-1))
cap = 2 * n + 4096
h_cnt, h_off, h_seq, h_so = (capi.PinnedArray(s, np.uint32) for s in (n, n + 1, cap, cap))


def run(locate, reps=5):
    best = 1e9
    for _ in range(reps + 2):
        t0 = time.perf_counter()
        ix.search_hits(h_in.array.reshape(-1), n, L, True, locate, None if locate else h_cnt.array, h_off.array if locate else None,
                       h_seq.array if locate else None, h_so.array if locate else None)
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, ix.last_batch_counters()["hits_stages"], ix.last_batch_ms()["total"]


for stage in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 24):
    ix.set_option(capi.OPT_HITS_STAGE, stage)
    for locate in (True, False):
        ms, st, dev = run(locate)
        print(f"stage {stage:9d} {'locate' if locate else 'count '}: {ms:7.3f} ms wall, {st:3d} stages, device timeline {dev:7.3f} ms, "
              f"{n / ms / 1e6:7.1f} G patterns/s", flush=True)
# raw PCIe
nb_in, nb_out = n * stride, (n + 1) * 4 + 2 * 4 * n
d_in = torch.empty(nb_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(nb_out, dtype=torch.uint8, device="cuda")
t_in = torch.from_numpy(h_in.array.reshape(-1))
h_o = capi.PinnedArray(nb_out, np.uint8)
t_out = torch.from_numpy(h_o.array)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(f, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        f()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(t_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        t_out.copy_(d_out, non_blocking=True)


a, b = timed(h2d), timed(d2h)
c = timed(lambda: (h2d(), d2h()))
print(f"raw PCIe: H2D {nb_in / 1e6:.0f} MB {a:.3f} ms ({nb_in / a / 1e6:.1f} GB/s), D2H {nb_out / 1e6:.0f} MB {b:.3f} ms ({nb_out / b / 1e6:.1f} GB/s), both at once {c:.3f} ms")
ix.close()
