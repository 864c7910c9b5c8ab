"""Time of seed_lookup_kernel / seed_verify*_kernel on a device-resident batch, normally and with E2FM_SEED_DEBUG=1 (every lookup
reads one of 64 table entries: the kernel's instruction work without its DRAM gathers). Usage: python profiles/exp/lookup_probe.py [workload]"""
import os
import sys

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
packed, _ = capi.pack_patterns(pats)
d = torch.from_numpy(packed).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dbg in ("0", "1", "0"):
    os.environ["E2FM_SEED_DEBUG"] = dbg
    best = {}
    for _ in range(5):
        flush.add_(1)
        torch.cuda.synchronize()
        r = ix.search_packed_device(d.data_ptr(), n, L, True)
        ms = ix.last_batch_ms()
        r.free()
        for k in ("backward_search", "walk", "results", "total"):
            best[k] = min(best.get(k, 1e9), ms[k])
    print(f"E2FM_SEED_DEBUG={dbg}: lookup {best['backward_search']:.3f} ms, verify {best['walk']:.3f} ms, results {best['results']:.3f} ms, "
          f"total {best['total']:.3f} ms for {n} patterns", flush=True)
ix.close()
