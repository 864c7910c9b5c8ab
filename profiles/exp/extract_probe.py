"""One large e2fm_extract call on a bench workload's index (for an ncu capture of extract_kernel). Usage: python profiles/exp/extract_probe.py [workload]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from e2fm_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "mid50"
prep = bench.prepare(wl, 0)
ix = capi.DeviceIndex(open(prep["efm"], "rb").read(), synth.test_key(), 0)
for _ in range(3):
    s = ix.extract_snippet(1000, 1000 + (64 << 20))
print(len(s), s[:40])
ix.close()
