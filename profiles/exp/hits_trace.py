"""Per-stage timeline of e2fm_search_hits (E2FM_HITS_TRACE=1) on a bench workload. Usage: python profiles/exp/hits_trace.py [workload] [stage]"""
import os
import sys
import time

import numpy as np

os.environ.setdefault("E2FM_HITS_TRACE", "1")     # 2: an event after every kernel group; E2FM_HITS_SERIAL=1: search after the whole H2D
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from e2fm_b200 import capi, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "mid50"
prep = bench.prepare(wl, 0)
w = prep["w"]
n, L = w["n_patterns"], w["pat_len"]
ix = capi.DeviceIndex(open(prep["efm"], "rb").read(), synth.test_key(), 0)
if len(sys.argv) > 2:
    ix.set_option(capi.OPT_HITS_STAGE, int(sys.argv[2]))
pats = np.ascontiguousarray(bench.load_patterns(prep)[:n])
if os.environ.get("E2FM_TRACE_N"):                           # a shorter batch (a GPU's share of a partitioned one)
    n = min(n, int(os.environ["E2FM_TRACE_N"]))
    pats = np.ascontiguousarray(pats[:n])
mult = int(os.environ.get("E2FM_TRACE_MULT", "1"))          # the batch repeated: more pipeline stages
if mult > 1:
    pats = np.ascontiguousarray(np.tile(pats, (mult, 1)))
    n *= mult
tight = os.environ.get("E2FM_TRACE_TIGHT", "0") != "0"    # ceil(L / 4) bytes per pattern instead of 16-byte words
stride = capi.packed_tight_stride(L) if tight else capi.packed_stride(L)
if os.environ.get("E2FM_TRACE_WC", "0") != "0":            # experiment: write-combined page-locked memory for the input
    import ctypes

    from cuda import cudart
    err, ptr = cudart.cudaHostAlloc(n * stride, cudart.cudaHostAllocWriteCombined)
    assert int(err) == 0, err

    class _WC:
        array = np.ctypeslib.as_array((ctypes.c_uint8 * (n * stride)).from_address(int(ptr))).reshape(n, stride)
    h_in = _WC()
else:
    h_in = capi.PinnedArray((n, stride), np.uint8)
(capi.pack_patterns_tight if tight else capi.pack_patterns)(pats, h_in.array.reshape(-1))
cap = 2 * n + 4096
h_off, h_seq, h_so = (capi.PinnedArray(s, np.uint32) for s in (n + 1, cap, cap))
for rep in range(int(os.environ.get("E2FM_TRACE_REPS", "4"))):
    print(f"--- call {rep}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    ix.search_hits(h_in.array.reshape(-1), n, L, 2 if tight else 1, True, None, h_off.array, h_seq.array, h_so.array)
    print(f"--- wall {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr, flush=True)
ix.close()
