"""Does the L2 fetch granularity limit change what HBM delivers for random 32-byte-sector gathers?
Runs the library's gather microbenchmark (e2fm_debug_gather_bench) at cudaLimitMaxL2FetchGranularity = 128 (default), 64, 32."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from e2fm_b200 import capi  # noqa: E402

capi.lib()
rt = C.CDLL("libcudart.so.12")
LIMIT = 0x05   # cudaLimitMaxL2FetchGranularity
out = {}
for gran in (128, 64, 32):
    rc = rt.cudaDeviceSetLimit(LIMIT, C.c_size_t(gran))
    got = C.c_size_t(0)
    rt.cudaDeviceGetLimit(C.byref(got), LIMIT)
    row = {"set_rc": int(rc), "limit_now": int(got.value)}
    for mb in (64, 1024, 16384):
        row[f"{mb}MB"] = round(capi.gather_bench(mb << 20, 1 << 28, 0), 2)
    out[str(gran)] = row
    print(gran, row, flush=True)
print(json.dumps(out))
