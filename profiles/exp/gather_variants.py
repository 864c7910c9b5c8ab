"""What does ONE random gather cost in DRAM traffic, and does the load instruction change it?
Runs e2fm_debug_gather_bench_ex for every load variant over a 16 GB buffer (and the L2 fetch-granularity limit at 32 / 128);
run it plain for the rates and under `ncu --metrics dram__bytes_read.sum,...` for the bytes per gather."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from e2fm_b200 import capi  # noqa: E402

NAMES = {0: "ld.global.nc.u64", 1: "ld.global.u64", 2: "ld.global.cg.u64", 3: "ld.global.nc.v8.u32 (32 B)", 4: "2 x ld.global.nc.v4.u32 (one sector)",
         5: "ld.global.nc.L2::64B.u64", 6: "ld.global.nc.L1::no_allocate.u64", 7: "ld.global.nc.L2::128B.u64",
         8: "ld.global.nc.L2::64B.v8.u32 (32 B, ld_sector)"}
capi.lib()
rt = C.CDLL("libcudart.so.12")
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = 1 << 27
out = {}
for gran in (None, 32):
    if gran:
        rt.cudaDeviceSetLimit(0x05, C.c_size_t(gran))
    for v in range(9):
        g = capi.gather_bench(mb << 20, n, 0, v)
        out[f"{NAMES[v]} | limit {gran or 'default'}"] = round(g, 2)
        print(f"variant {v} {NAMES[v]:40s} limit {gran or 'default'}: {g:7.2f} G gathers/s", flush=True)
print(json.dumps(out))
