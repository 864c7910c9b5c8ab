#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --workload chr1 --steps 5 --warmup 3 --no-cpp > gpurun_out/bench_chr1_r2b.json 2> gpurun_out/bench_chr1_r2b.err; echo "rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_chr1_r2b.json")); print(d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["parity"], (d.get("cpu_baseline") or {}).get("value"), d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
P
