#!/bin/bash
# configs[2] (100 genomes, 10 M patterns of length 24) on one GPU with the parity block; ncu of the widen kernel
mkdir -p gpurun_out
timeout 1200 python bench.py --workload collection100 --steps 5 --warmup 3 > gpurun_out/bench_c100_r2.json 2> gpurun_out/bench_c100_r2.err; echo "c100 rc=$?"
tail -4 gpurun_out/bench_c100_r2.err
python - <<'P'
import json
try:
    d=json.load(open("gpurun_out/bench_c100_r2.json")); print(d["value"]/1e9, d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["parity"], (d.get("cpu_baseline") or {}).get("value"), d["roofline"]["frac"], d["roofline"]["kernel"])
except Exception as e: print("failed", e)
P
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:"widen_packed" -s 8 -c 1 -o gpurun_out/prof_mid50_widen_r2 python bench.py --workload mid50 --steps 1 --warmup 3 --no-cpu-baseline --no-cpp > gpurun_out/n6.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_mid50_widen_r2.ncu-rep --page raw --csv > gpurun_out/prof_mid50_widen_r2.raw.csv 2>/dev/null
ls -la gpurun_out/prof_mid50_widen_r2*
