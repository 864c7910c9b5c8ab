#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload ecoli --steps 10 --warmup 3 --no-cpp > gpurun_out/bench_ecoli_r2.json 2> gpurun_out/bench_ecoli_r2.err; echo "rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_ecoli_r2.json")); print(d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"]["device_timeline_ms_per_step"], d["e2e"]["what"][-60:], d["parity"], d["roofline"]["bound"], d["roofline"]["frac"], d["phase_ms_per_step"])
P
export E2FM_TRACE_TIGHT=1 E2FM_HITS_TRACE=1 E2FM_TRACE_REPS=4
timeout 300 python profiles/exp/hits_trace.py ecoli 2>&1 | tail -8
