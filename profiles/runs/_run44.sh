#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t24.log 2>&1; echo "tests rc=$?"; grep -n "^E  " gpurun_out/t24.log | head -8 | cut -c1-400; tail -2 gpurun_out/t24.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp > gpurun_out/bench_mid50_1gpu_r2.json 2> gpurun_out/bench_mid50_1gpu_r2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_mid50_1gpu_r2.json")); print(d["value"]/1e9, d["e2e"]["value"]/1e9, d["gpu_launches"], d.get("gpu_launches_by_leg"), d["parity"], d["e2e"]["host_link"])
P
export E2FM_TRACE_MULT=2 E2FM_TRACE_TIGHT=1 E2FM_HITS_TRACE=0
timeout 600 python profiles/exp/hits_trace.py mid50 2>&1 | grep wall | tr '\n' ' '; echo " (default pinned)"
E2FM_TRACE_WC=1 timeout 600 python profiles/exp/hits_trace.py mid50 2>&1 | grep -E "wall|Error|error" | tr '\n' ' '; echo " (write-combined)"
timeout 600 python profiles/exp/hits_trace.py mid50 2>&1 | grep wall | tr '\n' ' '; echo " (default pinned again)"
