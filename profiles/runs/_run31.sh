#!/bin/bash
# tight packing: parity of e2fm_search_hits in the three input forms, then the end-to-end effect on a mid-sized workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search_hits" > gpurun_out/t19.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t19.log
timeout 600 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline --no-cpp > gpurun_out/m14_tight.json 2> gpurun_out/m14_tight.err; echo "tight rc=$?"
timeout 600 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline --no-cpp --no-tight > gpurun_out/m14_words.json 2> gpurun_out/m14_words.err; echo "words rc=$?"
python - <<'P'
import json
for f in ("tight","words"):
    try:
        d=json.load(open(f"gpurun_out/m14_{f}.json")); print(f, d["value"]/1e9, d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["parity"])
    except Exception as e: print(f, "failed", e)
P
