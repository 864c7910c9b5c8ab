#!/bin/bash
# full GPU suite, then the default bench command (human, configs[3]) as the driver runs it, then the launch list of the same command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t23.log 2>&1; echo "tests rc=$?"; grep -n "^E  " gpurun_out/t23.log | head -8 | cut -c1-400; tail -3 gpurun_out/t23.log
timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_human_r2c.json 2> gpurun_out/bench_human_r2c.err; echo "bench rc=$?"
tail -4 gpurun_out/bench_human_r2c.err
python - <<'P'
import json
d=json.load(open("gpurun_out/bench_human_r2c.json")); print(d["value"]/1e9, d["e2e"], d["parity"], d["config"].get("cpu_affinity_rank0"), d["gpu_launches"])
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_human_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cpp > gpurun_out/n0.log 2>&1; echo "launch list rc=$?"
