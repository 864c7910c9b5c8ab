#!/bin/bash
# configs[2] (100 genomes) on 2 GPUs: replicas + pattern partition, and the collection sharded by sequence (50 + 50) with its answers compared
mkdir -p gpurun_out
timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload collection100 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c100_2gpu_r2.json 2> gpurun_out/bench_c100_2gpu_r2.err; echo "rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_c100_2gpu_r2.err | tail -5 | cut -c1-300
python - <<'P'
import json
try:
    d=json.load(open("gpurun_out/bench_c100_2gpu_r2.json")); sh=d.get("sharded") or {}
    print(d["value"]/1e9, d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("host_link"), d["parity"], sh.get("value"), sh.get("same_answers_as_replicas"), sh.get("phase_ms_rank0"), sh.get("error"))
except Exception as e: print("failed", e)
P
