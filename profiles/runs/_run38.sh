#!/bin/bash
# the driver's N=2 command on the default (human, configs[3]) workload: replicas + pattern partition, sharded leg, reference arm
mkdir -p gpurun_out
nvidia-smi -L | head -3
free -g | head -2; nproc
timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_human_2gpu_r2.json 2> gpurun_out/bench_human_2gpu_r2.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/bench_human_2gpu_r2.err
python - <<'P'
import json
try:
    d=json.load(open("gpurun_out/bench_human_2gpu_r2.json")); print(d["value"]/1e9, d["e2e"]["value"]/1e9, d["parity"], json.dumps(d.get("sharded"))[:1500])
except Exception as e: print("failed", e)
P
