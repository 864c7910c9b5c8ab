#!/bin/bash
# N=2 path of bench.py after the host-buffer changes (deferred page-locking, CPU affinity, host-link figure), mid-sized workload
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload mid50 --steps 5 --warmup 3 > gpurun_out/bench_mid50_2gpu_r2b.json 2> gpurun_out/bench_mid50_2gpu_r2b.err; echo "bench2 rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_mid50_2gpu_r2b.err | tail -12
python - <<'P'
import json
try:
    d=json.load(open("gpurun_out/bench_mid50_2gpu_r2b.json")); print(d["value"]/1e9, d["e2e"]["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("host_link"), d["parity"], (d.get("sharded") or {}).get("value"), (d.get("sharded") or {}).get("same_answers_as_replicas"), (d.get("sharded") or {}).get("error"))
except Exception as e: print("failed", e)
P
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
