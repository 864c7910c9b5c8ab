set -u
OUT=gpurun_out
nvidia-smi -L | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/m13_8gpu.json 2> $OUT/m13_8gpu.err; echo "bench8 rc=$?"; tail -4 $OUT/m13_8gpu.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/m13_4gpu.json 2> $OUT/m13_4gpu.err; echo "bench4 rc=$?"; tail -2 $OUT/m13_4gpu.err | cut -c1-300
