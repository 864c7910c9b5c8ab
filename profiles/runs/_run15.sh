set -u
OUT=gpurun_out
nvidia-smi -L
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q -k "sharded_two_gpus" > $OUT/t13.log 2>&1; echo "tests rc=$?"; grep -n "^E  " $OUT/t13.log | head -8 | cut -c1-400; tail -3 $OUT/t13.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/m9_2gpu.json 2> $OUT/m9_2gpu.err; echo "bench2 rc=$?"; tail -5 $OUT/m9_2gpu.err | cut -c1-300
