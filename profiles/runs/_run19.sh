set -u
OUT=gpurun_out
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline --no-cpp > $OUT/m11.json 2> $OUT/m11.err; echo "bench rc=$?"; tail -3 $OUT/m11.err
