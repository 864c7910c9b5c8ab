set -u
OUT=gpurun_out
E2FM_SEED_DEBUG=1 timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline --no-cpp > $OUT/m12.json 2> $OUT/m12.err; echo "bench rc=$?"; tail -3 $OUT/m12.err; head -c 600 $OUT/m12.json
