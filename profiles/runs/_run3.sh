set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t5.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/t5.log
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp > $OUT/m1.json 2> $OUT/m1.err; echo "bench rc=$?"; tail -2 $OUT/m1.err
E2FM_SEED=0 timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp --no-cpu-baseline > $OUT/m1_noseed.json 2> $OUT/m1_noseed.err; echo "bench rc=$?"
