set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t7.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/t7.log
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp --no-cpu-baseline > $OUT/m3.json 2> $OUT/m3.err; echo "bench rc=$?"; tail -2 $OUT/m3.err
E2FM_SEED=0 timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp --no-cpu-baseline > $OUT/m3_noseed.json 2> $OUT/m3_noseed.err; echo "bench rc=$?"
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp --no-cpu-baseline --e2e-api fetch > $OUT/m3_fetch.json 2> $OUT/m3_fetch.err; echo "bench rc=$?"
