set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t9.log 2>&1; echo "tests rc=$?"; grep -n "^E " $OUT/t9.log | head -5 | cut -c1-300; tail -3 $OUT/t9.log
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp --no-cpu-baseline > $OUT/m5.json 2> $OUT/m5.err; echo "bench rc=$?"; tail -2 $OUT/m5.err
