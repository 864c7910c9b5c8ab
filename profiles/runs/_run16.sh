set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t14.log 2>&1; echo "tests rc=$?"; grep -n "^E  " $OUT/t14.log | head -5 | cut -c1-300; tail -3 $OUT/t14.log
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline --no-cpp > $OUT/m10.json 2> $OUT/m10.err; echo "bench rc=$?"; tail -3 $OUT/m10.err
