set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t12.log 2>&1; echo "tests rc=$?"; grep -n "^E  " $OUT/t12.log | head -5 | cut -c1-300; tail -3 $OUT/t12.log
timeout 600 python profiles/exp/hits_trace.py mid50 > $OUT/trace_mid50.log 2>&1; echo "rc=$?"; tail -8 $OUT/trace_mid50.log
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline --no-cpp > $OUT/m8.json 2> $OUT/m8.err; echo "bench rc=$?"
