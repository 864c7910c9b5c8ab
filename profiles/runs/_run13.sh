set -u
OUT=gpurun_out
timeout 1500 python profiles/exp/hits_trace.py human > $OUT/trace_human.log 2>&1; echo "rc=$?"; tail -22 $OUT/trace_human.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cpp > $OUT/h5.json 2> $OUT/h5.err; echo "bench rc=$?"
