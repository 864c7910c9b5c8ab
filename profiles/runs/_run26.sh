set -u
OUT=gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t15.log 2>&1; echo "tests rc=$?"; grep -n "^E  " $OUT/t15.log | head -8 | cut -c1-400; tail -3 $OUT/t15.log
timeout 900 python bench.py --workload chr1 --steps 5 --warmup 3 --no-cpp > $OUT/bench_chr1_r2.json 2> $OUT/bench_chr1_r2.err; echo "chr1 rc=$?"; tail -3 $OUT/bench_chr1_r2.err | cut -c1-300
