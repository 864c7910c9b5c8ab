set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q -k "hits" > $OUT/t8.log 2>&1; echo "tests rc=$?"; tail -12 $OUT/t8.log | cut -c1-300
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp --no-cpu-baseline > $OUT/m4.json 2> $OUT/m4.err; echo "bench rc=$?"; tail -2 $OUT/m4.err
