set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t11.log 2>&1; echo "tests rc=$?"; grep -n "^E  " $OUT/t11.log | head -5 | cut -c1-300; tail -3 $OUT/t11.log
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/m7.json 2> $OUT/m7.err; echo "bench rc=$?"; tail -2 $OUT/m7.err
timeout 600 python profiles/exp/gather_variants.py > $OUT/gv2_plain.log 2>&1; grep "variant [38]" $OUT/gv2_plain.log
timeout 600 python profiles/exp/e2e_stages.py mid50 > $OUT/e2e_stages2.log 2>&1; tail -16 $OUT/e2e_stages2.log
