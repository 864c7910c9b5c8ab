set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t10.log 2>&1; echo "tests rc=$?"; grep -n "^E  " $OUT/t10.log | head -5 | cut -c1-300; tail -3 $OUT/t10.log
timeout 900 python bench.py --workload mid50 --steps 5 --warmup 3 --no-cpp --no-cpu-baseline > $OUT/m6.json 2> $OUT/m6.err; echo "bench rc=$?"; tail -2 $OUT/m6.err
B="python bench.py --workload mid50 --steps 1 --warmup 3 --no-cpu-baseline --no-cpp"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"seed_lookup|seed_verify" -s 2 -c 2 -o $OUT/prof_mid50_seed_b_r2 -f $B > $OUT/n_m6.log 2>&1; echo "ncu rc=$?"
