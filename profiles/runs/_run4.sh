set -u
OUT=gpurun_out
B="python bench.py --workload mid50 --steps 1 --warmup 3 --no-cpu-baseline --no-cpp"
$B > $OUT/m2_plain.json 2> $OUT/m2_plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"seed_lookup|seed_verify" -s 2 -c 2 -o $OUT/prof_mid50_seed_r2 -f $B > $OUT/n_m2.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/*.ncu-rep
