set -u
OUT=gpurun_out
for w in chr1 collection100 ecoli config1; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpp > $OUT/bench_${w}_r2.json 2> $OUT/bench_${w}_r2.err; echo "$w rc=$?"; tail -1 $OUT/bench_${w}_r2.err | cut -c1-200
done
