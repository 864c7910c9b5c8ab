#!/bin/bash
# One GPU-box session that produces everything profiles/ holds for a round. Run it through gpurun from the repo root:
#
#   gpurun --timeout 2400 -- 'bash profiles/capture.sh r2'
#
# then, back in the authoring container:  python profiles/summarize.py r2
#
# Order matters: every command first runs WITHOUT ncu (a bench value is never taken under the profiler, and a program
# that fails plain is not profiled), then under ncu. ncu runs are single-GPU. -s/-c count launches that match -k: with
# 3 warm-up steps + 1 timed step + the end-to-end pass in front, the skips below land on a device-resident timed step;
# if the number of launches per step changes, re-derive them from the launch list (launches_*.csv).
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on"

# 1. parity first
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > $OUT/tests_$TAG.log

# 2. bench lines of every single-GPU workload (full contract: CPU baseline, clocks, e2e)
for w in ecoli config1 collection25 chr1; do
    timeout 600 python bench.py --workload $w > $OUT/bench_${w}_$TAG.json 2> $OUT/bench_${w}_$TAG.err
done

# 3. launch list of the default bench command (shares of the step)
$B > $OUT/plain_e.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_ecoli.csv $B > $OUT/n0.log 2>&1

# 4. full captures: search kernels, L2-resident (configs[1])
$NCU -k regex:"backward_search|encode_patterns|resolve_kernel|scan_expand" -s 60 -c 4 -o $OUT/prof_ecoli_$TAG $B > $OUT/n1.log 2>&1
# load pipeline (7083 buckets)
$NCU -k regex:"salsa20_keystream|bucket_decode|rank_fill|dense_build|sb_parse|pair_build" -c 6 -o $OUT/prof_load_$TAG $B --workload collection25 > $OUT/n2.log 2>&1
# search kernels, HBM-resident (25 genomes)
$B --workload collection25 > $OUT/plain_c.log 2>&1 &&
    $NCU -k regex:"backward_search|resolve_kernel" -s 102 -c 2 -o $OUT/prof_c25_$TAG $B --workload collection25 > $OUT/n3.log 2>&1
# per-query walks (dense tables off)
$B --workload collection25 --walk > $OUT/plain_w.log 2>&1 &&
    $NCU -k regex:"walk_kernel" -s 51 -c 1 -o $OUT/prof_c25walk_$TAG $B --workload collection25 --walk > $OUT/n4.log 2>&1
# wildcard scan + resolve on short patterns (configs[4])
$B --workload chr1 > $OUT/plain_h.log 2>&1 &&
    $NCU -k regex:"firstcheck|resolve_kernel" -s 96 -c 2 -o $OUT/prof_chr1_$TAG $B --workload chr1 > $OUT/n5.log 2>&1

cat $OUT/tests_$TAG.log
ls -la $OUT/*_$TAG.ncu-rep
