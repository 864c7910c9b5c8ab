#!/bin/bash
# One GPU-box session for the round-2 evidence on the north-star workload (configs[3], 3.1 Gbp): the bench line, the launch list and
# the ncu captures of every kernel group. Run through gpurun from the repo root:   gpurun --timeout 3000 -- 'bash profiles/capture_r2.sh'
# then, in the authoring container:   python profiles/summarize.py r2
# Every command runs WITHOUT ncu first (a bench value is never taken under the profiler). -s/-c count launches that match -k.
set -u
OUT=gpurun_out
mkdir -p $OUT
W=${1:-human}
B="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-cpp"
NCU="ncu --set full --clock-control none --import-source on -f"
# 1. the bench line of the default command (CPU baseline, e2e, parity, C++ host figure)
timeout 1800 python bench.py --workload $W --steps 10 --warmup 3 > $OUT/bench_${W}_r2.json 2> $OUT/bench_${W}_r2.err; echo "bench rc=$?"
$B > $OUT/plain_${W}.json 2> $OUT/plain_${W}.err || exit 1
# 2. launch list (shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_${W}_r2.csv $B > $OUT/n0.log 2>&1; echo "launch list rc=$?"
# 3. the two search kernels on a device-resident 4M-pattern chunk (the probe call launches them once each first)
timeout 900 $NCU -k regex:"seed_lookup|seed_verify" -s 2 -c 2 -o $OUT/prof_${W}_seed_r2 $B > $OUT/n1.log 2>&1; echo "seed rc=$?"
# 4. result pipeline of the device-resident batch (6 launches of the probe first), then of one pipeline stage of e2fm_search_hits
timeout 900 $NCU -k regex:"scan_block_sums|scan_sums|scan_apply|scatter_kernel|segment_sort|map_positions" -s 6 -c 6 -o $OUT/prof_${W}_results_r2 $B > $OUT/n2.log 2>&1; echo "results rc=$?"
timeout 900 $NCU -k regex:"hits_" -c 5 -o $OUT/prof_${W}_hits_r2 $B > $OUT/n3.log 2>&1; echo "hits rc=$?"
# 5. load pipeline
timeout 1200 $NCU -k regex:"salsa20|bucket_decode|sb_parse|rank_|dense_build|bwt16_build|pair_build|seed_fill|seed_build" -c 16 -o $OUT/prof_${W}_load_r2 $B > $OUT/n4.log 2>&1; echo "load rc=$?"
# only 64 MiB travel back: keep the raw-page CSV of every capture, and the report itself only when it is small
for r in $OUT/prof_${W}_*_r2.ncu-rep; do
    ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2> /dev/null
    if [ $(stat -c %s $r) -gt 12000000 ]; then rm -f $r; fi
done
rm -f $OUT/*.tmp
ls -la $OUT/ | tail -30
du -sh $OUT
