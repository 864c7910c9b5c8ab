#!/bin/bash
# ncu captures of the kernels the round-2 session on the north-star workload does not launch: the short-pattern lookup (configs[4]),
# seed table / verification record builders, ASCII packing, result narrowing, extract. One GPU; every command runs without ncu first.
set -u
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cpp"
$B --workload chr1 > $OUT/plain_chr1.json 2> $OUT/plain_chr1.err || exit 1
timeout 900 $NCU -k regex:"seed_lookup_enum|seed_verify_rec" -s 2 -c 2 -o $OUT/prof_chr1_seed_r2 $B --workload chr1 > $OUT/n7.log 2>&1; echo "chr1 seed rc=$?"
$B --workload mid50 --ascii > $OUT/plain_mid50_ascii.json 2> $OUT/plain_mid50_ascii.err || exit 1
timeout 900 $NCU -k regex:"seed_fill|seed_build|seed_flag|vrec_build" -c 8 -o $OUT/prof_mid50_seedbuild_r2 $B --workload mid50 > $OUT/n8.log 2>&1; echo "seed build rc=$?"
timeout 900 $NCU -k regex:"pack_patterns|narrow_kernel" -s 4 -c 3 -o $OUT/prof_mid50_ascii_r2 $B --workload mid50 --ascii --e2e-api fetch > $OUT/n9.log 2>&1; echo "ascii rc=$?"
python profiles/exp/extract_probe.py mid50 > $OUT/extract_probe.log 2>&1 || exit 1
timeout 600 $NCU -k regex:"extract_kernel" -s 1 -c 1 -o $OUT/prof_mid50_extract_r2 python profiles/exp/extract_probe.py mid50 > $OUT/n10.log 2>&1; echo "extract rc=$?"
for r in $OUT/prof_chr1_seed_r2 $OUT/prof_mid50_seedbuild_r2 $OUT/prof_mid50_ascii_r2 $OUT/prof_mid50_extract_r2; do
    ncu -i $r.ncu-rep --page raw --csv > $r.raw.csv 2> /dev/null
    if [ $(stat -c %s $r.ncu-rep) -gt 8000000 ]; then rm -f $r.ncu-rep; fi
done
ls -la $OUT/prof_*_r2.raw.csv | tail -8
