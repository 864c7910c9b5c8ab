#!/bin/bash
mkdir -p gpurun_out
timeout 900 python profiles/exp/encode_bench.py mid50 > gpurun_out/encode_mid50.json 2> gpurun_out/encode_mid50.err; echo "rc=$?"; cat gpurun_out/encode_mid50.json; tail -2 gpurun_out/encode_mid50.err
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:"bucket_mtf_rle|payload_offsets|bucket_substitute_pack|salsa20" -c 4 -o gpurun_out/prof_mid50_encode_r2 python profiles/exp/encode_bench.py mid50 > gpurun_out/n5.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_mid50_encode_r2.ncu-rep --page raw --csv > gpurun_out/prof_mid50_encode_r2.raw.csv 2>/dev/null
ls -la gpurun_out/prof_mid50_encode_r2*
