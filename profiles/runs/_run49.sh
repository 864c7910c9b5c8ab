#!/bin/bash
# the two shard-set kernels under ncu on ONE GPU: a shard set of world 1 (the library's NCCL calls on a single rank) runs the same kernels
mkdir -p gpurun_out
B="python bench.py --sharded --gpus 1 --workload mid50 --steps 1 --warmup 3"
timeout 200 $B > gpurun_out/sharded_w1.json 2> gpurun_out/sharded_w1.err; rc=$?; echo "plain rc=$rc"; tail -3 gpurun_out/sharded_w1.err | cut -c1-300
[ $rc -eq 0 ] || exit 0
timeout 150 ncu --set full --clock-control none --import-source on -f -k regex:"shard_" -s 2 -c 2 -o gpurun_out/prof_mid50_shard_r2 $B > gpurun_out/n11.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_mid50_shard_r2.ncu-rep --page raw --csv > gpurun_out/prof_mid50_shard_r2.raw.csv 2>/dev/null
ls -la gpurun_out/prof_mid50_shard_r2*
