set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $OUT/t3.log; tail -3 $OUT/t3.log
nproc; free -g | head -2
timeout 1500 python bench.py --steps 5 --warmup 3 > $OUT/h3.json 2> $OUT/h3.err; echo "bench rc=$?"; tail -3 $OUT/h3.err
timeout 600 python bench.py --steps 5 --warmup 3 --ascii --no-cpu-baseline --no-cpp > $OUT/h3_ascii.json 2> $OUT/h3_ascii.err; echo "ascii rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_human.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cpp > $OUT/n_h.log 2>&1; echo "ncu rc=$?"
