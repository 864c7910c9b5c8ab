set -u
OUT=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/t18.log 2>&1; echo "rc=$?"; grep -n "^E  " $OUT/t18.log | head -8 | cut -c1-500; tail -3 $OUT/t18.log
