set -u
OUT=gpurun_out
E2FM_DEBUG_SYNC=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/t17.log 2>&1; echo "rc=$?"; grep -n "^E  " $OUT/t17.log | head -8 | cut -c1-500; tail -3 $OUT/t17.log
