set -u
OUT=gpurun_out
E2FM_DEBUG_SYNC=1 timeout 900 python -m pytest tests -m gpu -x -q -k "seed_search_matches_reference and col2_k2 and None" > $OUT/t16.log 2>&1; echo "rc=$?"; grep -n "^E  " $OUT/t16.log | head -8 | cut -c1-400; tail -3 $OUT/t16.log
