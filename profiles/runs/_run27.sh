set -u
OUT=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests -m gpu -x -q -k "seed_search_matches_reference and col2_k2 and None" > $OUT/san.log 2>&1; echo "rc=$?"; grep -n "Invalid\|at 0x\|by thread\|Address\|kernel\|========= " $OUT/san.log | head -40 | cut -c1-250
