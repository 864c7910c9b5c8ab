set -u
OUT=gpurun_out
timeout 900 python profiles/exp/e2e_stages.py mid50 > $OUT/e2e_stages.log 2>&1; echo "rc=$?"; tail -18 $OUT/e2e_stages.log
