set -u
OUT=gpurun_out
timeout 900 python profiles/exp/lookup_probe.py mid50 > $OUT/lookup_probe.log 2>&1; echo "rc=$?"; tail -4 $OUT/lookup_probe.log
