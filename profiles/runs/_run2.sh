set -u
OUT=gpurun_out
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q > $OUT/t4.log 2>&1; echo "tests rc=$?"; grep -n "passed\|failed\|Fatal\|Segmentation\|Error" $OUT/t4.log | head
timeout 300 python profiles/exp/granularity.py > $OUT/gran.log 2>&1; tail -1 $OUT/gran.log
