set -u
OUT=gpurun_out
timeout 600 python profiles/exp/gather_variants.py > $OUT/gv_plain.log 2>&1; echo "plain rc=$?"; tail -17 $OUT/gv_plain.log | head -16
timeout 900 ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gather_bench --csv --log-file $OUT/gv_ncu.csv python profiles/exp/gather_variants.py > $OUT/gv_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python -X faulthandler -m pytest tests -m gpu -x -q -k "hits or seed" > $OUT/t6.log 2>&1; echo "tests rc=$?"; tail -15 $OUT/t6.log
