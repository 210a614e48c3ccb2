cd $GRAFT_REPO_ROOT
for v in nosig skip3 f4k; do LD_LIBRARY_PATH=build/variants/$v ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,sm__cycles_active.avg --clock-control none --csv -k regex:"fa_bwd_kernel" build/fa_check time 512 4096 128 1 2 1 2>/dev/null | grep -E "fa_bwd_kernel" | awk -F'","' -v v=$v '{print v, $(NF-2), $NF}'; done
FA_B200_BWD_FUSED_PASSES=0 LD_LIBRARY_PATH=build/variants/f4k ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,sm__cycles_active.avg --clock-control none --csv -k regex:"fa_bwd_kernel" build/fa_check time 512 4096 128 1 2 1 2>/dev/null | grep -E "fa_bwd_kernel" | awk -F'","' -v v=classic '{print v, $(NF-2), $NF}'
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
(for i in 1 2 3 4 5 6 7 8 9 10 11 12; do nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader; sleep 0.25; done) &
FA_B200_BWD_FUSED_PASSES=0 LD_LIBRARY_PATH=build/variants/f4k build/fa_check time 512 4096 128 1 300 1 | tail -1
wait
(for i in 1 2 3 4 5 6 7 8 9 10 11 12; do nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader; sleep 0.25; done) &
LD_LIBRARY_PATH=build/variants/f4k build/fa_check time 512 4096 128 1 300 1 | tail -1
wait
