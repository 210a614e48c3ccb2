cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/f4k ncu --set full --clock-control none -k regex:fa_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_f4k -f build/fa_check time 512 4096 128 1 2 1 > /dev/null 2>&1
FA_B200_BWD_FUSED_PASSES=0 LD_LIBRARY_PATH=build/variants/f4k ncu --set full --clock-control none -k regex:fa_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_classic -f build/fa_check time 512 4096 128 1 2 1 > /dev/null 2>&1
LD_LIBRARY_PATH=build/variants/nosig ncu --set full --clock-control none -k regex:fa_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_nosig -f build/fa_check time 512 4096 128 1 2 1 > /dev/null 2>&1
for n in f4k classic nosig; do ncu -i gpurun_out/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv; rm gpurun_out/prof_$n.ncu-rep; done
