cd $GRAFT_REPO_ROOT
export LD_LIBRARY_PATH=build/variants/p8
ncu --set full --clock-control none -k regex:"prep|convert" -s 2 -c 2 -o gpurun_out/prof_r02e_passes -f build/fa_check time 512 4096 128 1 3 1 > gpurun_out/ncu_passes.log 2>&1
ncu -i gpurun_out/prof_r02e_passes.ncu-rep --page raw --csv > gpurun_out/prof_r02e_passes_raw.csv
ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv -k regex:"prep|convert" build/fa_check time 512 4096 128 1 3 1 2>/dev/null | grep -E "prep|convert" | awk -F'","' '{print "nocacheflush", substr($5,1,60), $NF}'
