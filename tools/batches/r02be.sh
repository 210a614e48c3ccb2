cd $GRAFT_REPO_ROOT
for v in p8; do LD_LIBRARY_PATH=build/variants/$v timeout 300 build/fa_check bwd | tail -n 1; done
for v in base p8; do LD_LIBRARY_PATH=build/variants/$v ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"prep|convert" build/fa_check time 512 4096 128 1 3 1 2>/dev/null | grep -E "prep|convert" | awk -F'","' -v v=$v '{print v, substr($5,1,60), $NF}'; done
bash tools/ab.sh "time 512 4096 128 1 10 1" base p8
