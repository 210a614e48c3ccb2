cd $GRAFT_REPO_ROOT
for r in 1 2; do for v in fused skip1 skip2 skip3; do printf "%s " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done; 
printf "classic "; FA_B200_BWD_FUSED_PASSES=0 LD_LIBRARY_PATH=build/variants/fused timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done
LD_LIBRARY_PATH=build/variants/fused ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"fa_bwd" build/fa_check time 512 4096 128 1 3 1 2>/dev/null | grep -E "fa_bwd" | awk -F'","' '{print "fused", substr($5,1,50), $NF}'
FA_B200_BWD_FUSED_PASSES=0 LD_LIBRARY_PATH=build/variants/fused ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"fa_bwd" build/fa_check time 512 4096 128 1 3 1 2>/dev/null | grep -E "fa_bwd" | awk -F'","' '{print "classic", substr($5,1,50), $NF}'
