cd $GRAFT_REPO_ROOT
for v in f4k f4knw nosig skip3 skip1 skip2; do test -f build/variants/$v/libfa_b200.so || echo MISSING $v; done
for r in 1 2 3; do for v in f4k f4knw nosig skip3 skip1 skip2; do printf "%s " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done; 
printf "classic "; FA_B200_BWD_FUSED_PASSES=0 LD_LIBRARY_PATH=build/variants/f4k timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done
