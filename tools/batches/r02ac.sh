cd $GRAFT_REPO_ROOT
for pd in 0.1 0; do for v in 0 1; do printf "T16384 p_drop=%s persistent=%s " $pd $v; FA_B200_BWD_PERSISTENT=$v timeout 120 build/fa_check time 128 16384 128 1 3 1 $pd | tail -1; done; done
for v in w_pf1 b0v0 e0 r01; do printf "%s " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -1; done
for pd in 0.1; do for v in 0 1; do printf "T4096 p_drop=%s persistent=%s " $pd $v; FA_B200_BWD_PERSISTENT=$v timeout 120 build/fa_check time 512 4096 128 1 5 1 $pd | tail -1; done; done
