cd $GRAFT_REPO_ROOT
for r in 1 2; do for v in d1 d2 d2p0; do printf "%s cfg4 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -1; done; done
for v in d1 d2; do printf "%s T4096 drop " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 5 1 0.1 2>&1 | tail -1; done
LD_LIBRARY_PATH=build/variants/d2 timeout 200 build/fa_check drop | tail -1
