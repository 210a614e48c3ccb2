cd $GRAFT_REPO_ROOT
for r in 1 2; do for v in new sr1 sr1u4 sr0u4; do printf "%s cfg4 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -1; done; done
LD_LIBRARY_PATH=build/variants/sr1 timeout 300 build/fa_check drop | tail -n 1
