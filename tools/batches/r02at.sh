cd $GRAFT_REPO_ROOT
for r in 1 2 3; do for v in old new; do printf "%s cfg4 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -2 | tr '\n' ' '; echo; done; done
for r in 1 2; do for v in old new; do printf "%s T16384 nodrop " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 0 2>&1 | tail -1; done; done
