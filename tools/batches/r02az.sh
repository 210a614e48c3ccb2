cd $GRAFT_REPO_ROOT
for r in 1 2 3; do for v in old new; do printf "%s cfg3 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -2 | tr '\n' ' '; echo; done; done
