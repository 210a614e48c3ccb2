cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/o1 timeout 300 build/fa_check bwd | tail -n 1
for r in 1 2 3; do for v in o0 o1; do printf "%s " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done; done
