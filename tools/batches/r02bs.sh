cd $GRAFT_REPO_ROOT
for r in 1 2 3 4; do for v in f3 f1 f2 f0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
