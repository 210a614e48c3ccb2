cd $GRAFT_REPO_ROOT
for r in 1 2 3; do for v in cur pf0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 64 4096 128 1 30 1 | tail -1; done; done
for r in 1 2; do for v in cur pf0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 128 16384 128 1 3 1 0.1 | tail -1; done; done
for r in 1 2; do for v in cur pf0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 0 5 1 | tail -1; done; done
for r in 1 2; do for v in cur pf0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 64 1 10 1 | tail -1; done; done
