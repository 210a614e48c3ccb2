cd $GRAFT_REPO_ROOT
for r in 1 2 3; do for v in t0 t1 t2 t4 t8 t15; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
