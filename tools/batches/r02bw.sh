cd $GRAFT_REPO_ROOT
for r in 1 2 3; do for v in u4 u8 u2 d4; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
for r in 1 2; do for v in u4 u8 u2 d4; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 128 16384 128 1 3 1 0.1 | tail -1; done; done
