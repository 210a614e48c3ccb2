cd $GRAFT_REPO_ROOT
for r in 1 2 3; do for v in h14 h2 h4 h8 h0 h6 h12 h10; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
