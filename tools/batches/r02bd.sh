cd $GRAFT_REPO_ROOT
for v in h1 hq; do LD_LIBRARY_PATH=build/variants/$v timeout 300 build/fa_check bwd 3 1000 128 1 | tail -n 1; done
for r in 1 2 3; do for v in h0 h1 h2 hq; do printf "%s " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done; done
