cd $GRAFT_REPO_ROOT
for r in 1 2 3; do for v in cur pf0 hg8 hg32; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
sleep 2
for v in cur pf0 hg8; do LD_LIBRARY_PATH=build/variants/$v tools/sustained_probe.sh "$v bwd x300" build/fa_check time 512 4096 128 1 300 1; sleep 2; done
