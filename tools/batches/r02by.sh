cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/pr0 timeout 300 build/fa_check bwd | tail -n 1
LD_LIBRARY_PATH=build/variants/pr0 timeout 120 build/fa_check bwd 40 1000 128 1 | tail -n 2
LD_LIBRARY_PATH=build/variants/pr0 timeout 120 build/fa_check bwd 37 2125 64 0 | tail -n 2
for r in 1 2 3; do for v in pr1 pr0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
for r in 1 2; do for v in pr1 pr0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 0 5 1 | tail -1; done; done
for r in 1 2; do for v in pr1 pr0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 64 1 10 1 | tail -1; done; done
sleep 2
for v in pr1 pr0; do LD_LIBRARY_PATH=build/variants/$v tools/sustained_probe.sh "$v bwd x300" build/fa_check time 512 4096 128 1 300 1; sleep 2; done
