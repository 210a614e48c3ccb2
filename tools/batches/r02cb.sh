cd $GRAFT_REPO_ROOT
export LD_LIBRARY_PATH=build/variants/new
timeout 300 build/fa_check bwd | tail -n 1
timeout 300 build/fa_check drop | tail -n 1
timeout 300 build/fa_check tf32 | tail -n 1
timeout 300 build/fa_check fwd | tail -n 1
timeout 120 build/fa_check bwd 40 1000 128 1 | tail -n 1
unset LD_LIBRARY_PATH
for r in 1 2 3; do for v in head new; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
for r in 1 2; do for v in head new; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 128 16384 128 1 3 1 0.1 | tail -1; done; done
for r in 1 2; do for v in head new; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 64 1 10 1 | tail -1; done; done
sleep 2
for r in 1 2; do for v in head new; do LD_LIBRARY_PATH=build/variants/$v tools/sustained_probe.sh "$v bwd x300" build/fa_check time 512 4096 128 1 300 1; sleep 2; done; done
