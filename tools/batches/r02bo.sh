cd $GRAFT_REPO_ROOT
P=tools/sustained_probe.sh
for v in notrace; do LD_LIBRARY_PATH=build/variants/$v timeout 300 build/fa_check bwd | tail -n 1; LD_LIBRARY_PATH=build/variants/$v timeout 300 build/fa_check fwd | tail -n 1; done
for r in 1 2; do for v in base notrace; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
for r in 1 2; do for v in base notrace; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
sleep 2
for r in 1 2; do for v in base notrace; do LD_LIBRARY_PATH=build/variants/$v $P "$v bwd x300" build/fa_check time 512 4096 128 1 300 1; sleep 2; done; done
for r in 1 2; do for v in base notrace; do LD_LIBRARY_PATH=build/variants/$v $P "$v fwd x1000" build/fa_check time 512 4096 128 1 1000 0; sleep 2; done; done
