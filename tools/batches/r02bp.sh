cd $GRAFT_REPO_ROOT
P=tools/sustained_probe.sh
for r in 1 2 3; do for v in base notrace bar; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
for r in 1 2; do for v in notrace bar; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
sleep 2
for v in base notrace bar; do LD_LIBRARY_PATH=build/variants/$v $P "$v fwd x1000" build/fa_check time 512 4096 128 1 1000 0; sleep 2; done
for v in notrace bar; do LD_LIBRARY_PATH=build/variants/$v $P "$v bwd x300" build/fa_check time 512 4096 128 1 300 1; sleep 2; done
