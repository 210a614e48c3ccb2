cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/ps1 timeout 200 build/fa_check fwd | tail -n 1
for r in 1 2 3; do for v in ps2 ps1; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
sleep 2
for v in ps2 ps1; do LD_LIBRARY_PATH=build/variants/$v tools/sustained_probe.sh "$v fwd x1000" build/fa_check time 512 4096 128 1 1000 0; sleep 2; done
