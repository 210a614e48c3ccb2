cd $GRAFT_REPO_ROOT
export LD_LIBRARY_PATH=build/variants/uni
timeout 300 build/fa_check fwd | tail -n 1
timeout 300 build/fa_check bwd | tail -n 1
timeout 300 build/fa_check drop | tail -n 1
timeout 300 build/fa_check tf32 | tail -n 1
timeout 300 build/fa_check bwd32 | tail -n 1
unset LD_LIBRARY_PATH
for r in 1 2 3; do for v in prev uni; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
for r in 1 2 3; do for v in prev uni; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
for v in prev uni; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time32 512 1024 64 0 20 | tail -1; done
for v in prev uni; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 64 1 20 0 | tail -1; done
for v in prev uni; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 64 1 10 1 | tail -1; done
for v in prev uni; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 128 16384 128 1 3 1 0.1 | tail -1; done
sleep 2
for v in prev uni; do LD_LIBRARY_PATH=build/variants/$v tools/sustained_probe.sh "$v bwd x300" build/fa_check time 512 4096 128 1 300 1; sleep 2; done
for v in prev uni; do LD_LIBRARY_PATH=build/variants/$v tools/sustained_probe.sh "$v fwd x1000" build/fa_check time 512 4096 128 1 1000 0; sleep 2; done
