cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/w1h0 timeout 200 build/fa_check fwd | tail -n 1
for r in 1 2 3; do for v in h6 w0h0 w1h0 w1h6; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
for r in 1 2; do for v in w0h0 w1h0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 10 1 | tail -1; done; done
