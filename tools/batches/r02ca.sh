cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/fu1 timeout 300 build/fa_check tf32 | tail -n 1
for r in 1 2 3; do for v in fu0 fu1 fu1h0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time32 512 1024 64 0 50 | head -1; done; done
for r in 1 2; do for v in fu0 fu1 fu1h0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time32 512 4096 64 1 20 | head -1; done; done
for r in 1 2; do for v in fu0 fu1 fu1h0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time32 88 512 32 1 200 | head -1; done; done
for r in 1 2; do for v in fu0 fu1 fu1h0; do printf "$v "; LD_LIBRARY_PATH=build/variants/$v build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done
