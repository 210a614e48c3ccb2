cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check fwd | tail -n 1; timeout 300 build/fa_check bwd | tail -n 1; timeout 300 build/fa_check drop | tail -n 1
timeout 600 python tools/fuzz_persistent.py 60 21 | tail -n 2
for BH in 512 64; do for r in 1 2 3; do for v in hm cur; do printf "%s BH=%s " $v $BH; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time $BH 4096 128 1 20 1 2>&1 | tail -2 | tr '\n' ' '; echo; done; done; done
