cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check bwd > gpurun_out/r02bb_bwd.log 2>&1; tail -n 1 gpurun_out/r02bb_bwd.log
timeout 300 build/fa_check drop | tail -n 1
timeout 600 python tools/fuzz_persistent.py 80 51 | tail -n 2
for BH in 512 64; do for r in 1 2 3; do for v in dk0 dk1; do printf "%s BH=%s " $v $BH; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time $BH 4096 128 1 10 1 2>&1 | tail -1; done; done; done
for v in dk0 dk1; do printf "%s cfg4 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -1; done
for v in dk0 dk1; do printf "%s d64 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 64 1 10 1 2>&1 | tail -1; done
for v in dk0 dk1; do printf "%s noncausal " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 0 5 1 2>&1 | tail -1; done
