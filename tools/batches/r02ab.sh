cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check fwd > gpurun_out/r02ab_fwd.log 2>&1; tail -n 1 gpurun_out/r02ab_fwd.log
timeout 300 build/fa_check drop > gpurun_out/r02ab_drop.log 2>&1; tail -n 1 gpurun_out/r02ab_drop.log
timeout 300 build/fa_check tf32 > gpurun_out/r02ab_tf32.log 2>&1; tail -n 1 gpurun_out/r02ab_tf32.log
for r in 1 2 3; do for v in 0 1; do printf "fwd persistent=%s " $v; FA_B200_FWD_PERSISTENT=$v timeout 120 build/fa_check time 512 4096 128 1 20 0 | tail -1; done; done > gpurun_out/r02ab_ab.log 2>&1; cat gpurun_out/r02ab_ab.log
for v in 0 1; do printf "noncausal fwd persistent=%s " $v; FA_B200_FWD_PERSISTENT=$v timeout 120 build/fa_check time 512 4096 128 0 10 0 | tail -1; done
for v in 0 1; do printf "64 heads fwd persistent=%s " $v; FA_B200_FWD_PERSISTENT=$v timeout 120 build/fa_check time 64 4096 128 1 50 0 | tail -1; done
for v in 0 1; do printf "cfg2 tf32 persistent=%s " $v; FA_B200_FWD_PERSISTENT=$v timeout 120 build/fa_check time32 512 1024 64 0 50 | tail -1; done
