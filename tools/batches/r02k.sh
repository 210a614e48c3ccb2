set -x
cd $GRAFT_REPO_ROOT
timeout 120 build/fa_check fwd 2 1000 128 1 > gpurun_out/r02k_fwd1.log 2>&1; tail -2 gpurun_out/r02k_fwd1.log
for r in 1 2; do
  for v in 0 1; do printf "pair=%s " $v; FA_B200_FWD_PAIR=$v timeout 120 build/fa_check time 512 4096 128 1 20 0 | tail -1; done
done > gpurun_out/r02k_ab.log 2>&1; cat gpurun_out/r02k_ab.log
timeout 120 build/fa_check tracefwd > gpurun_out/r02k_tracefwd.log 2>&1; cat gpurun_out/r02k_tracefwd.log
