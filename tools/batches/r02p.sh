set -x
cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check bwd > gpurun_out/r02p_bwd.log 2>&1; tail -3 gpurun_out/r02p_bwd.log
FA_B200_BWD_PERSISTENT=0 timeout 300 build/fa_check bwd > gpurun_out/r02p_bwd_np.log 2>&1; tail -2 gpurun_out/r02p_bwd_np.log
timeout 300 build/fa_check drop > gpurun_out/r02p_drop.log 2>&1; tail -2 gpurun_out/r02p_drop.log
for r in 1 2 3; do
  for v in 0 1; do printf "persistent=%s " $v; FA_B200_BWD_PERSISTENT=$v timeout 120 build/fa_check time 512 4096 128 1 10 1 | tail -1; done
done > gpurun_out/r02p_ab.log 2>&1; cat gpurun_out/r02p_ab.log
for v in 0 1; do printf "noncausal persistent=%s " $v; FA_B200_BWD_PERSISTENT=$v timeout 120 build/fa_check time 512 4096 128 0 5 1 | tail -1; done > gpurun_out/r02p_ab_nc.log 2>&1; cat gpurun_out/r02p_ab_nc.log
for v in 0 1; do printf "cfg4 persistent=%s " $v; FA_B200_BWD_PERSISTENT=$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 | tail -1; done > gpurun_out/r02p_ab_cfg4.log 2>&1; cat gpurun_out/r02p_ab_cfg4.log
timeout 120 build/fa_check trace 512 4096 128 1 0 > gpurun_out/r02p_trace0.log 2>&1
timeout 120 build/fa_check trace 512 4096 128 1 24 > gpurun_out/r02p_trace24.log 2>&1; cat gpurun_out/r02p_trace24.log
