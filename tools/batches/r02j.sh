set -x
cd $GRAFT_REPO_ROOT
timeout 120 build/fa_check fwd 2 512 128 1 > gpurun_out/r02j_fwd1.log 2>&1; cat gpurun_out/r02j_fwd1.log
timeout 120 build/fa_check fwd 2 512 128 0 >> gpurun_out/r02j_fwd1.log 2>&1; tail -2 gpurun_out/r02j_fwd1.log
timeout 300 build/fa_check fwd > gpurun_out/r02j_fwd.log 2>&1; grep -c OK gpurun_out/r02j_fwd.log; grep -v OK gpurun_out/r02j_fwd.log | tail
for r in 1 2 3; do
  for v in 0 1; do printf "pair=%s " $v; FA_B200_FWD_PAIR=$v timeout 120 build/fa_check time 512 4096 128 1 20 0 | tail -1; done
done > gpurun_out/r02j_ab.log 2>&1; cat gpurun_out/r02j_ab.log
for v in 0 1; do printf "pair=%s " $v; FA_B200_FWD_PAIR=$v timeout 120 build/fa_check time 512 4096 128 0 10 0 | tail -1; done >> gpurun_out/r02j_ab.log 2>&1; tail -2 gpurun_out/r02j_ab.log
