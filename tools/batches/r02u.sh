cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check bwd > gpurun_out/r02u_bwd.log 2>&1; tail -n 1 gpurun_out/r02u_bwd.log
timeout 300 build/fa_check drop > gpurun_out/r02u_drop.log 2>&1; tail -n 1 gpurun_out/r02u_drop.log
bash tools/ab.sh "time 512 4096 128 1 10 1" u4 w_pf0 w_pf1 > gpurun_out/r02u_ab.log 2>&1; cat gpurun_out/r02u_ab.log
for v in w_pf1; do LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check trace 512 4096 128 1 20 > gpurun_out/r02u_trace20_$v.log 2>&1; done
cat gpurun_out/r02u_trace20_w_pf1.log
