cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check bwd > gpurun_out/r02x_bwd.log 2>&1; tail -n 1 gpurun_out/r02x_bwd.log
timeout 300 build/fa_check drop > gpurun_out/r02x_drop.log 2>&1; tail -n 1 gpurun_out/r02x_drop.log
bash tools/ab.sh "time 512 4096 128 1 10 1" w_pf1 b0v0 b1v0 b1v1 > gpurun_out/r02x_ab.log 2>&1; cat gpurun_out/r02x_ab.log
for s in 24 44; do LD_LIBRARY_PATH=build/variants/b1v1 timeout 120 build/fa_check trace 512 4096 128 1 $s > gpurun_out/r02x_trace${s}_b1v1.log 2>&1; done
