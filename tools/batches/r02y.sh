cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/e1 timeout 300 build/fa_check bwd > gpurun_out/r02y_bwd.log 2>&1; tail -n 1 gpurun_out/r02y_bwd.log
bash tools/ab.sh "time 512 4096 128 1 10 1" w_pf1 e0 e1 > gpurun_out/r02y_ab.log 2>&1; cat gpurun_out/r02y_ab.log
for s in 24 44; do LD_LIBRARY_PATH=build/variants/e1 timeout 120 build/fa_check trace 512 4096 128 1 $s > gpurun_out/r02y_trace${s}_e1.log 2>&1; done
