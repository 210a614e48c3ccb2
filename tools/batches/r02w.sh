cd $GRAFT_REPO_ROOT
bash tools/ab.sh "time 512 4096 128 1 10 1" w_pf1 a0v0 a1v0 a0v1 a1v1 > gpurun_out/r02w_ab.log 2>&1; cat gpurun_out/r02w_ab.log
for s in 24 44; do LD_LIBRARY_PATH=build/variants/a1v1 timeout 120 build/fa_check trace 512 4096 128 1 $s > gpurun_out/r02w_trace${s}_a1v1.log 2>&1; done
for s in 24 44; do LD_LIBRARY_PATH=build/variants/a0v0 timeout 120 build/fa_check trace 512 4096 128 1 $s > gpurun_out/r02w_trace${s}_a0v0.log 2>&1; done
