cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check bwd > gpurun_out/r02v_bwd.log 2>&1; tail -n 1 gpurun_out/r02v_bwd.log
timeout 300 build/fa_check drop > gpurun_out/r02v_drop.log 2>&1; tail -n 1 gpurun_out/r02v_drop.log
mkdir -p build/variants/cur; cp flash-attention-minitorch_b200/lib/libfa_b200.so build/variants/cur/
bash tools/ab.sh "time 512 4096 128 1 10 1" w_pf1 cur > gpurun_out/r02v_ab.log 2>&1; cat gpurun_out/r02v_ab.log
timeout 120 build/fa_check trace 512 4096 128 1 20 > gpurun_out/r02v_trace20.log 2>&1
cat gpurun_out/r02v_trace20.log
