cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check bwd > gpurun_out/r02r_bwd.log 2>&1; tail -n 1 gpurun_out/r02r_bwd.log
bash tools/ab.sh "time 512 4096 128 1 10 1" r01 u1 u2 u4 u8 > gpurun_out/r02r_ab.log 2>&1; cat gpurun_out/r02r_ab.log
for v in u1 u8; do LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check trace 512 4096 128 1 20 > gpurun_out/r02r_trace20_$v.log 2>&1; done
cat gpurun_out/r02r_trace20_u1.log
