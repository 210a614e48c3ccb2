cd $GRAFT_REPO_ROOT
export LD_LIBRARY_PATH=build/variants/g2u4
timeout 300 build/fa_check bwd > gpurun_out/r02t_bwd.log 2>&1; tail -n 1 gpurun_out/r02t_bwd.log
unset LD_LIBRARY_PATH
bash tools/ab.sh "time 512 4096 128 1 10 1" u4 g2u4 g2u8 > gpurun_out/r02t_ab.log 2>&1; cat gpurun_out/r02t_ab.log
for v in g2u4; do LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check trace 512 4096 128 1 0 > gpurun_out/r02t_trace0_$v.log 2>&1; done
cat gpurun_out/r02t_trace0_g2u4.log
