cd $GRAFT_REPO_ROOT
export LD_LIBRARY_PATH=build/variants/h4
timeout 300 build/fa_check bwd > gpurun_out/r02s_bwd.log 2>&1; tail -n 1 gpurun_out/r02s_bwd.log
timeout 300 build/fa_check drop > gpurun_out/r02s_drop.log 2>&1; tail -n 1 gpurun_out/r02s_drop.log
unset LD_LIBRARY_PATH
bash tools/ab.sh "time 512 4096 128 1 10 1" u4 u8 h4 h8 > gpurun_out/r02s_ab.log 2>&1; cat gpurun_out/r02s_ab.log
for v in h4; do LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check trace 512 4096 128 1 0 > gpurun_out/r02s_trace0_$v.log 2>&1; done
cat gpurun_out/r02s_trace0_h4.log
