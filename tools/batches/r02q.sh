cd $GRAFT_REPO_ROOT
timeout 120 build/fa_check trace 512 4096 128 1 20 > gpurun_out/r02q_trace20.log 2>&1; cat gpurun_out/r02q_trace20.log
