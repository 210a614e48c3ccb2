set -x
cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check bwd32 > gpurun_out/r02h_bwd32.log 2>&1; grep -c OK gpurun_out/r02h_bwd32.log; grep -v OK gpurun_out/r02h_bwd32.log | tail -12
