cd $GRAFT_REPO_ROOT
bash tools/ab.sh "time 512 4096 128 1 20 0" p0 p8 p4 p2 p1 > gpurun_out/r02aj_ab.log 2>&1; cat gpurun_out/r02aj_ab.log
LD_LIBRARY_PATH=build/variants/p2 timeout 200 build/fa_check fwd | tail -n 1
