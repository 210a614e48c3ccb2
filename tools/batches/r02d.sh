set -x
cd $GRAFT_REPO_ROOT
bash tools/ab.sh "time 512 4096 128 1 10 1" r01 new b_ord2 > gpurun_out/r02d_ab_bwd.log 2>&1; cat gpurun_out/r02d_ab_bwd.log
timeout 100 build/fa_check bwd 3 1000 128 1 > gpurun_out/r02d_chk.log 2>&1; LD_LIBRARY_PATH=build/variants/b_ord2 timeout 100 build/fa_check bwd 3 1000 128 1 >> gpurun_out/r02d_chk.log 2>&1; cat gpurun_out/r02d_chk.log
