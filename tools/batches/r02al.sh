cd $GRAFT_REPO_ROOT
LD_LIBRARY_PATH=build/variants/ds1 timeout 300 build/fa_check bwd > gpurun_out/r02al_bwd_ds1.log 2>&1; tail -n 1 gpurun_out/r02al_bwd_ds1.log
grep 'T= 1000 d=128 causal=1' gpurun_out/r02al_bwd_ds1.log | tail -n 1
LD_LIBRARY_PATH=build/variants/ds0 timeout 300 build/fa_check bwd 3 1000 128 1 | tail -n 1
bash tools/ab.sh "time 512 4096 128 1 10 1" ds0 ds1 > gpurun_out/r02al_ab.log 2>&1; cat gpurun_out/r02al_ab.log
