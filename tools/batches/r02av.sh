cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check drop > gpurun_out/r02av_drop.log 2>&1; tail -n 1 gpurun_out/r02av_drop.log
for r in 1 2; do for v in old new; do printf "%s cfg4 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -1; done; done
for v in old new; do printf "%s T4096 drop " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 5 1 0.1 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dropout or cfg4 or persistent" 2>&1 | tail -n 2
