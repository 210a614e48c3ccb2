cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check drop | tail -n 1
for r in 1 2; do for v in old new; do printf "%s cfg4 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -2 | tr '\n' ' '; echo; done; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dropout or cfg4" 2>&1 | tail -n 2
