set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_lm_kernels.py tests/test_gpu_mt.py -x -q -m gpu > gpurun_out/pytest_mt_r01g.log 2>&1; tail -5 gpurun_out/pytest_mt_r01g.log
for L in 512; do timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length $L 2>&1 | tail -1 > gpurun_out/mt_r01g_L$L.log; done
cut -c1-400 gpurun_out/mt_r01g_L512.log
timeout 200 python tools/lm_bench.py > gpurun_out/lm_bench_r01g.json 2>&1; tail -c 1500 gpurun_out/lm_bench_r01g.json
