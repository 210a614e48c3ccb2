set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_lm_kernels.py tests/test_gpu_mt.py -x -q -m gpu > gpurun_out/pytest_mt_r01g.log 2>&1; tail -5 gpurun_out/pytest_mt_r01g.log
for i in 1 2; do timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 2>&1 | tail -1 > gpurun_out/mt_r01g_L512.log; cut -c1-330 gpurun_out/mt_r01g_L512.log; done
FA_MT_FUSED_BIAS_GRAD=0 timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 2>&1 | tail -1 | cut -c1-120
timeout 200 python tools/lm_bench.py > gpurun_out/lm_bench_r01g.json 2>&1; grep -o '"kernel": "gemm[^}]*' gpurun_out/lm_bench_r01g.json | cut -c1-90
