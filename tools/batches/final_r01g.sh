cd $GRAFT_REPO_ROOT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01g.log 2>&1; tail -1 gpurun_out/smoke_r01g.log | cut -c1-200
timeout 420 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r01g.log 2>&1; tail -1 gpurun_out/pytest_gpu_r01g.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01g_cfg3.json 2> gpurun_out/bench_r01g_cfg3.err; cut -c1-200 gpurun_out/bench_r01g_cfg3.json
for L in 128 512 1024; do timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length $L 2>&1 | tail -1 > gpurun_out/mt_r01g_L$L.log; cut -c1-110 gpurun_out/mt_r01g_L$L.log; done
timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 --attention-precision bf16 2>&1 | tail -1 > gpurun_out/mt_r01g_L512_bf16.log; cut -c1-110 gpurun_out/mt_r01g_L512_bf16.log
timeout 200 python tools/lm_bench.py > gpurun_out/lm_bench_r01g.json 2>&1
