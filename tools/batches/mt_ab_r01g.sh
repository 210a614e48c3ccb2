cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_lm_kernels.py tests/test_gpu_mt.py -x -q -m gpu > gpurun_out/pytest_mt_r01g.log 2>&1; tail -3 gpurun_out/pytest_mt_r01g.log
run() { timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 2>&1 | tail -1 | cut -c60-90; }
echo "both fused"; run; run
echo "logits bias off"; FA_MT_FUSED_LOGITS_BIAS_GRAD=0 run; FA_MT_FUSED_LOGITS_BIAS_GRAD=0 run
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mt_launches_r01g2.csv python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 --iters 4 --cuda-graph False > gpurun_out/mt_ncu_r01g.log 2>&1
