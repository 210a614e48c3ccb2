set -x
cd $GRAFT_REPO_ROOT
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01f.log 2>&1; tail -1 gpurun_out/smoke_r01f.log
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r01f.log 2>&1; tail -1 gpurun_out/pytest_gpu_r01f.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01f_cfg3.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01f_ref.log 2>&1
build/fa_check bwd32 512 1024 1 > gpurun_out/bwd32_time.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fa_bwd_tf32_kernel -s 1 -c 1 -o gpurun_out/prof_r01f_bwd_tf32 build/fa_check bwd32 512 1024 1 > gpurun_out/ncu_bwd_tf32_r01f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mt_launches_r01f2.csv python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 --iters 4 --cuda-graph False > gpurun_out/mt_ncu_r01f2.log 2>&1
for L in 128 512 1024; do python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length $L 2>&1 | tail -1 > gpurun_out/mt_r01f_L$L.log; done
python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 --attention-precision bf16 2>&1 | tail -1 > gpurun_out/mt_r01f_L512_bf16.log
python tools/lm_bench.py > gpurun_out/lm_bench_r01f.json 2>&1
tail -c 300 gpurun_out/bench_r01f_cfg3.log; cat gpurun_out/mt_r01f_L512.log | cut -c1-150
