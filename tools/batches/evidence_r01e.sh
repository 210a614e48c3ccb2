set -x
cd $GRAFT_REPO_ROOT
timeout 200 build/umma_rate > gpurun_out/umma_rate_r01e.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e_cfg3.log 2>&1
python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_r01e_cfg2.log 2>&1
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_r01e_cfg4.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01e_ref.log 2>&1
python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch_r01e.log 2>&1
for L in 512 128; do python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length $L 2>&1 | tail -1 > gpurun_out/mt_r01e_L$L.log; done
python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 --attention-precision bf16 2>&1 | tail -1 > gpurun_out/mt_r01e_L512_bf16.log
python tools/lm_bench.py > gpurun_out/lm_bench_r01e.json 2>&1
tail -c 400 gpurun_out/bench_r01e_cfg3.log gpurun_out/bench_r01e_cfg2.log gpurun_out/bench_r01e_cfg4.log gpurun_out/mt_r01e_L512.log
