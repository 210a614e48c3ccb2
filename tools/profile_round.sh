set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
build/fa_check time 512 4096 128 1 2 1 > gpurun_out/plain_time.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_r01b_bwd build/fa_check time 512 4096 128 1 2 1 > gpurun_out/ncu_bwd.log 2>&1
build/fa_check time 512 4096 128 1 2 1 > gpurun_out/plain_time.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_r01b_fwd build/fa_check time 512 4096 128 1 2 1 > gpurun_out/ncu_fwd.log 2>&1
build/fa_check time32 512 1024 64 0 2 > gpurun_out/plain_time32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_r01b_tf32 build/fa_check time32 512 1024 64 0 2 > gpurun_out/ncu_tf32.log 2>&1
python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3.log 2>&1
tail -c 600 gpurun_out/bench_cfg2.log gpurun_out/bench_cfg4.log gpurun_out/bench_cfg3.log
ls -la gpurun_out | tail -20
