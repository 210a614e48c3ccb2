cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_pytest.log 2>&1; tail -n 3 gpurun_out/r02f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_cfg3.json 2> gpurun_out/r02f_bench_cfg3.err; head -c 400 gpurun_out/r02f_bench_cfg3.json; echo
bash tools/profile_round.sh r02f > gpurun_out/r02f_profile_round.log 2>&1
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-mt > gpurun_out/r02f_bench_cfg4.json 2> gpurun_out/r02f_bench_cfg4.err; head -c 300 gpurun_out/r02f_bench_cfg4.json; echo
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-mt > gpurun_out/r02f_bench_cfg2.json 2> gpurun_out/r02f_bench_cfg2.err; head -c 300 gpurun_out/r02f_bench_cfg2.json; echo
tools/sustained_probe.sh "bwd x300" build/fa_check time 512 4096 128 1 300 1
