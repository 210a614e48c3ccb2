cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_pytest.log 2>&1; tail -n 3 gpurun_out/r02e_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench_cfg3.json 2> gpurun_out/r02e_bench_cfg3.err; tail -c 2500 gpurun_out/r02e_bench_cfg3.json
bash tools/profile_round.sh r02e > gpurun_out/r02e_profile_round.log 2>&1
tools/sustained_probe.sh "bwd x300" build/fa_check time 512 4096 128 1 300 1
tools/sustained_probe.sh "fwd x1000" build/fa_check time 512 4096 128 1 1000 0
