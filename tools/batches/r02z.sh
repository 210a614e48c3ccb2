cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02z_pytest.log 2>&1; tail -n 4 gpurun_out/r02z_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; tail -c 1500 gpurun_out/r02z_bench.json
