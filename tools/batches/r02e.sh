set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; tail -5 gpurun_out/r02e_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 3000 gpurun_out/r02e_bench.json; tail -5 gpurun_out/r02e_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; tail -3 gpurun_out/r02e_smoke.log
