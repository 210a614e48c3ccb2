cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; tail -n 2 gpurun_out/r02k_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench_cfg3.json 2> gpurun_out/r02k_bench_cfg3.err; head -c 260 gpurun_out/r02k_bench_cfg3.json; echo
timeout 200 python tools/fuzz_persistent.py 40 5 2>&1 | tail -1
