cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; tail -n 3 gpurun_out/r02i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench_cfg3.json 2> gpurun_out/r02i_bench_cfg3.err; head -c 300 gpurun_out/r02i_bench_cfg3.json; echo
bash tools/profile_round.sh r02i > gpurun_out/r02i_profile_round.log 2>&1
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-mt > gpurun_out/r02i_bench_cfg4.json 2> gpurun_out/r02i_bench_cfg4.err; head -c 200 gpurun_out/r02i_bench_cfg4.json; echo
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-mt > gpurun_out/r02i_bench_cfg2.json 2> gpurun_out/r02i_bench_cfg2.err; head -c 200 gpurun_out/r02i_bench_cfg2.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -c 400
