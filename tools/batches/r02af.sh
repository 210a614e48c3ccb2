cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-mt > gpurun_out/r02af_cfg2.json 2> gpurun_out/r02af_cfg2.err; tail -c 600 gpurun_out/r02af_cfg2.json
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-mt > gpurun_out/r02af_cfg4.json 2> gpurun_out/r02af_cfg4.err; tail -c 600 gpurun_out/r02af_cfg4.json
