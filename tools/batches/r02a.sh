# round 2, first GPU batch: correctness of the re-scheduled forward / backward + same-box A/B against the round-1 library
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 build/fa_check fwd > gpurun_out/r02a_fwd_check.log 2>&1; tail -3 gpurun_out/r02a_fwd_check.log
timeout 300 build/fa_check bwd > gpurun_out/r02a_bwd_check.log 2>&1; tail -3 gpurun_out/r02a_bwd_check.log
timeout 300 build/fa_check drop > gpurun_out/r02a_drop_check.log 2>&1; tail -3 gpurun_out/r02a_drop_check.log
timeout 300 build/fa_check tf32 > gpurun_out/r02a_tf32_check.log 2>&1; tail -2 gpurun_out/r02a_tf32_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; tail -5 gpurun_out/r02a_pytest.log
bash tools/ab.sh "time 512 4096 128 1 20 0" r01 f_s0p1 f_s1p1 f_s0p2 new f_poly4 f_poly8 > gpurun_out/r02a_ab_fwd.log 2>&1
bash tools/ab.sh "time 512 4096 128 0 10 0" r01 new f_poly4 > gpurun_out/r02a_ab_fwd_nc.log 2>&1
bash tools/ab.sh "time 512 4096 128 1 10 1" r01 new b_ord1 > gpurun_out/r02a_ab_bwd.log 2>&1
bash tools/ab.sh "time 512 4096 128 0 5 1" r01 new b_ord1 > gpurun_out/r02a_ab_bwd_nc.log 2>&1
bash tools/ab.sh "time 128 16384 128 1 3 1 0.1" r01 new > gpurun_out/r02a_ab_cfg4.log 2>&1
cat gpurun_out/r02a_ab_fwd.log gpurun_out/r02a_ab_fwd_nc.log gpurun_out/r02a_ab_bwd.log gpurun_out/r02a_ab_bwd_nc.log gpurun_out/r02a_ab_cfg4.log
timeout 120 build/fa_check tracefwd > gpurun_out/r02a_tracefwd.log 2>&1
timeout 120 build/fa_check trace > gpurun_out/r02a_tracebwd.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.json
