set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_bench_g2.json 2> gpurun_out/r02g_bench_g2.err; tail -c 2500 gpurun_out/r02g_bench_g2.json; tail -5 gpurun_out/r02g_bench_g2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 flash-attention-minitorch_b200/project/run_machine_translation.py --iters 30 > gpurun_out/r02g_mt_g2.log 2>&1; tail -4 gpurun_out/r02g_mt_g2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 flash-attention-minitorch_b200/project/run_machine_translation.py --iters 30 --overlap-allreduce False > gpurun_out/r02g_mt_g2_nooverlap.log 2>&1; tail -2 gpurun_out/r02g_mt_g2_nooverlap.log
timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --iters 30 > gpurun_out/r02g_mt_g1.log 2>&1; tail -2 gpurun_out/r02g_mt_g1.log
