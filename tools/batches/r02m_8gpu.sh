set -x
cd $GRAFT_REPO_ROOT
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02m_bench_g$n.json 2> gpurun_out/r02m_bench_g$n.err; tail -c 600 gpurun_out/r02m_bench_g$n.json; tail -3 gpurun_out/r02m_bench_g$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02m_ref_g8.json 2>&1; tail -c 400 gpurun_out/r02m_ref_g8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 flash-attention-minitorch_b200/project/run_machine_translation.py --iters 42 --overlap-allreduce False > gpurun_out/r02m_mt_g8_nooverlap.log 2>&1; tail -1 gpurun_out/r02m_mt_g8_nooverlap.log
