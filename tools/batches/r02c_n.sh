cd $GRAFT_REPO_ROOT
n=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2981$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02d_bench_g$n.json 2> gpurun_out/r02d_bench_g$n.err; tail -c 200 gpurun_out/r02d_bench_g$n.json
