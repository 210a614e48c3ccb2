set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-mt > gpurun_out/r02h_bench_g1_same_box.json 2> gpurun_out/r02h_bench_g1.err; head -c 300 gpurun_out/r02h_bench_g1_same_box.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02h_bench_g8.json 2> gpurun_out/r02h_bench_g8.err; head -c 300 gpurun_out/r02h_bench_g8.json; echo; tail -3 gpurun_out/r02h_bench_g8.err
