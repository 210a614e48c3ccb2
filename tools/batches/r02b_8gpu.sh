set -x
cd $GRAFT_REPO_ROOT
for n in 8 1; do
if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02b_bench_g1.json 2> gpurun_out/r02b_bench_g1.err; else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02b_bench_g$n.json 2> gpurun_out/r02b_bench_g$n.err; fi
tail -c 300 gpurun_out/r02b_bench_g$n.json; tail -2 gpurun_out/r02b_bench_g$n.err
done
