set -x
cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 > gpurun_out/mt_r01f_g8_fp32.log 2>&1; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r01f_g8.log 2>&1; echo rc=$?
tail -1 gpurun_out/mt_r01f_g8_fp32.log | cut -c1-200; tail -1 gpurun_out/bench_r01f_g8.log | cut -c1-200
