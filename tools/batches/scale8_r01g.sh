cd $GRAFT_REPO_ROOT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 > gpurun_out/mt_r01g_g8_fp32.log 2>&1; echo rc=$?
tail -1 gpurun_out/mt_r01g_g8_fp32.log | cut -c1-400
