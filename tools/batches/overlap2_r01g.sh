cd $GRAFT_REPO_ROOT
for ov in True False; do
timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$( [ $ov = True ] && echo 1 || echo 2 ) flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 --overlap-allreduce $ov > gpurun_out/mt_r01g_g2_overlap_$ov.log 2>&1; echo "overlap=$ov rc=$?"
tail -1 gpurun_out/mt_r01g_g2_overlap_$ov.log | cut -c1-420
done
