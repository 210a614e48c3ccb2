set -x
cd $GRAFT_REPO_ROOT
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mt_launches_r01g.csv python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 --iters 4 --cuda-graph False > gpurun_out/mt_ncu_r01g.log 2>&1
for i in 1 2; do timeout 200 python flash-attention-minitorch_b200/project/run_machine_translation.py --use-flash-attention True --model-max-length 512 2>&1 | tail -1 | cut -c1-120; done
