# tools/profile_round.sh <tag>   (run under gpurun, ONE GPU): the evidence set of a round.
#   1. plain bench run, then the ncu launch list of the same command (shares of a step)
#   2. ncu --set full of the two headline kernels (backward main kernel, forward) at the config-3 shapes, each after a plain run
# Reports land in gpurun_out/; summarise them here (no GPU) with tools/profile_summarise.sh <tag>.
tag=${1:-r02}
set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 1 --no-mt > gpurun_out/plain_bench_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-mt > gpurun_out/ncu_launch_$tag.log 2>&1
build/fa_check time 512 4096 128 1 2 1 > gpurun_out/plain_time_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_${tag}_bwd build/fa_check time 512 4096 128 1 2 1 > gpurun_out/ncu_bwd_$tag.log 2>&1
build/fa_check time 512 4096 128 1 2 1 > gpurun_out/plain_time_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_kernel -s 1 -c 1 -o gpurun_out/prof_${tag}_fwd build/fa_check time 512 4096 128 1 2 1 > gpurun_out/ncu_fwd_$tag.log 2>&1
ls -la gpurun_out | tail -8
