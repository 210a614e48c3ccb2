set -x
cd $GRAFT_REPO_ROOT
timeout 120 build/tma_mix > gpurun_out/r02b_tma_mix.log 2>&1; cat gpurun_out/r02b_tma_mix.log
timeout 120 build/fa_check trace > gpurun_out/r02b_tracebwd.log 2>&1; cat gpurun_out/r02b_tracebwd.log
for m in none sk st; do
  for r in 1 2; do
    printf "b_dbg %s " $m; FA_B200_BWD_DEBUG_DQ=$m LD_LIBRARY_PATH=build/variants/b_dbg timeout 120 build/fa_check time 512 4096 128 1 10 1 | tail -1
  done
done > gpurun_out/r02b_dbg.log 2>&1; cat gpurun_out/r02b_dbg.log
