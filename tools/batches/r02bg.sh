cd $GRAFT_REPO_ROOT
export LD_LIBRARY_PATH=build/variants/fused
timeout 120 build/fa_check bwd 40 1000 128 1 | tail -n 3
timeout 120 build/fa_check bwd 40 1000 128 0 | tail -n 2
timeout 120 build/fa_check bwd 37 2125 64 1 | tail -n 2
timeout 120 build/fa_check bwd 3 300 128 1 | tail -n 2
timeout 300 build/fa_check bwd | tail -n 1
timeout 300 build/fa_check drop | tail -n 1
for r in 1 2 3; do for f in 0 1; do printf "fused=%s " $f; FA_B200_BWD_FUSED_PASSES=$f timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done; done
for f in 0 1; do printf "fused=%s " $f; FA_B200_BWD_FUSED_PASSES=$f timeout 120 build/fa_check time 64 4096 128 1 20 1 2>&1 | tail -1; done
for f in 0 1; do printf "fused=%s " $f; FA_B200_BWD_FUSED_PASSES=$f timeout 120 build/fa_check time 512 4096 128 0 5 1 2>&1 | tail -1; done
