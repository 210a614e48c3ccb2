cd $GRAFT_REPO_ROOT
timeout 300 build/fa_check fwd | tail -n 1
for BH in 512 64 128; do printf "BH=%s " $BH; timeout 120 build/fa_check time $BH 4096 128 1 20 0 2>&1 | tail -1; done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 2
