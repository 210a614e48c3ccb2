cd $GRAFT_REPO_ROOT
for BH in 64 128 512; do for r in 1 2; do for v in g8 g16 g32; do printf "%s BH=%s " $v $BH; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time $BH 4096 128 1 20 1 2>&1 | tail -2 | tr '\n' ' '; echo; done; done; done
