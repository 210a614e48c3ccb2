cd $GRAFT_REPO_ROOT
for r in 1 2; do for v in c_u2 c_u4 c_u8; do printf "%s cfg4 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 128 16384 128 1 3 1 0.1 2>&1 | tail -1; done; done
for v in c_u4 c_u8; do printf "%s cfg3 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 128 1 10 1 2>&1 | tail -1; done
for v in c_u4 c_u8; do printf "%s d64 " $v; LD_LIBRARY_PATH=build/variants/$v timeout 120 build/fa_check time 512 4096 64 1 10 1 2>&1 | tail -1; done
