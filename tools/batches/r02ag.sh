cd $GRAFT_REPO_ROOT
for args in "fwd 3 1000 128 1" "bwd 5 700 128 1" "bwd 3 384 64 0" "bwd 40 384 128 1 1 0.1"; do
  echo "== memcheck fa_check $args"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 build/fa_check $args 2>&1 | tail -n 4
done > gpurun_out/r02ag_sanitizer.log 2>&1
cat gpurun_out/r02ag_sanitizer.log
