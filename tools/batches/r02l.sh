set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest.log 2>&1; tail -4 gpurun_out/r02l_pytest.log
bash tools/profile_round.sh r02a
