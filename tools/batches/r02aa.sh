cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "persistent or bf16_forward_backward or full_size or cfg4" > gpurun_out/r02aa_pytest.log 2>&1; tail -n 4 gpurun_out/r02aa_pytest.log
