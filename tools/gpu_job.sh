cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grouped_sharded" > gpurun_out/r2l_pytest.log 2>&1
tail -15 gpurun_out/r2l_pytest.log
