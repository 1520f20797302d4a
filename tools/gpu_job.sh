cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2as_pytest.log 2>&1
tail -2 gpurun_out/r2as_pytest.log
SPL_DEBUG=1 QUIET=1 timeout 300 python tools/explore.py --beam 30000000 --reps 2 > gpurun_out/r2as_debug_30m.log 2>&1; grep -E "SUMMARY" gpurun_out/r2as_debug_30m.log | tail -1
grep -E "grouped\] L1[34]" gpurun_out/r2as_debug_30m.log | tail -4 | cut -c150-400
