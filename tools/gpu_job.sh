cd $GRAFT_REPO_ROOT
SPL_NO_GROUPED=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes.py tests/test_gpu_properties.py -m gpu -q -x -k "not beam_3m and not growth" > gpurun_out/r2k_pytest.log 2>&1
tail -3 gpurun_out/r2k_pytest.log
SPL_NO_GROUPED=1 SPL_NO_ORDER=1 QUIET=1 timeout 300 python tools/explore.py --kv --beam 30000000 --reps 2 > gpurun_out/r2k_kv_noorder.log 2>&1; grep -E "rep|SUMMARY" gpurun_out/r2k_kv_noorder.log | tail -2
SPL_NO_GROUPED=1 timeout 300 python tools/explore.py --kv --beam 30000000 --reps 2 > gpurun_out/r2k_kv_order.log 2>&1; grep -E "rep |SUMMARY| L1[0-4]" gpurun_out/r2k_kv_order.log | tail -8
