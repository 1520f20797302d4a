cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes.py -m gpu -q -x -k "not beam_3m" > gpurun_out/r2ab_pytest.log 2>&1
tail -3 gpurun_out/r2ab_pytest.log
SPL_DEBUG=1 QUIET=1 timeout 300 python tools/explore.py --beam 30000000 --reps 2 > gpurun_out/r2ab_debug_30m.log 2>&1; grep -E "rep|SUMMARY" gpurun_out/r2ab_debug_30m.log | tail -3
QUIET=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"psort_scatter_kernel|sort_hist_kernel|m2_runs_kernel" -s 90 -c 7 -o gpurun_out/r2ab_sort -f python tools/explore.py --beam 30000000 --reps 1 > gpurun_out/r2ab_ncu.log 2>&1
tail -2 gpurun_out/r2ab_ncu.log; ls -la gpurun_out/r2ab_sort.ncu-rep
