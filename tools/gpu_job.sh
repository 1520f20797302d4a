cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -k "not beam_3m" > gpurun_out/r2i_pytest.log 2>&1
tail -5 gpurun_out/r2i_pytest.log
SPL_DEBUG=1 QUIET=1 timeout 300 python tools/explore.py --beam 3000000 --reps 2 > gpurun_out/r2i_debug_3m.log 2>&1; grep -E "grouped|rep|SUMMARY" gpurun_out/r2i_debug_3m.log | tail -3
SPL_DEBUG=1 QUIET=1 timeout 300 python tools/explore.py --beam 30000000 --reps 2 > gpurun_out/r2i_debug_30m.log 2>&1; grep -E "grouped|rep|SUMMARY" gpurun_out/r2i_debug_30m.log | tail -5
