cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes.py -m gpu -q -x -k "not beam_3m" > gpurun_out/r2ai_pytest.log 2>&1
tail -3 gpurun_out/r2ai_pytest.log
SPL_DEBUG=1 QUIET=1 timeout 300 python tools/explore.py --beam 30000000 --reps 2 > gpurun_out/r2ai_debug_30m.log 2>&1; grep -E "rep|SUMMARY" gpurun_out/r2ai_debug_30m.log | tail -3
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/sharded_check.py --grouped --beam 300000 2>&1 | grep -v "^\*\*\|OMP" | tail -2
SPL_TIMING=1 timeout 300 $T tools/sharded_check.py --grouped --beam 60000000 --no-oracle --no-links --reps 2 2>&1 | grep -v "^\*\*\|OMP" | tail -3
