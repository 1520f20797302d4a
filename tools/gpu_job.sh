cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes.py -m gpu -q -x -k "not beam_3m" > gpurun_out/r2ah_pytest.log 2>&1
tail -3 gpurun_out/r2ah_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/sharded_check.py --beam 50000 --identity pyhash 2>&1 | grep -v "^\*\*\|OMP" | tail -2
timeout 300 $T tools/sharded_check.py --bfs 7 --identity pyhash --block 5000 2>&1 | grep -v "^\*\*\|OMP" | tail -2
timeout 300 $T tools/sharded_check.py --grouped --beam 300000 2>&1 | grep -v "^\*\*\|OMP" | tail -2
SPL_TIMING=1 timeout 300 $T tools/sharded_check.py --grouped --beam 60000000 --no-oracle --no-links --reps 2 2>&1 | grep -v "^\*\*\|OMP" | tail -3
