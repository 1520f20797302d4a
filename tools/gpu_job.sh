cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/sharded_check.py --grouped --beam 300000 2>&1 | grep "world="
SPL_TIMING=1 timeout 300 $T tools/sharded_check.py --grouped --beam 60000000 --no-oracle --no-links --reps 2 2>&1 | grep -v "^\*\*\|OMP" | tail -3
SPL_NO_P2P=1 SPL_TIMING=1 timeout 300 $T tools/sharded_check.py --grouped --beam 60000000 --no-oracle --no-links --reps 2 2>&1 | grep -v "^\*\*\|OMP" | tail -1
