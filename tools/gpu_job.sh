cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/sharded_check.py --grouped --beam 300000 > gpurun_out/r2aj_p2p.log 2>&1
grep -v "^\*\*\|OMP" gpurun_out/r2aj_p2p.log | grep -i "error\|Traceback\|File\|spl_\|cuda\|world=" | head -20
timeout 300 $T tools/sharded_check.py --grouped --beam 20000 --block 3000 2>&1 | grep -v "^\*\*\|OMP" | tail -2
SPL_TIMING=1 timeout 300 $T tools/sharded_check.py --grouped --beam 60000000 --no-oracle --no-links --reps 3 2>&1 | grep -v "^\*\*\|OMP" | tail -4
