cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SPL_TIMING=1 timeout 300 $TR tools/sharded_check.py --grouped --beam 60000000 --no-oracle --no-links --reps 3 > gpurun_out/r2q_time60m.log 2>&1; grep -v "^\*\*\|OMP" gpurun_out/r2q_time60m.log | tail -6
timeout 300 $TR tools/sharded_check.py --grouped --beam 60000000 --no-oracle --no-links --reps 3 > gpurun_out/r2q_time60m_nt.log 2>&1; grep -v "^\*\*\|OMP" gpurun_out/r2q_time60m_nt.log | tail -4
