cd $GRAFT_REPO_ROOT
SPL_TIMING=1 timeout 100 python tools/sharded_check.py --grouped --beam 30000000 --no-oracle --no-links --reps 2 2>&1 | grep -v "^\*\*\|OMP" | tail -3 | tee gpurun_out/r2i_world1.log
