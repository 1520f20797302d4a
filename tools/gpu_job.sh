cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err; tail -c 2600 gpurun_out/r2t_bench_n8.json; grep -v "^\*\*\|OMP" gpurun_out/r2t_bench_n8.err | tail -5
SPL_TIMING=1 timeout 300 $TR tools/sharded_check.py --grouped --beam 240000000 --no-oracle --no-links --reps 2 > gpurun_out/r2t_phases_n8.log 2>&1; grep -v "^\*\*\|OMP" gpurun_out/r2t_phases_n8.log | tail -4
