cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "overflow or spill or grouped" 2>&1 | tail -3
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --config C4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2g_bench_c4_n2.json 2> gpurun_out/r2g_bench_c4_n2.err; tail -c 400 gpurun_out/r2g_bench_c4_n2.json; grep -v "^\*\*\|OMP\|^$" gpurun_out/r2g_bench_c4_n2.err | grep -A8 Traceback | head -20
