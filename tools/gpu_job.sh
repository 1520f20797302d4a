cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -q -x 2>&1 | tail -2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench_n2.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['seconds_each_solve_rank0'], d['e2e']['value'])
PY
