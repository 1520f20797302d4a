cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --no-cpu-baseline --no-parity > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2g_bench_n2.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e'])
PY
