cd $GRAFT_REPO_ROOT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 8 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench_n8.json') if l.startswith('{')][-1])
print('value %.4e'%d['value'], d['ms_per_step'], d['stage_ms_per_step'], 'frac', d['roofline']['frac'], 'e2e %.4e'%d['e2e']['value'], d['e2e']['seconds_each_solve_rank0'], d['parity_check']['ok'])
PY
