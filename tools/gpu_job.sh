cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; tail -c 2500 gpurun_out/r2s_bench_n2.json; grep -v "^\*\*\|OMP" gpurun_out/r2s_bench_n2.err | tail -5
timeout 300 python bench.py --steps 3 --warmup 2 --no-parity --no-cpu-baseline > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_n1.json')); print('N1 value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds_per_solve'])"
