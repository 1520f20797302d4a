cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_pytest_full.log 2>&1; tail -2 gpurun_out/r2i_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench_n1.json') if l.startswith('{')][-1])
print('value %.4e'%d['value'], d['ms_per_step'], d['stage_ms_per_step'], 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'e2e %.4e'%d['e2e']['value'], d['e2e']['seconds_each_solve_rank0'], d['parity_check']['ok'])
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -c 400
