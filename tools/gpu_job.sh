cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x -k "realistic or multiplayer or identity_key" > gpurun_out/r2u_pytest.log 2>&1
tail -12 gpurun_out/r2u_pytest.log
timeout 300 python bench.py --config C5 --beam 2000000 --steps 2 --warmup 1 --no-parity --no-cpu-baseline 2>&1 | cut -c1-330
timeout 300 python bench.py --config C5 --players 3 --beam 2000000 --steps 2 --warmup 1 --no-parity --no-cpu-baseline 2>&1 | cut -c1-330
