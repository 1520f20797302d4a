cd $GRAFT_REPO_ROOT
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "grouped or spill or overflow" 2>&1 | tail -1
