"""`random.seed(S); solve(..., noise='mt')` under torchrun must return the line of the seeded, unmodified reference
(tests/golden/mt_runs.json) on any number of GPUs:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mt_multi_check.py"""
import json
import os
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

import splendor_rl_gym_b200 as S

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
runs = [r for r in json.load(open(ROOT / 'tests/golden/mt_runs.json')) if r['mode'] == 'speedrun']
for run in runs:
    random.seed(run['seed'])
    sol = S.State.newgame().solve(goal_pts=run['goal'], use_heuristic=True, heuristic_name=run['heuristic'], beam_width=run['beam'],
                                  verbose=False, noise='mt', device=local)
    assert [repr(s) for s in sol] == [p['repr'] for p in run['path']], run['heuristic']
    assert [(s.saved, s.pts) for s in sol] == [(p['saved'], p['pts']) for p in run['path']]
    assert [random.randint(1, 100) for _ in range(4)] == run['after']
    if int(os.environ.get('RANK', '0')) == 0:
        print(f"mt {run['heuristic']} goal={run['goal']} beam={run['beam']} seed={run['seed']}: line of the seeded reference reproduced "
              f"on {os.environ.get('WORLD_SIZE', '1')} GPU(s)", flush=True)
if dist.is_initialized():
    dist.destroy_process_group()
