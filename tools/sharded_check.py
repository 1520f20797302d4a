"""Multi-GPU parity + timing check, run under torchrun on the GPU box:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_check.py --beam 100000
Every rank runs the sharded solver; rank 0 also runs the CPU oracle and compares every level."""
import argparse
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

import splendor_rl_gym_b200 as S
from splendor_rl_gym_b200.sharded import Comm, CudaBackend, GroupedShardedSolver, ShardedSolver

ap = argparse.ArgumentParser()
ap.add_argument('--goal', type=int, default=15)
ap.add_argument('--beam', type=int, default=100_000)
ap.add_argument('--heuristic', default='aggressive')
ap.add_argument('--tie', default='stable')
ap.add_argument('--noise', default='const')
ap.add_argument('--bfs', type=int, default=0)
ap.add_argument('--no-oracle', action='store_true')
ap.add_argument('--reps', type=int, default=1)
ap.add_argument('--block', type=int, default=1 << 20, help='parents per block (round granularity)')
ap.add_argument('--slots', type=int, default=0, help='visited-table slots per rank')
ap.add_argument('--no-links', action='store_true')
ap.add_argument('--grouped', action='store_true', help='queue sharded by card set (GroupedShardedSolver); --block = global ranks per round')
ap.add_argument('--nodes', type=int, default=0, help='node-table slots per rank (grouped)')
ap.add_argument('--identity', default='key', help="'pyhash': visited set keyed by the reference's State.hash (key-sharded driver)")
a = ap.parse_args()

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
world_env = max(1, int(os.environ.get('WORLD_SIZE', '1')))
if a.grouped:
    eng = S.Engine(local, table_slots=1 << 22, node_slots=a.nodes or min(int(140e9 / 384), max(1 << 14, a.beam * 4 // world_env)),
                   max_node_bytes=int(150e9))
else:
    eng = S.Engine(local, table_slots=a.slots or max(1 << 22, int(a.beam * 110 / 0.6 / world_env)), max_table_bytes=int(120e9))
comm = Comm(eng.tdev)
use_h = a.bfs == 0
check = comm.rank == 0 and not a.no_oracle
from splendor_rl_gym_b200 import sharded as _sh0
for rep in range(a.reps):
    if rep > 0:
        if _sh0.TIMING and comm.rank == 0:
            print(f'  phases of rep {rep - 1}: ' + ', '.join(f'{k}={v:.3f}' for k, v in _sh0.PHASES.items()), flush=True)
        _sh0.PHASES.clear()  # the summary line below shows the last (warm) repetition
    if check and rep == 0:
        import oracle
        orc = oracle.Solver(255 if a.bfs else a.goal, use_heuristic=use_h, heuristic_name=a.heuristic, beam_width=a.beam,
                            policy=a.tie, noise=a.noise, identity=a.identity)
    torch.cuda.synchronize()
    t0 = time.time()
    if a.grouped:
        sol = GroupedShardedSolver(eng, comm, 0, 0, a.goal, a.heuristic, a.beam, a.noise, round_parents=max(a.block, 1024) if a.block != 1 << 20 else 1 << 27,
                                   keep_links=not a.no_links)
    else:
        sol = ShardedSolver(CudaBackend(eng), comm, 0, 0, 255 if a.bfs else a.goal, use_h, a.heuristic, a.beam, a.tie, a.noise,
                            block_parents=a.block, keep_links=not a.no_links, identity=a.identity)
    while True:
        gi = sol.step()
        if rep == 0 and not a.no_oracle:
            fr = sol.gather_frontier() if not gi['ended'] else None
            if check:
                oi = orc.step()
                fields = ('frontier', 'goal_rank') if gi['ended'] else ('frontier', 'generated', 'unique', 'kept', 'goal_rank', 'visited')
                for f in fields:
                    assert gi[f] == oi[f], (f, gi, oi)
                if not gi['ended']:
                    h = fr.cpu().numpy().view(np.uint64)
                    st, lk = orc.level(oi['level'] + 1)
                    assert (h[:, 0] == st['lo']).all() and (h[:, 1] == st['hi']).all() and (h[:, 2] == st['aux']).all() and (h[:, 3] == lk).all()
        if a.bfs and comm.rank == 0:
            torch.cuda.synchronize()
            print(f"  level {gi['level']}: frontier={gi['frontier']} generated={gi['generated']} unique={gi['unique']} visited={gi.get('visited')} "
                  f"t={time.time() - t0:.2f}s free={torch.cuda.mem_get_info()[0] / 2**30:.0f} GiB", flush=True)
        if gi['ended'] or (a.bfs and len(sol.infos) >= a.bfs):
            break
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rep == 0 and not a.no_oracle and not a.no_links and not a.bfs:
        ranks, ords = sol.path()
        if check:
            assert len(ords) == orc.nlevels - 1, (len(ords), orc.nlevels)
    if a.grouped:
        sol.close()
    if comm.rank == 0:
        exp = sum(i['expanded'] for i in sol.infos)
        print(f'world={comm.world} rep={rep} levels={len(sol.infos)} expanded={exp} wall={dt:.3f}s -> {exp / dt / 1e6:.2f} M expanded/s'
              + (' PARITY OK vs oracle' if check and rep == 0 else ''), flush=True)
from splendor_rl_gym_b200 import sharded as _sh
if _sh.TIMING and comm.rank == 0:
    tot = sum(_sh.PHASES.values())
    if a.grouped:
        print('rank-0 stage ms (last rep): ' + ', '.join(f"{k}={sum(i.get('ms_' + k, 0.0) for i in sol.infos):.1f}" for k in ('sort', 'thread', 'warp', 'cta')))
    print('phases (s, last rep): ' + ', '.join(f'{k}={v:.3f} ({100 * v / tot:.0f}%)' for k, v in _sh.PHASES.items()))
if comm.on:
    dist.destroy_process_group()
