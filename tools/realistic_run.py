"""Timing of realistic mode (BASELINE configs[4]) on one GPU: python tools/realistic_run.py --players 2 --beam 2000000"""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import splendor_rl_gym_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument('--players', type=int, default=2)
ap.add_argument('--goal', type=int, default=15)
ap.add_argument('--beam', type=int, default=20_000)
ap.add_argument('--seed', type=int, default=0)
ap.add_argument('--reps', type=int, default=2)
a = ap.parse_args()
gpc = {2: 4, 3: 5, 4: 7}[a.players]
cfg = S.GameConfig(num_players=a.players, target_points=a.goal, gems_per_color=gpc, infinite_resources=False)
eng = S.Engine(0, table_slots=max(1 << 22, int(a.beam * 400)))
for rep in range(a.reps):
    st = []
    torch.cuda.synchronize()
    t0 = time.time()
    sol = S.MultiPlayerState.newgame(cfg, shuffle_market=True, seed=a.seed).solve(beam_width=a.beam, verbose=False, engine=eng, stats=st)
    torch.cuda.synchronize()
    dt = time.time() - t0
    exp = sum(i['expanded'] for i in st)
    gen = sum(i['generated'] for i in st)
    print(f'players={a.players} beam={a.beam} rep={rep}: plies={len(sol) - 1} winner={sol[-1].get_winner()} '
          f'pts={[p.pts for p in sol[-1].players]} expanded={exp} generated={gen} visited={st[-1]["visited"]} '
          f'wall={dt:.3f}s -> {exp / dt / 1e6:.2f} M expanded/s', flush=True)
