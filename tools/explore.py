"""Ad-hoc timing of the level solver on the GPU box (not part of the test suite)."""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import splendor_rl_gym_b200 as S

ap = argparse.ArgumentParser()
ap.add_argument('--goal', type=int, default=15)
ap.add_argument('--beam', type=int, default=300_000)
ap.add_argument('--heuristic', default='aggressive')
ap.add_argument('--bfs', type=int, default=0, help='run pure BFS for this many levels instead')
ap.add_argument('--slots', type=int, default=0)
ap.add_argument('--chunk', type=int, default=0)
ap.add_argument('--reps', type=int, default=2)
ap.add_argument('--nodes', type=int, default=0, help='node-table slots of the grouped level (default 4 per beam slot)')
ap.add_argument('--kv', action='store_true', help='size the key table as round 1 did (for SPL_NO_GROUPED=1 runs)')
a = ap.parse_args()

slots = a.slots or (min(3 << 30, max(1 << 22, int(a.beam * 110 / 0.6))) if (a.kv or a.bfs) else 1 << 22)
nodes = a.nodes or min(int(140e9 / 384), max(1 << 14, a.beam * 4))
eng = S.Engine(0, table_slots=slots, chunk_parents=a.chunk, node_slots=0 if (a.kv or a.bfs) else nodes, max_node_bytes=int(150e9))
k, aux = S.State.newgame().record()
for rep in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.time()
    if a.bfs:
        sol = eng.solver(k, aux, 255, False, 'simple', 0, keep_links=False)
        infos = sol.run(max_levels=a.bfs)
    else:
        sol = eng.solver(k, aux, a.goal, True, a.heuristic, a.beam, 'stable', 'const')
        infos = sol.run()
    torch.cuda.synchronize()
    dt = time.time() - t0
    exp = sum(i['expanded'] for i in infos)
    gen = sum(i['generated'] for i in infos)
    print(f'rep {rep}: levels={len(infos)} expanded={exp} generated={gen} visited={infos[-1]["visited"]} '
          f'wall={dt:.3f}s -> {exp / dt / 1e6:.2f} M expanded/s, {gen / dt / 1e6:.1f} M generated/s  '
          f'mem={torch.cuda.mem_get_info()[0] / 2**30:.1f} GiB free')
    if rep == a.reps - 1:
        print('SUMMARY lib=%s wall=%.1fms expand=%.2f resolve=%.2f select=%.2f sort=%.2f count=%.2f' % (
            __import__('os').environ.get('SPLENDOR_B200_LIB', 'default').split('/')[-1], dt * 1e3,
            *(sum(i['ms_' + k] for i in infos) for k in ('expand', 'resolve', 'select', 'sort', 'count'))))
    if rep == a.reps - 1 and not __import__('os').environ.get('QUIET'):
        for i in infos:
            print('  L%-2d front=%-10d gen=%-11d uniq=%-10d kept=%-9d ms: count %.2f expand %.2f resolve %.2f select %.2f sort %.2f' % (
                i['level'], i['frontier'], i['generated'], i['unique'], i['kept'], i['ms_count'], i['ms_expand'],
                i['ms_resolve'], i['ms_select'], i['ms_sort']))
    sol.close()
