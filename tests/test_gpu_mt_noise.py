"""GPU parity with the UNMODIFIED, merely seeded reference (noise policy `mt`):
`random.seed(S); reference.solve(...)` vs `random.seed(S); gpu.solve(..., noise='mt')` -- the reference's
own randint(1, 100) stream (Python's global Mersenne Twister), one draw per scored state in next_queue
order, ties by arrival order.  No patched randint and no plug-in heuristic on the reference side."""
import hashlib
import random

import numpy as np
import pytest

import splendor_rl_gym_b200 as S
from common import assert_digest, level_digest

pytestmark = pytest.mark.gpu


def test_speedrun_seeded_reference(golden):
    for run in [r for r in golden['mt_runs'] if r['mode'] == 'speedrun']:
        what = f"mt {run['heuristic']} goal={run['goal']} beam={run['beam']} seed={run['seed']}"
        random.seed(run['seed'])
        stats = []
        sol = S.State.newgame().solve(goal_pts=run['goal'], use_heuristic=True, heuristic_name=run['heuristic'],
                                      beam_width=run['beam'], verbose=False, noise='mt', stats=stats)
        assert [repr(s) for s in sol] == [p['repr'] for p in run['path']], what
        assert [(s.saved, s.pts) for s in sol] == [(p['saved'], p['pts']) for p in run['path']], what
        # the global generator is left exactly where the reference left it
        assert [random.randint(1, 100) for _ in range(4)] == run['after'], what
        for info, want in zip(stats, run['levels']):
            assert info['frontier'] == want['frontier'], what
            if not info['ended']:
                assert (info['generated'], info['unique'], info['kept']) == (want['generated'], want['unique']['n'], want['kept']['n']), what


def test_speedrun_seeded_reference_level_digests(golden):
    eng = S.Engine.get(0)
    run = next(r for r in golden['mt_runs'] if r['mode'] == 'speedrun' and r['heuristic'] == 'aggressive')
    random.seed(run['seed'])
    k, a = S.State.newgame().record()
    sol = eng.solver(k, a, run['goal'], True, run['heuristic'], run['beam'], 'stable', 'mt')
    for want in run['levels']:
        gi = sol.step()
        if gi['ended']:
            break
        fr = sol.frontier().cpu().numpy().view(np.uint64)
        assert_digest(level_digest(fr[:, 0], fr[:, 1], fr[:, 2], fr[:, 3]), want['kept'], f"mt level {want['level']}")
    sol.close()


def test_realistic_seeded_reference(golden):
    for run in [r for r in golden['mt_runs'] if r['mode'] == 'realistic']:
        what = f"mt realistic p={run['players']} goal={run['goal']} beam={run['beam']} seed={run['seed']}"
        P = run['players']
        cfg = S.GameConfig(num_players=P, target_points=run['goal'], gems_per_color={2: 4, 3: 5, 4: 7}[P], infinite_resources=False)
        random.seed(run['seed'])
        sol = S.MultiPlayerState.newgame(cfg, shuffle_market=run['market_seed'] is not None, seed=run['market_seed']).solve(
            beam_width=run['beam'], verbose=False, noise='mt')
        assert len(sol) - 1 == run['plies'] and sol[-1].get_winner() == run['winner'], what
        got = [hashlib.sha256(s.record()['p'][0][:P].tobytes() + s.record()['vis'][0].tobytes() + bytes([s.current_player])).hexdigest()[:16]
               for s in sol]
        assert got == run['path_sha'], what
        assert [dict(pts=p.pts, cards=list(p.cards), saved=p.saved) for p in sol[-1].players] == run['final'], what
        assert [random.randint(1, 100) for _ in range(4)] == run['after'], what
