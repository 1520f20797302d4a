"""Parity at BASELINE.json widths (the sizes the bench and the reference's defaults actually use).

CPU half (`-m "not gpu"`): the C oracle against SURVEY.md 8(c)'s known answers of the UNMODIFIED
reference at beam 20 000 / 2 000 (realistic mode, src/solver.py:750-860) -- this pins the oracle at
the width configs[4] names, not only at the fixture widths.
GPU half (`-m gpu`): the CUDA path through the C ABI against the oracle's full arrays, level by level,
at configs[0] (goal 10, simple, beam 300 000), configs[2] (goal 15, aggressive, beam 3 000 000) and
configs[4] (realistic 2 and 3 players, goal 15, market seed 0, beam 20 000).
"""
import numpy as np
import pytest

import oracle

# SURVEY.md 8(c), realistic-mode table (reference run with randint -> 50, ties by arrival order)
SURVEY_REALISTIC = {
    (2, 15, 20_000): dict(plies=53, winner=0, final=[(16, 13, 37), (15, 10, 26)], expanded=977_184,
                          generated=5_661_529, visited=4_194_569, smin=-5534.680148893811, smax=5587.378147681974),
    (3, 15, 2_000): dict(plies=84, winner=2, final=[(15, 16, 55), (13, 15, 46), (16, 14, 45)], expanded=161_768,
                         generated=1_207_994, visited=1_021_114, smin=-5592.961739816975, smax=12956.389628148994),
}


def _seed0_market(golden):
    return next(r for r in golden['realistic_runs'] if r['seed'] == 0 and r['players'] == 2)['market']


def _ocfg(golden, players):
    mk = _seed0_market(golden)
    return oracle.make_rconfig(players, 15, {2: 4, 3: 5, 4: 7}[players], [mk['t1'], mk['t2'], mk['t3']])


def _ncards(rec, q):
    return bin(int(rec['p'][q]['mlo'])).count('1') + bin(int(rec['p'][q]['mhi'])).count('1')


def _check_survey(orc, infos, players, want):
    path = orc.path()
    fin = path[-1]
    assert len(path) - 1 == want['plies']
    assert [(oracle.r_pts(fin, q), _ncards(fin, q), int(fin['p'][q]['saved'])) for q in range(players)] == want['final']
    pts = [oracle.r_pts(fin, q) for q in range(players)]
    assert pts.index(max(pts)) == want['winner']
    assert sum(i['expanded'] for i in infos) == want['expanded']
    assert sum(i['generated'] for i in infos) == want['generated']
    assert infos[-1]['visited'] == want['visited']
    assert orc.score_range() == (want['smin'], want['smax'])


@pytest.mark.parametrize('players,beam', [(2, 20_000), (3, 2_000)])
def test_oracle_realistic_survey_known_answers(golden, players, beam):
    orc = oracle.RSolver(_ocfg(golden, players), beam)
    infos = orc.run()
    _check_survey(orc, infos, players, SURVEY_REALISTIC[(players, 15, beam)])
    orc.close()


# ------------------------------------------------------------------ GPU half
@pytest.fixture(scope='module')
def eng():
    import splendor_rl_gym_b200 as S
    return S.Engine.get(0)


@pytest.mark.gpu
@pytest.mark.parametrize('players', [2, 3])
def test_gpu_realistic_beam_20000_vs_oracle(eng, golden, players):
    """configs[4] at the reference's default width: every ply's queue (96-byte records incl. parent links) equals
    the oracle's, which in turn reproduces the unmodified reference's known answers (2 players: SURVEY.md 8c)."""
    from splendor_rl_gym_b200.realistic import RConfig
    mk = _seed0_market(golden)
    gpc = {2: 4, 3: 5}[players]
    cfg = RConfig(players, 15, gpc, 0)
    for t, seq in enumerate((mk['t1'], mk['t2'], mk['t3'])):
        cfg.deck_len[t] = len(seq)
        for i, c in enumerate(seq):
            cfg.deck[t][i] = c
    ocfg = _ocfg(golden, players)
    sol = eng.rsolver(cfg, oracle.r_root(ocfg), 20_000)
    orc = oracle.RSolver(ocfg, 20_000)
    while True:
        gi, oi = sol.step(), orc.step()
        assert (gi['frontier'], gi['goal_rank']) == (oi['frontier'], oi['goal_rank']), gi
        if gi['ended']:
            break
        assert (gi['generated'], gi['unique'], gi['kept'], gi['visited']) == (oi['generated'], oi['unique'], oi['kept'], oi['visited']), gi
        assert sol.frontier().tobytes() == orc.level(oi['level'] + 1).tobytes(), gi['level']
    assert orc.done
    if players == 2:
        _check_survey(orc, orc.infos, 2, SURVEY_REALISTIC[(2, 15, 20_000)])
    _, ords = sol.path()
    assert len(ords) == orc.nlevels - 1
    sol.close()
    orc.close()


def _speedrun_vs_oracle(eng, goal, hname, beam, tie='stable', noise='const'):
    import splendor_rl_gym_b200 as S
    k, a = S.State.newgame().record()
    sol = eng.solver(k, a, goal, True, hname, beam, tie, noise)
    orc = oracle.Solver(goal, use_heuristic=True, heuristic_name=hname, beam_width=beam, policy=tie, noise=noise)
    while True:
        gi, oi = sol.step(), orc.step()
        what = f'{hname} goal {goal} beam {beam} level {gi["level"]}'
        for f in ('frontier', 'generated', 'unique', 'kept', 'goal_rank', 'visited'):
            assert gi[f] == oi[f], (what, f, gi, oi)
        if gi['ended']:
            assert orc.done
            break
        fr = sol.frontier().cpu().numpy().view(np.uint64)
        st, lk = orc.level(oi['level'] + 1)
        assert (fr[:, 0] == st['lo']).all() and (fr[:, 1] == st['hi']).all() and (fr[:, 2] == st['aux']).all(), what
        assert (fr[:, 3] == lk).all(), what
    _, ords = sol.path()
    assert len(ords) == orc.nlevels - 1
    sol.close()
    orc.close()


@pytest.mark.gpu
def test_gpu_config0_goal10_simple_beam_300k_vs_oracle(eng):
    """BASELINE configs[0] exactly: `splendor_fastest_win.py 10 -u` (simple heuristic, beam 300 000)."""
    _speedrun_vs_oracle(eng, 10, 'simple', 300_000)


@pytest.mark.gpu
def test_gpu_config2_goal15_aggressive_beam_3m_vs_oracle(eng):
    """BASELINE configs[2] at ten times the reference's default width: full arrays of every level vs the oracle
    (about a minute of host time for the oracle's side)."""
    _speedrun_vs_oracle(eng, 15, 'aggressive', 3_000_000)
