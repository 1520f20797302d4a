"""GPU smoke tests of the command line front-end (the reference's CLI flags, src: splendor_fastest_win.py:14-148)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(*args):
    r = subprocess.run([sys.executable, str(ROOT / 'splendor_fastest_win.py'), *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_cli_speedrun_bfs_goal3():
    """reference CI smoke run `3 -q` + the exact line of tests/test_solver.py:92-102"""
    out = _run('3', '-q')
    assert out.strip().splitlines()[-6:] == ['(0, 0, 0, 0, 0)', '(0, 0, 1, 1, 1)', '(0, 0, 1, 1, 3)', '(0, 0, 1, 1, 5)',
                                             '(0, 0, 2, 2, 6)', '(0, 0, 2, 2, 0) 3K6']


def test_cli_speedrun_beam_verbose():
    out = _run('6', '-u', '-H', 'balanced', '-w', '1000', '--noise', 'hash')
    assert 'SPEEDRUN MODE SOLVER' in out and 'Beam Width: 1,000' in out and 'turn=0' in out and 'Solution:' in out


def test_cli_realistic():
    """reference CI smoke run `6 --realistic -w 3000 -q`: 34 plies, winner P0 6-5 (SURVEY.md 8c known answer)"""
    out = _run('6', '--realistic', '-w', '3000', '-q')
    assert 'Game Over! Winner: Player 0' in out and 'Player 0: 6 points, 8 cards' in out and 'Player 1: 5 points, 8 cards' in out
    assert 'Total moves: 34' in out


def _strip_cache_noise(text):
    # the reference prints `Unpickling buys...` / `Generating buys...` when its pickle cache is first touched
    # (src/buys.py:25-36); that I/O is outside the path
    return [l for l in text.splitlines() if not l.endswith('buys...') and l != 'Pickling finished.']


def test_verbose_output_equals_reference(golden, capsys):
    """stdout of solve(verbose=True) -- banner, `turn=` and prefix-max `max_pts=` lines --
    is line for line what the reference prints (fixtures: tests/golden/verbose.json)."""
    import splendor_rl_gym_b200 as S
    want = golden['verbose']
    S.State.newgame().solve(goal_pts=6, use_heuristic=True, heuristic_name='balanced', beam_width=1000, verbose=True)
    assert capsys.readouterr().out.splitlines() == _strip_cache_noise(want['speedrun_goal6_balanced_beam1000'])
    S.State.newgame().solve(goal_pts=4, verbose=True)
    assert capsys.readouterr().out.splitlines() == _strip_cache_noise(want['speedrun_goal4_bfs'])
    cfg = S.GameConfig(num_players=2, target_points=6, gems_per_color=4, infinite_resources=False)
    S.MultiPlayerState.newgame(cfg).solve(beam_width=300, verbose=True)
    assert capsys.readouterr().out.splitlines() == _strip_cache_noise(want['realistic_2p_goal6_beam300'])
