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
