"""Time the UNMODIFIED Python reference (staged by __graft_entry__.build() under baseline/_ref/, git-ignored) on one
host core: State.newgame().solve(goal, use_heuristic=True, heuristic_name=H, beam_width=W, verbose=False) with
`src.solver.randint` patched to 50 (noise policy `const`).  Prints one JSON line {expanded, seconds, moves}.
Expanded states are counted by wrapping State.__iter__ (one call per dequeued, expanded state).
Usage: python tools/ref_timing.py <ref_dir> <goal> <heuristic> <beam>"""
import json
import os
import sys
import time

ref, goal, hname, beam = os.path.abspath(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
import setuptools  # more_itertools (the reference's only dependency) is vendored inside setuptools here

sys.path.append(os.path.join(os.path.dirname(setuptools.__file__), '_vendor'))
sys.path.insert(0, ref)
os.chdir(ref)
import src.buys as buys

buys.BUYS_PATH = type(buys.BUYS_PATH)('/tmp/spl_ref_buys.pickle')  # the reference caches its buy table next to src/
import src.solver as rs

rs.randint = lambda a, b: 50
count = [0]
orig_iter = rs.State.__iter__


def counting_iter(self):
    count[0] += 1
    return orig_iter(self)


rs.State.__iter__ = counting_iter
rs.get_buys()  # build / load the table outside the timed region
t0 = time.perf_counter()
path = rs.State.newgame().solve(goal_pts=goal, use_heuristic=True, heuristic_name=hname, beam_width=beam, verbose=False)
dt = time.perf_counter() - t0
print(json.dumps({'expanded': count[0], 'seconds': dt, 'moves': len(path) - 1}))
