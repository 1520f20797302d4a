#!/usr/bin/env python
"""Generate golden fixtures for the frontier-expansion path FROM THE REFERENCE ITSELF.

Runs only in the authoring container (needs /root/reference); the GPU box never runs
this.  It imports the unmodified reference modules (src/solver.py, src/gems.py,
src/buys.py, src/cardparser.py) and drives them with a stepper that mirrors the
`while queue` loop of src/solver.py:434-457 while calling the reference's own
`State.__iter__`, `State.__hash__`, `HEURISTICS[...]` and `sorted(...)`.

The only patches applied to the reference are the ones SURVEY.md §8a-N names:
  * `src.buys.BUYS_PATH` -> a /tmp file (the reference tree is read-only),
  * `src.solver.randint` -> a deterministic noise source (the reference draws its
    tie-break noise from the unseeded global Mersenne Twister),
  * plug-in entries added to `HEURISTICS` (the reference's own extension point,
    src/solver.py:299-305,429) for the key-tie-break (`det`) policy.

Usage:  python tests/golden/make_golden.py [--bfs-depth 8] [--out tests/golden]
"""
import argparse
import hashlib
import json
import os
import struct
import sys
import time
from pathlib import Path

REF = os.environ.get('SPLENDOR_REFERENCE', '/root/reference')
import setuptools  # noqa: E402  (more_itertools is vendored inside setuptools)

sys.path.append(os.path.join(os.path.dirname(setuptools.__file__), '_vendor'))
sys.path.insert(0, REF)

import src.buys as ref_buys  # noqa: E402

ref_buys.BUYS_PATH = Path('/tmp/splendor_ref_buys.pickle')

import src.solver as ref_solver  # noqa: E402
from src.gems import get_takes, subtract_with_bonus, take_gems  # noqa: E402
from src.solver import HEURISTICS, GameConfig, MultiPlayerState, State, deck  # noqa: E402

M64 = (1 << 64) - 1


def key_int(s) -> int:
    """Canonical 128-bit key: card mask << 15 | gems (3 bits each)."""
    m = 0
    for c in s.cards:
        m |= 1 << c
    g = 0
    for i, x in enumerate(s.gems):
        g |= x << (3 * i)
    return (m << 15) | g


def mix64(x: int) -> int:
    """splitmix64 finaliser over the folded 128-bit key (noise policy `hash`)."""
    x = (x ^ (x >> 64)) & M64
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


class Noise:
    """Deterministic stand-in for `random.randint` inside src.solver."""

    def __init__(self):
        self.mode = 'const'
        self.cur = 50

    def __call__(self, a, b):
        return self.cur


NOISE = Noise()
ref_solver.randint = NOISE


def _with_noise(fn):
    def h(s):
        if NOISE.mode == 'hash':
            NOISE.cur = 1 + mix64(key_int(s)) % 100
        else:
            NOISE.cur = 50
        return fn(s)
    return h


BASE = {n: HEURISTICS[n] for n in ('simple', 'balanced', 'aggressive', 'efficiency', 'competitive')}
for _n, _f in BASE.items():
    HEURISTICS[_n + '@stable'] = _with_noise(_f)                       # ties -> arrival order
    HEURISTICS[_n + '@det'] = (lambda f: (lambda s: (f(s), key_int(s))))(_with_noise(_f))  # ties -> key desc


def level_digest(states, parents=None):
    """Order-independent sums + order-dependent sha over (key, saved) records."""
    n = len(states)
    sl = sh = xl = xh = 0
    ssaved = spts = 0
    h = hashlib.sha256()
    for s in states:
        k = key_int(s)
        lo, hi = k & M64, k >> 64
        sl = (sl + lo) & M64
        sh = (sh + hi) & M64
        xl ^= lo
        xh ^= hi
        ssaved += s.saved
        spts += s.pts
        h.update(struct.pack('<QQH', lo, hi, s.saved))
    d = dict(n=n, sum_lo=sl, sum_hi=sh, xor_lo=xl, xor_hi=xh, sum_saved=ssaved, sum_pts=spts,
             sha=h.hexdigest())
    if parents is not None:
        d['sum_parent_rank'] = sum(p for p, _ in parents)
        d['sum_ordinal'] = sum(o for _, o in parents)
    return d


def stepper(goal_pts, use_heuristic, hname, beam, max_levels=None, keep_levels=False):
    """Mirror of src/solver.py:425-457 that records per-level digests."""
    root = State.newgame()
    queue = [root]
    trail = {root: None}
    heuristic = HEURISTICS.get(hname, ref_solver.simple_heuristic)
    levels = []
    expanded_total = generated_total = 0
    turn = 0
    puzzle = root
    t0 = time.time()
    kept_levels = []
    while queue:
        next_queue = []
        links = []
        expanded = generated = 0
        goal_rank = None
        for rank, puzzle in enumerate(queue):
            if puzzle.pts >= goal_pts:
                next_queue.clear()
                links.clear()
                goal_rank = rank
                break
            expanded += 1
            for ordinal, nxt in enumerate(puzzle):
                generated += 1
                if nxt in trail:
                    continue
                trail[nxt] = puzzle
                next_queue.append(nxt)
                links.append((rank, ordinal))
        expanded_total += expanded
        generated_total += generated
        rec = dict(level=turn, frontier=len(queue), expanded=expanded, generated=generated,
                   goal_rank=goal_rank, unique=level_digest(next_queue, links))
        if use_heuristic:
            order = sorted(range(len(next_queue)), key=lambda i: heuristic(next_queue[i]), reverse=True)[:beam]
            queue = [next_queue[i] for i in order]
            rec['kept'] = level_digest(queue, [links[i] for i in order])
        else:
            queue = next_queue
        rec['secs'] = round(time.time() - t0, 3)
        levels.append(rec)
        if keep_levels:
            kept_levels.append(queue)
        turn += 1
        print(f'  level {turn}: frontier={len(queue)} generated={generated} t={rec["secs"]}s', file=sys.stderr)
        if max_levels is not None and turn >= max_levels:
            break
    sol = []
    while puzzle:
        sol.append(puzzle)
        puzzle = trail[puzzle]
    sol.reverse()
    out = dict(goal=goal_pts, use_heuristic=use_heuristic, heuristic=hname, beam=beam,
               moves=len(sol) - 1, expanded=expanded_total, generated=generated_total,
               visited=len(trail), levels=levels,
               path=[dict(repr=repr(s), key=str(key_int(s)), pts=s.pts, saved=s.saved,
                          bonus=list(s.bonus), cards=list(s.cards), gems=list(s.gems)) for s in sol])
    return (out, kept_levels) if keep_levels else out


def gen_tables():
    takes = get_takes()
    h = hashlib.sha256()
    edges = 0
    for g in sorted(takes):
        h.update(bytes(g))
        h.update(struct.pack('<H', len(takes[g])))
        for t in takes[g]:
            h.update(bytes(t))
            edges += 1
    buys = ref_buys.possible_buys()
    hb = hashlib.sha256()
    refs = 0
    for g in sorted(buys):
        hb.update(bytes(g))
        hb.update(struct.pack('<H', len(buys[g])))
        hb.update(bytes(buys[g]))
        refs += len(buys[g])
    sample_hands = [(0,0,0,0,0),(6,0,0,0,0),(7,0,0,0,0),(2,1,1,0,0),(2,2,2,1,1),(4,4,0,0,0),
                    (3,3,2,0,0),(3,3,3,0,0),(4,4,1,0,0),(2,2,2,2,2),(7,3,0,0,0),(4,3,2,1,0),
                    (1,2,0,0,3),(0,0,2,2,6),(3,2,2,2,1),(5,5,0,0,0),(7,2,1,0,0),(1,1,1,1,6)]
    sample_keys = [(0,0,0,0,0),(0,0,0,0,2),(0,4,0,0,0),(0,0,0,2,4),(4,4,0,1,0),(7,7,7,7,7),
                   (3,3,3,3,3),(2,5,1,0,7),(6,0,6,0,6)]
    return dict(
        deck=[dict(cost=list(c.cost), pt=c.pt, bonus=c.bonus.value, str_id=c.str_id) for c in deck],
        takes=dict(keys=len(takes), live=sum(1 for v in takes.values() if v), edges=edges,
                   max_fanout=max(len(v) for v in takes.values()), sha=h.hexdigest(),
                   samples={','.join(map(str, g)): [list(t) for t in take_gems(g)] for g in sample_hands}),
        buys=dict(keys=len(buys), refs=refs, sha=hb.hexdigest(),
                  samples={','.join(map(str, g)): list(buys[g]) for g in sample_keys}),
        subtract_with_bonus=[dict(gems=g, cost=c, bonus=b, out=[list(subtract_with_bonus(g, c, b)[0]),
                                                               subtract_with_bonus(g, c, b)[1]])
                             for g, c, b in [((5,4,3,2,1),(1,2,3,4,5),(0,1,2,3,4)),
                                             ((4,3,0,7,2),(0,0,0,5,0),(0,0,0,0,0)),
                                             ((4,3,0,2,2),(0,0,0,0,3),(1,0,0,0,0)),
                                             ((7,7,7,7,7),(3,3,5,3,0),(2,0,6,1,1))]],
    )


def gen_successors_and_scores(levels, per_level=24):
    """Sample reachable states; record their full ordered successor lists and f64 scores."""
    import random
    rng = random.Random(1234)
    picks = []
    for q in levels:
        picks += rng.sample(q, min(per_level, len(q)))
    succ = []
    for s in picks:
        kids = [dict(key=str(key_int(c)), saved=c.saved, pts=c.pts, bonus=list(c.bonus)) for c in s]
        succ.append(dict(cards=list(s.cards), gems=list(s.gems), bonus=list(s.bonus), pts=s.pts,
                         saved=s.saved, key=str(key_int(s)), children=kids))
    scores = []
    for s in picks:
        row = dict(key=str(key_int(s)), saved=s.saved, pts=s.pts, bonus=list(s.bonus),
                   gems=list(s.gems), ncards=len(s.cards))
        for mode in ('const', 'hash'):
            NOISE.mode = mode
            for n in ('simple', 'balanced', 'aggressive', 'efficiency'):
                v = HEURISTICS[n + '@stable'](s)
                row[f'{n}:{mode}'] = struct.unpack('<Q', struct.pack('<d', v))[0]
        NOISE.mode = 'const'
        scores.append(row)
    # synthetic wide-range states (as tests/test_heuristics.py builds them: fields need not be consistent)
    for pts, saved, bonus, gems, ncards in [(0,0,(0,0,0,0,0),(0,0,0,0,0),0),(2,5,(0,0,0,0,0),(0,0,0,0,0),0),
                                            (12,5,(0,0,0,0,0),(0,0,0,0,0),0),(15,28,(3,2,4,1,5),(1,0,2,0,0),15),
                                            (22,61,(5,5,5,4,4),(0,0,0,0,0),23),(7,1000,(18,0,0,0,0),(7,3,0,0,0),18),
                                            (140,65535,(18,18,18,18,18),(2,2,2,2,2),90)]:
        cards = tuple(range(ncards))
        s = State(cards=cards, bonus=bonus, gems=gems, pts=pts, saved=saved)
        row = dict(key=str(key_int(s)), saved=saved, pts=pts, bonus=list(bonus), gems=list(gems), ncards=ncards)
        for mode in ('const', 'hash'):
            NOISE.mode = mode
            for n in ('simple', 'balanced', 'aggressive', 'efficiency'):
                v = HEURISTICS[n + '@stable'](s)
                row[f'{n}:{mode}'] = struct.unpack('<Q', struct.pack('<d', v))[0]
        NOISE.mode = 'const'
        scores.append(row)
    return succ, scores


# ---------------------------------------------------------------- realistic mode (src/solver.py:471-860)
def rrec_bytes(s) -> bytes:
    """Canonical identity bytes of a MultiPlayerState: per player (card mask u64+u32, gems u16, saved u16),
    12 visible slots in slot order (255 = empty), current player."""
    out = b''
    for p in s.players:
        m = 0
        for c in p.cards:
            m |= 1 << c
        g = 0
        for i, x in enumerate(p.gems):
            g |= x << (3 * i)
        out += struct.pack('<QIHH', m & M64, m >> 64, g, p.saved)
    vis = []
    for tier in (s.market.tier1_visible, s.market.tier2_visible, s.market.tier3_visible):
        vis += list(tier) + [255] * (4 - len(tier))
    out += bytes(vis) + bytes([s.current_player])
    return out


def rlevel_digest(states, links=None):
    h = hashlib.sha256()
    for s in states:
        h.update(rrec_bytes(s))
    d = dict(n=len(states), sha=h.hexdigest(),
             sum_saved=sum(p.saved for s in states for p in s.players),
             sum_pts=sum(p.pts for s in states for p in s.players),
             sum_pool=sum(sum(s.gem_pool.available) for s in states))
    if links is not None:
        d['sum_parent_rank'] = sum(p for p, _ in links)
        d['sum_ordinal'] = sum(o for _, o in links)
    return d


def rstepper(players, goal, seed, beam, gems_per_color=None):
    """Mirror of MultiPlayerState.solve's loop (src/solver.py:820-852) recording per-level digests."""
    gpc = gems_per_color or {2: 4, 3: 5, 4: 7}[players]
    cfg = GameConfig(num_players=players, target_points=goal, gems_per_color=gpc, infinite_resources=False)
    root = MultiPlayerState.newgame(cfg, shuffle_market=seed is not None, seed=seed)
    # the reference's closure heuristic (src/solver.py:778-812) is not reachable from outside solve();
    # obtain the identical code object by running solve() on a finished game and capturing `sorted`'s key
    captured = {}
    import builtins
    real_sorted = builtins.sorted

    def spy(it, key=None, reverse=False):
        if key is not None and 'h' not in captured:
            captured['h'] = key
        return real_sorted(it, key=key, reverse=reverse)
    ref_solver.sorted = spy
    try:
        tiny = MultiPlayerState.newgame(GameConfig(num_players=players, target_points=0, gems_per_color=gpc,
                                                   infinite_resources=False))
        # target 0 -> `any(p.pts >= 0)` ends the game at the first dequeue; sorted() is still called once
        tiny.solve(beam_width=1, verbose=False)
    finally:
        del ref_solver.sorted
    heuristic = captured['h']
    queue = [root]
    trail = {root: None}
    levels = []
    turn = 0
    puzzle = root
    smin, smax = None, None
    expanded_total = generated_total = 0
    t0 = time.time()
    while queue:
        next_queue, links = [], []
        expanded = generated = 0
        goal_rank = None
        for rank, puzzle in enumerate(queue):
            if puzzle.is_game_over():
                next_queue.clear()
                links.clear()
                goal_rank = rank
                break
            expanded += 1
            for ordinal, nxt in enumerate(puzzle):
                generated += 1
                if nxt in trail:
                    continue
                trail[nxt] = puzzle
                next_queue.append(nxt)
                links.append((rank, ordinal))
        scores = [heuristic(s) for s in next_queue]
        if scores:
            smin = min(scores) if smin is None else min(smin, min(scores))
            smax = max(scores) if smax is None else max(smax, max(scores))
        sh = hashlib.sha256()
        for v in scores:
            sh.update(struct.pack('<d', v))
        order = sorted(range(len(next_queue)), key=lambda i: scores[i], reverse=True)[:beam]
        rec = dict(level=turn, frontier=len(queue), expanded=expanded, generated=generated, goal_rank=goal_rank,
                   unique=rlevel_digest(next_queue, links), scores_sha=sh.hexdigest())
        queue = [next_queue[i] for i in order]
        rec['kept'] = rlevel_digest(queue, [links[i] for i in order])
        levels.append(rec)
        expanded_total += expanded
        generated_total += generated
        turn += 1
        if turn > 1000:
            break
    sol = []
    while puzzle:
        sol.append(puzzle)
        puzzle = trail.get(puzzle)
    sol.reverse()
    final = sol[-1]
    root_market = dict(t1=list(root.market.tier1_visible) + list(root.market.tier1_deck),
                       t2=list(root.market.tier2_visible) + list(root.market.tier2_deck),
                       t3=list(root.market.tier3_visible) + list(root.market.tier3_deck))
    out = dict(players=players, goal=goal, seed=seed, beam=beam, gems_per_color=gpc, plies=len(sol) - 1,
               winner=final.get_winner(), expanded=expanded_total, generated=generated_total, visited=len(trail),
               min_score=smin, max_score=smax, market=root_market, levels=levels, wall_s=round(time.time() - t0, 2),
               final=[dict(pts=p.pts, cards=list(p.cards), saved=p.saved, gems=list(p.gems)) for p in final.players],
               path_sha=[hashlib.sha256(rrec_bytes(s)).hexdigest()[:16] for s in sol])
    # cross-check against the reference's own unmodified solve()
    ref_sol = MultiPlayerState.newgame(cfg, shuffle_market=seed is not None, seed=seed).solve(beam_width=beam, verbose=False)
    assert [rrec_bytes(s) for s in ref_sol] == [rrec_bytes(s) for s in sol], 'stepper != reference solve()'
    return out


def gen_realistic_successors(n_states=60):
    import random
    rng = random.Random(99)
    out = []
    for players, seed in ((2, 0), (3, 1), (2, None), (4, 5)):
        gpc = {2: 4, 3: 5, 4: 7}[players]
        cfg = GameConfig(num_players=players, target_points=15, gems_per_color=gpc, infinite_resources=False)
        s = MultiPlayerState.newgame(cfg, shuffle_market=seed is not None, seed=seed)
        market = dict(t1=list(s.market.tier1_visible) + list(s.market.tier1_deck),
                      t2=list(s.market.tier2_visible) + list(s.market.tier2_deck),
                      t3=list(s.market.tier3_visible) + list(s.market.tier3_deck))
        for _ in range(n_states):
            kids = list(s)
            out.append(dict(players=players, gems_per_color=gpc, market=market, state=rrec_bytes(s).hex(),
                            children=[rrec_bytes(k).hex() for k in kids],
                            over=[k.is_game_over() for k in kids]))
            if not kids:
                break
            # random walk biased towards buys so that markets refill and decks run down
            buys = [k for k in kids if sum(len(p.cards) for p in k.players) > sum(len(p.cards) for p in s.players)]
            s = rng.choice(buys) if buys and rng.random() < 0.7 else rng.choice(kids)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bfs-depth', type=int, default=8)
    ap.add_argument('--out', default=str(Path(__file__).parent))
    ap.add_argument('--only', default='')
    a = ap.parse_args()
    out = Path(a.out)
    only = set(a.only.split(',')) if a.only else None

    def want(x):
        return only is None or x in only

    if want('tables'):
        json.dump(gen_tables(), open(out / 'tables.json', 'w'), indent=0)
        print('tables done', file=sys.stderr)

    if want('bfs'):
        t0 = time.time()
        bfs, lv = stepper(10**9, False, 'simple', 0, max_levels=a.bfs_depth, keep_levels=True)
        bfs['path'] = []
        bfs['wall_s'] = round(time.time() - t0, 2)
        json.dump(bfs, open(out / 'bfs_levels.json', 'w'), indent=0)
        succ, scores = gen_successors_and_scores([[State.newgame()]] + lv[:7])
        json.dump(succ, open(out / 'successors.json', 'w'))
        json.dump(scores, open(out / 'scores.json', 'w'), indent=0)
        del lv
        # reference's own unmodified solve(): BFS winning lines (tests/test_solver.py:92-113)
        lines = {}
        for goal in (3, 4):
            lines[str(goal)] = [repr(s) for s in State.newgame().solve(goal_pts=goal, verbose=False)]
        json.dump(lines, open(out / 'bfs_lines.json', 'w'), indent=0)

    if want('realistic'):
        NOISE.mode = 'const'
        json.dump(gen_realistic_successors(), open(out / 'realistic_successors.json', 'w'))
        runs = []
        for players, goal, seed, beam in [(2, 6, None, 300), (2, 6, None, 3000), (2, 10, 0, 500), (3, 8, 0, 400),
                                          (2, 15, 0, 2000), (3, 15, 0, 1000), (4, 6, 3, 200), (2, 15, 7, 20)]:
            r = rstepper(players, goal, seed, beam)
            runs.append(r)
            print(f'realistic p={players} goal={goal} seed={seed} beam={beam}: plies={r["plies"]} winner={r["winner"]} '
                  f'expanded={r["expanded"]} visited={r["visited"]} {r["wall_s"]}s', file=sys.stderr)
        json.dump(runs, open(out / 'realistic_runs.json', 'w'), indent=0)

    if want('mt'):
        # The UNMODIFIED reference, only seeded: random.seed(S); solve(...).  No patched randint, no plug-ins.
        import random as pyrandom
        ref_solver.randint = pyrandom.randint
        try:
            runs = []
            for h, goal, beam, seed in [('simple', 10, 2000, 0), ('aggressive', 15, 5000, 1), ('balanced', 15, 1000, 7),
                                        ('efficiency', 12, 3000, 3), ('simple', 15, 300, 11)]:
                pyrandom.seed(seed)
                sol = State.newgame().solve(goal_pts=goal, use_heuristic=True, heuristic_name=h, beam_width=beam, verbose=False)
                after = [pyrandom.randint(1, 100) for _ in range(4)]
                pyrandom.seed(seed)
                st = stepper(goal, True, h, beam)
                assert [p['repr'] for p in st['path']] == [repr(x) for x in sol]
                runs.append(dict(mode='speedrun', heuristic=h, goal=goal, beam=beam, seed=seed, moves=len(sol) - 1,
                                 path=[dict(repr=repr(x), saved=x.saved, pts=x.pts) for x in sol], after=after,
                                 visited=st['visited'], levels=st['levels']))
                print(f'mt speedrun {h} goal={goal} beam={beam} seed={seed}: moves={len(sol) - 1}', file=sys.stderr)
            for players, goal, mseed, beam, seed in [(2, 10, 0, 500, 5), (3, 8, 1, 300, 9), (2, 15, None, 1000, 2)]:
                gpc = {2: 4, 3: 5, 4: 7}[players]
                cfg = GameConfig(num_players=players, target_points=goal, gems_per_color=gpc, infinite_resources=False)
                pyrandom.seed(seed)
                sol = MultiPlayerState.newgame(cfg, shuffle_market=mseed is not None, seed=mseed).solve(beam_width=beam, verbose=False)
                after = [pyrandom.randint(1, 100) for _ in range(4)]
                runs.append(dict(mode='realistic', players=players, goal=goal, market_seed=mseed, beam=beam, seed=seed,
                                 plies=len(sol) - 1, winner=sol[-1].get_winner(), after=after,
                                 path_sha=[hashlib.sha256(rrec_bytes(x)).hexdigest()[:16] for x in sol],
                                 final=[dict(pts=p.pts, cards=list(p.cards), saved=p.saved) for p in sol[-1].players]))
                print(f'mt realistic p={players} goal={goal} beam={beam} seed={seed}: plies={len(sol) - 1}', file=sys.stderr)
            json.dump(runs, open(out / 'mt_runs.json', 'w'), indent=0)
        finally:
            ref_solver.randint = NOISE

    if want('pyhash'):
        # the reference's identity: State.hash = hash((cards, gems)) (src/solver.py:316), as an unsigned 64-bit
        # value, for a spread of reachable states (shallow BFS states + card-rich states of a beam search)
        NOISE.mode = 'const'
        _, lv = stepper(10**9, False, 'simple', 0, max_levels=6, keep_levels=True)
        sample = [State.newgame()]
        for states in lv:
            sample += states[::max(1, len(states) // 300)]
        _, lv = stepper(15, True, 'aggressive@stable', 400, keep_levels=True)
        for states in lv[6:]:
            sample += states[::max(1, len(states) // 60)]
        rows = [[str(key_int(x)), str(x.hash & M64)] for x in sample]
        json.dump(dict(note='[canonical 128-bit key, hash((cards, gems)) & (2**64-1)] per state', rows=rows),
                  open(out / 'pyhash.json', 'w'))
        print(f'pyhash: {len(rows)} states, max cards {max(len(x.cards) for x in sample)}', file=sys.stderr)

    if want('verbose'):
        # stdout of the reference's own solve(verbose=True) (noise const, ties by arrival order)
        import contextlib
        import io
        NOISE.mode = 'const'
        texts = {}
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            State.newgame().solve(goal_pts=6, use_heuristic=True, heuristic_name='balanced@stable', beam_width=1000, verbose=True)
        texts['speedrun_goal6_balanced_beam1000'] = buf.getvalue().replace('balanced@stable', 'balanced')
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            State.newgame().solve(goal_pts=4, verbose=True)
        texts['speedrun_goal4_bfs'] = buf.getvalue()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            MultiPlayerState.newgame(GameConfig(num_players=2, target_points=6, gems_per_color=4, infinite_resources=False)).solve(
                beam_width=300, verbose=True)
        texts['realistic_2p_goal6_beam300'] = buf.getvalue()
        json.dump(texts, open(out / 'verbose.json', 'w'), indent=0)

    if want('beam'):
        runs = []
        cfgs = []
        for h in ('simple', 'balanced', 'aggressive', 'efficiency'):
            for pol in ('stable', 'det'):
                for noise in ('const', 'hash'):
                    cfgs.append((h, pol, noise, 6, 1000))
                    cfgs.append((h, pol, noise, 15, 1000))
        cfgs += [('aggressive', 'stable', 'const', 15, 20000), ('aggressive', 'det', 'const', 15, 20000),
                 ('balanced', 'det', 'hash', 15, 20000), ('simple', 'stable', 'const', 10, 5000),
                 ('efficiency', 'stable', 'hash', 15, 7)]
        for h, pol, noise, goal, beam in cfgs:
            NOISE.mode = noise
            t0 = time.time()
            r = stepper(goal, True, f'{h}@{pol}', beam)
            r.update(policy=pol, noise=noise, base=h, wall_s=round(time.time() - t0, 2))
            # cross-check with the reference's unmodified solve()
            sol = State.newgame().solve(goal_pts=goal, use_heuristic=True, heuristic_name=f'{h}@{pol}',
                                        beam_width=beam, verbose=False)
            assert [repr(s) for s in sol] == [p['repr'] for p in r['path']], (h, pol, noise, goal, beam)
            assert [s.saved for s in sol] == [p['saved'] for p in r['path']]
            runs.append(r)
            print(f'beam {h}@{pol}/{noise} goal={goal} beam={beam}: moves={r["moves"]} '
                  f'expanded={r["expanded"]} visited={r["visited"]} {r["wall_s"]}s', file=sys.stderr)
        NOISE.mode = 'const'
        json.dump(runs, open(out / 'beam_runs.json', 'w'), indent=0)


if __name__ == '__main__':
    main()
