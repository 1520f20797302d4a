"""GPU parity tests (-m gpu) of realistic multi-player mode (SURVEY.md 8 rows a-R1..a-R5): the CUDA path
through the C ABI against the reference-generated fixtures and against the CPU oracle."""
import hashlib

import numpy as np
import pytest

import oracle
import splendor_rl_gym_b200 as S
from common import assert_digest, rlevel_digest
from splendor_rl_gym_b200.realistic import RConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    return S.Engine.get(0)


def _cfg(players, goal, gpc, market, noise=0):
    cfg = RConfig(players, goal, gpc, noise)
    for t, seq in enumerate((market['t1'], market['t2'], market['t3'])):
        cfg.deck_len[t] = len(seq)
        for i, c in enumerate(seq):
            cfg.deck[t][i] = c
    return cfg


def _rec_from_ident(raw: bytes, P: int):
    rec = np.zeros(1, oracle.RREC_DTYPE)
    rec['p'][0][:P] = np.frombuffer(raw[:16 * P], oracle.RPLAYER_DTYPE)
    rec['vis'][0] = np.frombuffer(raw[16 * P:16 * P + 12], np.uint8)
    rec['cur'] = raw[16 * P + 12]
    return rec


def test_rexpand_matches_reference_successors(eng, golden):
    """MultiPlayerState.__iter__ (buys in slot order, market refill, colour triples, double takes)."""
    for s in golden['realistic_successors']:
        P = s['players']
        cfg = _cfg(P, 15, s['gems_per_color'], s['market'])
        kids = eng.rexpand(cfg, _rec_from_ident(bytes.fromhex(s['state']), P))
        assert [oracle.rident_bytes(k, P).hex() for k in kids] == s['children']
        assert [int(k['link']) & 0xff for k in kids] == list(range(len(kids)))


def test_rscore_bit_exact_vs_oracle(eng, golden):
    """multi_competitive_heuristic: f64 bit patterns (negative scores, -0.0 point differences included)."""
    run = golden['realistic_runs'][4]  # 2 players, goal 15, seed 0, beam 2000
    P, gpc = run['players'], run['gems_per_color']
    ocfg = oracle.make_rconfig(P, run['goal'], gpc, [run['market']['t1'], run['market']['t2'], run['market']['t3']])
    cfg = _cfg(P, run['goal'], gpc, run['market'])
    s = oracle.RSolver(ocfg, run['beam'])
    for _ in range(30):
        s.step()
    recs = np.concatenate([s.level(i) for i in (5, 12, 20, 29)])
    got = eng.rscore(cfg, recs).view(np.uint64)
    want = oracle.r_score(ocfg, recs).view(np.uint64)
    assert (got == want).all()
    assert (oracle.r_score(ocfg, recs) < 0).any()
    s.close()


@pytest.mark.parametrize('run_idx', [4, 5, 6])
def test_exact_identity_key_is_injective(eng, golden, run_idx):
    """spl_rpack: two states get the same 384-bit key iff their identity bytes (players incl. saved, visible cards in
    slot order, current player: src/solver.py:495-500) are equal -- over the queues of a 2-, 3- and 4-player search."""
    run = golden['realistic_runs'][run_idx]
    P, gpc = run['players'], run['gems_per_color']
    ocfg = oracle.make_rconfig(P, run['goal'], gpc, [run['market']['t1'], run['market']['t2'], run['market']['t3']])
    cfg = _cfg(P, run['goal'], gpc, run['market'])
    s = oracle.RSolver(ocfg, run['beam'])
    for _ in range(25):
        s.step()
    recs = np.concatenate([s.level(i) for i in range(3, 25)])
    s.close()
    keys = eng.rpack(cfg, recs)
    ident = [oracle.rident_bytes(r, P) for r in recs]
    assert len({bytes(k) for k in keys}) == len(set(ident))
    by_key = {}
    for k, i in zip(keys, ident):
        assert by_key.setdefault(bytes(k), i) == i


def test_rsolver_vs_reference_and_oracle(eng, golden):
    """MultiPlayerState.solve: every ply's kept set in rank order (incl. saved, market order, parent links),
    visited count, ply count and the winning line vs the unmodified reference; full arrays vs the oracle."""
    for run in golden['realistic_runs']:
        P, gpc = run['players'], run['gems_per_color']
        what = f"realistic p={P} goal={run['goal']} seed={run['seed']} beam={run['beam']}"
        cfg = _cfg(P, run['goal'], gpc, run['market'])
        ocfg = oracle.make_rconfig(P, run['goal'], gpc, [run['market']['t1'], run['market']['t2'], run['market']['t3']])
        root = oracle.r_root(ocfg)
        sol = eng.rsolver(cfg, root, run['beam'])
        orc = oracle.RSolver(ocfg, run['beam'])
        for want in run['levels']:
            gi, oi = sol.step(), orc.step()
            assert gi['frontier'] == want['frontier'] and gi['goal_rank'] == (want['goal_rank'] if want['goal_rank'] is not None else -1), what
            if gi['ended']:
                break
            assert (gi['generated'], gi['unique'], gi['visited']) == (want['generated'], want['unique']['n'], oi['visited']), (what, want['level'])
            fr = sol.frontier()
            assert_digest(rlevel_digest(fr, P, gpc), want['kept'], f"{what} level {want['level']}")
            assert fr.tobytes() == orc.level(oi['level'] + 1).tobytes(), (what, want['level'])
        assert gi['ended'] and orc.done, what
        ranks, ords = sol.path()
        assert len(ords) == run['plies'], what
        sol.close()
        orc.close()


@pytest.mark.parametrize('players,goal,seed,beam', [(2, 6, None, 300), (3, 8, 0, 400), (2, 15, 7, 20)])
def test_multiplayer_solve_api(golden, players, goal, seed, beam):
    """The reference-facing API: MultiPlayerState.newgame(...).solve(...) returns the reference's line."""
    run = next(r for r in golden['realistic_runs'] if (r['players'], r['goal'], r['seed'], r['beam']) == (players, goal, seed, beam))
    gpc = {2: 4, 3: 5, 4: 7}[players]
    cfg = S.GameConfig(num_players=players, target_points=goal, gems_per_color=gpc, infinite_resources=False)
    root = S.MultiPlayerState.newgame(cfg, shuffle_market=seed is not None, seed=seed)
    assert list(root.market.tier1_visible + root.market.tier1_deck) == run['market']['t1']
    sol = root.solve(beam_width=beam, verbose=False)
    assert len(sol) - 1 == run['plies'] and sol[-1].turn_number == run['plies']
    assert sol[-1].is_game_over() and sol[-1].get_winner() == run['winner']
    assert [dict(pts=p.pts, cards=list(p.cards), saved=p.saved, gems=list(p.gems)) for p in sol[-1].players] == run['final']
    assert [hashlib.sha256(s.record()[0:1]['p'][0][:players].tobytes() + s.record()['vis'][0].tobytes()
                           + bytes([s.current_player])).hexdigest()[:16] for s in sol] == run['path_sha']
    # structural invariants of the reference's tests/test_realistic.py:189-219
    for a, b in zip(sol, sol[1:]):
        assert b.current_player == (a.current_player + 1) % players
        assert all(x >= 0 for x in b.gem_pool.available) and all(sum(p.gems) <= 10 for p in b.players)


def test_realistic_defaults_and_pool_ops():
    """tests/test_realistic.py:8-87 of the reference: defaults, pool arithmetic, market refill."""
    c = S.GameConfig()
    assert (c.num_players, c.target_points, c.gems_per_color, c.cards_visible_per_tier, c.infinite_resources) == (2, 15, 4, 4, True)
    pool = S.GemPool.new_pool(4)
    assert pool.available == (4,) * 5 and pool.can_take_three_different((1, 1, 1, 0, 0)) and not pool.can_take_three_different((1, 1, 0, 0, 0))
    assert pool.can_take_two_same((2, 0, 0, 0, 0)) and not pool.take((1, 0, 0, 0, 0)).can_take_two_same((2, 0, 0, 0, 0))
    assert pool.take((1, 1, 1, 0, 0)).return_gems((1, 1, 1, 0, 0)) == pool
    m = S.CardMarket.from_full_deck()
    assert len(m.all_visible_cards()) == 12 and (len(m.tier1_deck), len(m.tier2_deck), len(m.tier3_deck)) == (31, 26, 21)
    first = m.tier1_visible[0]
    m2 = m.buy_card(first)
    assert first not in m2.all_visible_cards() and len(m2.tier1_visible) == 4 and m2.tier1_visible[-1] == m.tier1_deck[0]
    g = S.MultiPlayerState.newgame()
    assert len(g.players) == 2 and g.current_player == 0 and not g.is_game_over() and g.get_winner() is None
