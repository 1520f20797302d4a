"""GPU tests at sizes the CPU oracle cannot follow: size-independent properties of the levels
(SURVEY.md 8c / task rule 3: sortedness, uniqueness, round trips, checksums) instead of stored answers."""
import numpy as np
import pytest
import torch

import splendor_rl_gym_b200 as S

pytestmark = pytest.mark.gpu


def _assert_unique(keys):
    # 128-bit keys: sort by (hi, lo) through two stable passes and compare neighbours
    o = torch.argsort(keys[:, 0], stable=True)
    k = keys[o]
    o2 = torch.argsort(k[:, 1], stable=True)
    k = k[o2]
    same = (k[1:, 0] == k[:-1, 0]) & (k[1:, 1] == k[:-1, 1])
    assert not bool(same.any())


def test_bfs_depth10_properties():
    """Exhaustive BFS (configs[1]) to depth 10 on one GPU: 88.7 M-state frontier."""
    eng = S.Engine(0, table_slots=400_000_000)
    k, a = S.State.newgame().record()
    sol = eng.solver(k, a, 255, False, 'simple', 0)
    sizes = []
    prev = last_parent = None
    for lv in range(10):
        info = sol.step()
        fr = sol.frontier()
        last_parent = prev
        sizes.append(fr.shape[0])
        if lv >= 8:
            link = fr[:, 3]
            # arrival order: parent ranks are non-decreasing, ordinals increase within a parent
            assert bool((link[1:] > link[:-1]).all())
            assert int(link[-1] >> 8) < info['frontier']
            _assert_unique(fr[:, :2])
            # a gem take keeps the card mask: every child's mask is a superset of its parent's mask
            parents = prev[(link >> 8)[::997]]
            kids = fr[::997]
            assert bool(((kids[:, 1] & parents[:, 1]) == parents[:, 1]).all())
            assert bool((((kids[:, 0] >> 15) & (parents[:, 0] >> 15)) == (parents[:, 0] >> 15)).all())
        prev = fr.clone() if lv >= 7 else None
    assert sizes == [15, 110, 790, 4939, 30065, 166688, 799295, 3949711, 18533132, 88749859]
    assert info['visited'] == 1 + sum(sizes)
    # round trip: re-expanding a parent reproduces the child named by (parent rank, ordinal)
    link = sol.frontier()[:, 3]
    pick = torch.randint(0, link.shape[0], (64,), device=link.device)
    child = sol.frontier()[pick].cpu().numpy().view(np.uint64)
    par = last_parent[(link[pick] >> 8)].cpu().numpy().view(np.uint64)
    for c, p in zip(child, par):
        st = S.State.from_record(int(p[0]), int(p[1]), int(p[2]))
        kid = list(st)[int(c[3]) & 0xff]
        assert kid.record() == (int(c[0]) | int(c[1]) << 64, int(c[2]))
    sol.close()
    eng.close()


@pytest.mark.parametrize('tie', ['stable', 'det'])
def test_beam_3m_properties(tie):
    """Beam search at 10x the reference's default width: rank order, tie order, uniqueness, beam saturation."""
    beam = 3_000_000
    eng = S.Engine(0, table_slots=550_000_000)
    k, a = S.State.newgame().record()
    sol = eng.solver(k, a, 15, True, 'aggressive', beam, tie, 'const')
    while True:
        info = sol.step()
        if info['ended']:
            break
        fr = sol.frontier()
        assert fr.shape[0] == min(beam, info['unique']) == info['kept']
        if info['unique'] > beam:
            sc = eng.score('aggressive', fr[:, :2].contiguous(), fr[:, 2].contiguous(), 'const')
            assert bool((sc[1:] <= sc[:-1]).all())                      # sorted by score, descending
            tie_mask = sc[1:] == sc[:-1]
            if tie == 'stable':                                          # ties keep arrival order
                assert bool((fr[1:, 3] > fr[:-1, 3])[tie_mask].all())
            else:                                                        # ties: larger key first
                hi_gt = fr[:-1, 1] > fr[1:, 1]
                hi_eq = fr[:-1, 1] == fr[1:, 1]
                lo_gt = (fr[:-1, 0] >> 1) * 2 + (fr[:-1, 0] & 1) != (fr[1:, 0] >> 1) * 2 + (fr[1:, 0] & 1)  # keys differ
                assert bool((hi_gt | (hi_eq & lo_gt))[tie_mask].all())
            _assert_unique(fr[:, :2])
    assert info['goal_rank'] == 0  # the goal state tops the queue of the last level
    ranks, ords = sol.path()
    assert len(ords) == len(sol.infos) - 1 == 15
    sol.close()
    eng.close()
