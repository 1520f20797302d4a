"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol that
include/splendor_b200.h declares; host-side tables and the API mirror match the reference's
golden vectors; the product path fails loudly without a GPU."""
import hashlib
import itertools
import re
import struct
from pathlib import Path

import pytest
import torch

import splendor_rl_gym_b200 as S

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    import ctypes
    header = (ROOT / 'include' / 'splendor_b200.h').read_text()
    declared = set(re.findall(r'\b(spl_[a-z_]+)\s*\(', header))
    assert declared, 'no declarations found'
    raw = ctypes.CDLL(str(S.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(raw, name), f'{name} declared in the header but not exported'
    assert declared == set(S.EXPORTS), declared ^ set(S.EXPORTS)
    assert S.lib.spl_abi_version() == 2


def test_takes_table_matches_reference(golden):
    t = golden['tables']['takes']
    h = hashlib.sha256()
    edges = 0
    takes = S.get_takes()
    assert len(takes) == 8 ** 5  # tests/test_gems.py:296-298
    for g in itertools.product(range(8), repeat=5):
        h.update(bytes(g))
        h.update(struct.pack('<H', len(takes[g])))
        for x in takes[g]:
            h.update(bytes(x))
            edges += 1
    assert edges == t['edges'] and h.hexdigest() == t['sha']


def test_take_gems_reference_vectors():
    """Order-exact vectors of the reference's tests/test_gems.py:11-60."""
    assert S.take_gems((0, 0, 0, 0, 0)) == (
        (0, 0, 1, 1, 1), (0, 1, 0, 1, 1), (0, 1, 1, 0, 1), (0, 1, 1, 1, 0), (1, 0, 0, 1, 1),
        (1, 0, 1, 0, 1), (1, 0, 1, 1, 0), (1, 1, 0, 0, 1), (1, 1, 0, 1, 0), (1, 1, 1, 0, 0),
        (0, 0, 0, 0, 2), (0, 0, 0, 2, 0), (0, 0, 2, 0, 0), (0, 2, 0, 0, 0), (2, 0, 0, 0, 0))
    assert len(S.take_gems((6, 0, 0, 0, 0))) == 14
    assert len(S.take_gems((7, 0, 0, 0, 0))) == 8
    for g in [(2, 1, 1, 0, 0), (2, 2, 2, 1, 1), (4, 4, 0, 0, 0), (3, 3, 3, 0, 0), (2, 2, 2, 2, 2)]:
        out = S.take_gems(g)
        assert len(set(out)) == len(out)
        assert all(sum(x) <= 10 and all(0 <= v <= 7 for v in x) for x in out)


def test_buys_table_matches_reference(golden):
    t = golden['tables']['buys']
    buys = S.possible_buys()
    deck = S.get_deck()
    h = hashlib.sha256()
    for g in itertools.product(range(8), repeat=5):
        h.update(bytes(g))
        h.update(struct.pack('<H', len(buys[g])))
        h.update(bytes(buys[g]))
    assert h.hexdigest() == t['sha']
    ids = lambda k: [deck[c].str_id for c in buys[k]]  # noqa: E731  tests/test_buys.py:9-27
    assert len(buys[7, 7, 7, 7, 7]) == 90
    assert ids((0, 0, 0, 0, 0)) == [] and ids((0, 0, 0, 0, 2)) == []
    assert ids((0, 4, 0, 0, 0)) == ['0W3', '1K4']
    assert ids((0, 0, 0, 2, 4)) == ['0B3', '0W12', '1G4']
    assert ids((4, 4, 0, 1, 0)) == ['0W3', '0R3', '0G12', '0K122', '1R4', '1K4']


def test_deck_matches_reference(golden):
    deck = S.get_deck()
    assert len(deck) == 90 and len({c.str_id for c in deck}) == 90
    for c, w in zip(deck, golden['tables']['deck']):
        assert (list(c.cost), c.pt, c.bonus.value, c.str_id) == (w['cost'], w['pt'], w['bonus'], w['str_id'])
    assert str(deck[0]) == '0W3' and str(deck[89]) == '5K37'


def test_gem_arithmetic_reference_vectors(golden):
    for r in golden['tables']['subtract_with_bonus']:
        g, saved = S.subtract_with_bonus(tuple(r['gems']), tuple(r['cost']), tuple(r['bonus']))
        assert [list(g), saved] == r['out']
    assert S.increase_bonus((0, 0, 0, 0, 0), S.Color.RED) == (0, 0, 0, 1, 0)


def test_buy_card_chain():
    """tests/test_solver.py:23-52 of the reference (incl. `saved`)."""
    st = S.State.newgame()
    st.gems = (4, 3, 0, 7, 2)
    st = st.buy_card(50)
    assert (st.cards, st.bonus, st.gems, st.pts, st.saved) == ((50,), (1, 0, 0, 0, 0), (4, 3, 0, 2, 2), 2, 0)
    st = st.buy_card(6)
    assert (st.cards, st.bonus, st.gems, st.pts, st.saved) == ((6, 50), (1, 1, 0, 0, 0), (4, 3, 0, 2, 0), 2, 1)
    st = st.buy_card(57)
    assert (st.cards, st.bonus, st.gems, st.pts, st.saved) == ((6, 50, 57), (1, 1, 1, 0, 0), (1, 2, 0, 2, 0), 4, 3)
    s1 = S.State.newgame()
    for card in (40, 5, 21):
        s1 = s1.buy_card(card)
    s1.gems = (1, 2, 0, 0, 3)
    assert s1.cards == (5, 21, 40) and s1.bonus == (2, 1, 0, 0, 0)
    assert repr(s1) == '(1, 2, 0, 0, 3) 0W12-0B113-1W223'
    assert repr(S.State.newgame()) == '(0, 0, 0, 0, 0)'


def test_registry_and_fallback_name():
    assert set(S.HEURISTICS) == {'simple', 'balanced', 'aggressive', 'efficiency', 'competitive'}
    from splendor_rl_gym_b200.engine import heuristic_id
    assert heuristic_id('no-such-heuristic') == heuristic_id('simple')  # src/solver.py:429


def test_pack_roundtrip():
    from splendor_rl_gym_b200.engine import pack_aux, pack_key, unpack_record
    k = pack_key((0, 48, 49, 89), (7, 0, 3, 1, 2))
    a = pack_aux((1, 2, 3, 4, 18), 140, 65535)
    assert unpack_record(k & (2 ** 64 - 1), k >> 64, a) == ((0, 48, 49, 89), (1, 2, 3, 4, 18), (7, 0, 3, 1, 2), 140, 65535)
    with pytest.raises(ValueError):
        pack_key((), (8, 0, 0, 0, 0))


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        S.State.newgame().solve(goal_pts=3, verbose=False)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        list(S.State.newgame())
    import ctypes as C
    from splendor_rl_gym_b200._lib import Config
    h = C.c_void_p()
    rc = S.lib.spl_create(C.byref(Config(0, 0, 0, 0, 0)), C.byref(h))
    assert rc == -2 and b'no CPU fallback' in S.lib.spl_last_error(None)


def test_buys_cache_files_are_reference_compatible(tmp_path, golden):
    """buys.pickle / buys.txt written from the native table load as the reference's own (src/buys.py:20-49)."""
    import pickle
    from splendor_rl_gym_b200.buys import export_buys_to_txt, get_buys, store_buys
    store_buys(get_buys(), tmp_path / 'buys.pickle')
    back = pickle.load(open(tmp_path / 'buys.pickle', 'rb'))
    assert back == get_buys() and (tmp_path / 'buys.pickle').stat().st_size > 1_000_000  # tests/test_buys.py:30-38
    for k, want in golden['tables']['buys']['samples'].items():
        assert list(back[tuple(map(int, k.split(',')))]) == want
    export_buys_to_txt(tmp_path / 'buys.txt')
    first = open(tmp_path / 'buys.txt').readline()
    assert first == '(0, 0, 0, 0, 0): ()\n'


def test_python_ids_match_header_enums():
    """The name -> id tables of the host mirror are the enum values include/splendor_b200.h declares."""
    from splendor_rl_gym_b200 import engine
    header = (ROOT / 'include' / 'splendor_b200.h').read_text()
    enum = {k: int(v) for k, v in re.findall(r'\b(SPL_[A-Z_]+)\s*=\s*(-?\d+)', header)}
    assert engine.HEURISTIC_IDS == {'simple': enum['SPL_H_SIMPLE'], 'balanced': enum['SPL_H_BALANCED'],
                                    'aggressive': enum['SPL_H_AGGRESSIVE'], 'efficiency': enum['SPL_H_EFFICIENCY'],
                                    'competitive': enum['SPL_H_BALANCED']}  # alias, src/solver.py:289-296
    assert engine.NOISE_IDS == {'const': enum['SPL_NOISE_CONST'], 'hash': enum['SPL_NOISE_HASH'],
                                'mt': enum['SPL_NOISE_EXTERNAL']}
    assert engine.TIE_IDS == {'stable': enum['SPL_TIE_STABLE'], 'det': enum['SPL_TIE_KEY'], 'key': enum['SPL_TIE_KEY'],
                              'det_ordered': enum['SPL_TIE_KEY_ORDERED']}
    assert engine.IDENTITY_IDS == {'key': enum['SPL_IDENT_KEY'], 'pyhash': enum['SPL_IDENT_PYHASH']}
