"""Terminal settlement (finish_game / _share_value) against golden terminal states produced by the reference's own
GameEnv (tests/golden/make_settle_golden.py): chips won, game result and card result per seat, two- and three-player
tables with all-in layers.  CPU: the oracle restatement and the host-instantiated product routine; GPU: the kernel."""
import ctypes as C
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import GOLDEN

GOLD = json.load(open(os.path.join(GOLDEN, "settle_run.json")))


def test_golden_covers_side_pots():
    assert len(GOLD) >= 1000
    assert sum(1 for g in GOLD if g["P"] == 3) >= 200
    layered = [g for g in GOLD if any(len({v for a, v in zip(acts, vals) if a == 2}) > 1 for acts, vals in zip(g["actions"], g["values"]))]
    assert len(layered) >= 200
    assert all(sum(g["won"]) == sum(g["bet"]) for g in GOLD)          # chip conservation in the reference


def test_oracle_restatement_matches_reference(oracle):
    from oracle import settle as S
    for g in GOLD:
        ranks = [oracle.rank7(g["board"] + h) for h in g["holes"]]
        won, res, card = S.settle(ranks, g["actions"], g["values"], g["onboard"], g["button"], g["small_blind"])
        assert (won, res, card) == (g["won"], g["game_result"], g["card_result"])


def _arrays(games, P):
    import texasholdempocker_rl_b200 as phe
    G = len(games)
    boards = phe.ids_to_masks(np.array([g["board"] for g in games]))
    holes = phe.ids_to_masks(np.array([g["holes"] for g in games]))
    act = np.zeros((G, 4, P), dtype=np.int8)
    val = np.zeros((G, 4, P), dtype=np.int32)
    for i, g in enumerate(games):
        R = len(g["actions"])
        act[i, :R] = g["actions"]
        val[i, :R] = g["values"]
    nr = np.array([len(g["actions"]) for g in games], dtype=np.int32)
    bt = np.array([g["button"] for g in games], dtype=np.int32)
    ob = np.array([sum(b << p for p, b in enumerate(g["onboard"])) for g in games], dtype=np.uint32)
    return boards, holes, act, val, nr, bt, ob


def test_host_instantiated_routine_matches_reference(hostcheck):
    hostcheck.hc_settle.argtypes = [C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for P in (2, 3):
        games = [g for g in GOLD if g["P"] == P]
        boards, holes, act, val, nr, bt, ob = _arrays(games, P)
        for i, g in enumerate(games):
            won = np.zeros(P, dtype=np.int32); res = np.zeros(P, dtype=np.int8); card = np.zeros(P, dtype=np.int8)
            paid = np.zeros(1, dtype=np.uint32)
            a = np.ascontiguousarray(act[i, :nr[i]]); v = np.ascontiguousarray(val[i, :nr[i]])
            h = np.ascontiguousarray(holes[i])
            hostcheck.hc_settle(int(boards[i]), h.ctypes.data, P, int(nr[i]), int(bt[i]), g["small_blind"], int(ob[i]), a.ctypes.data,
                                v.ctypes.data, won.ctypes.data, res.ctypes.data, card.ctypes.data, paid.ctypes.data)
            assert won.tolist() == g["won"] and res.tolist() == [int(x) for x in g["game_result"]]
            assert card.tolist() == [int(x) for x in g["card_result"]]


@pytest.mark.gpu
def test_gpu_settle_matches_reference(phe):
    import torch
    for P in (2, 3):
        games = [g for g in GOLD if g["P"] == P]
        arrs = _arrays(games, P)
        won, res, card, paid = phe.settle(*arrs, games[0]["small_blind"])
        assert won.tolist() == [g["won"] for g in games]
        assert res.tolist() == [[int(x) for x in g["game_result"]] for g in games]
        assert card.tolist() == [[int(x) for x in g["card_result"]] for g in games]
        dev = [torch.from_numpy(a.astype(np.int32) if a.dtype == np.uint32 else a).cuda() for a in arrs]
        won_d, res_d, card_d, paid_d = phe.settle(*dev, games[0]["small_blind"])
        assert (won_d.cpu().numpy() == won).all() and (res_d.cpu().numpy() == res).all() and (card_d.cpu().numpy() == card).all()
        assert (paid_d.cpu().numpy() == paid).all()


@pytest.mark.gpu
def test_gpu_settle_env_surface(phe):
    """settle_env on a reference-shaped object returns the dictionaries finish_game / _share_value build."""
    D = phe.DECK
    g = next(g for g in GOLD if g["P"] == 2 and len(g["actions"]) == 4 and max(g["won"]) > 0)
    names = ["player_0", "player_1"]
    FOLD, OTHER = SimpleNamespace(name="FOLD"), SimpleNamespace(name="RAISE")
    code = {0: None, 1: FOLD, 2: OTHER}
    env = SimpleNamespace(
        players={n: SimpleNamespace(status=SimpleNamespace(name="ONBOARD" if g["onboard"][p] else "ALLIN")) for p, n in enumerate(names)},
        button_player_name=names[g["button"]], small_blind=g["small_blind"],
        flop_cards=[D[i] for i in g["board"][:3]], turn_cards=[D[g["board"][3]]], river_cards=[D[g["board"][4]]],
        info_sets={n: SimpleNamespace(player_hand_cards=[D[i] for i in g["holes"][p]]) for p, n in enumerate(names)},
        historical_round_action_list=[{n: code[a[p]] for p, n in enumerate(names)} for a in g["actions"]],
        historical_round_value_list=[{n: v[p] for p, n in enumerate(names)} for v in g["values"]])
    won, result, card = phe.settle_env(env)
    assert {n: won.get(n, 0) for n in names} == {n: g["won"][p] for p, n in enumerate(names)}
    assert [result[n].value for n in names] == g["game_result"] and [card[n].value for n in names] == g["card_result"]
