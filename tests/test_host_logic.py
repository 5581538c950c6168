"""Host-side logic that needs no GPU: card contract, state packing, task parsing, shard ranges."""
import numpy as np
import pytest


def test_card_contract(phe):
    # utils/pokerhandevaluator_utils.py:3-4 and env/game.py:14-18: deck[i] has id i
    for i, c in enumerate(phe.DECK):
        assert phe.card_to_evaluator(c) == i == phe.card_id(c)
    assert phe.card_id("As") == 51 and phe.card_id("2c") == 0 and phe.card_id("Td") == 33
    assert str(phe.DECK[51]) == "SA"
    assert sorted([phe.DECK[5], phe.DECK[4], phe.DECK[50]]) == [phe.DECK[4], phe.DECK[5], phe.DECK[50]]
    with pytest.raises(ValueError):
        phe.card_id(52)
    with pytest.raises(ValueError):
        phe.card_id("1x")


def test_masks_roundtrip(phe):
    rng = np.random.default_rng(0)
    ids = np.argsort(rng.random((100, 52)), axis=1)[:, :7]
    m = phe.ids_to_masks(ids)
    for row, mask in zip(ids, m):
        assert sorted(phe.mask_to_ids(mask)) == sorted(row.tolist())
        assert phe.cards_to_mask(row.tolist()) == int(mask)
    with pytest.raises(ValueError):
        phe.cards_to_mask([1, 2, 2])
    with pytest.raises(ValueError):
        phe.ids_to_masks(np.array([[1, 1, 2, 3, 4, 5, 6]]))


def test_evaluate_cards_argument_errors(phe):
    with pytest.raises(ValueError):
        phe.evaluate_cards(1, 2, 3)
    with pytest.raises(ValueError):
        phe.evaluate_cards(1, 2, 3, 4, 5, 6, 6)


def test_pack_states(phe):
    s = phe.pack_states([["As", "Ah"]], [["2c", "7d", "Tc"]], n_opp=5)
    assert s.shape == (1, 2)
    word = int(s[0, 1]) & (2**64 - 1)
    hero = word & ((1 << 61) - 1)
    assert word >> 61 == 5 and bin(hero).count("1") == 2 and bin(int(s[0, 0])).count("1") == 5
    with pytest.raises(ValueError):
        phe.pack_states([["As", "Ah"]], [["As", "7d", "Tc"]])
    with pytest.raises(ValueError):
        phe.pack_states([["As", "Ah"]], [["7d", "Tc"]])


def test_random_states_shape(phe):
    s = phe.random_states(4096, 20261019)
    pc = np.bitwise_count(s[:, 0].astype(np.uint64))
    assert set(np.unique(pc).tolist()) == {2, 5, 6, 7}
    assert (np.bitwise_count(s[:, 1].astype(np.uint64)) == 2).all() and ((s[:, 0] & s[:, 1]) == s[:, 1]).all()
    assert (s == phe.random_states(4096, 20261019)).all()


def test_task_parsing(phe):
    from importlib import import_module
    svc = import_module("texasholdempocker-rl_b200.service")
    D = phe.DECK
    t = ("g", "p0", 2, [D[0], D[21], D[32]], [D[10]], None, [D[51], D[50]])
    assert svc.task_exist_ids(t) == [51, 50, 0, 21, 32, 10]
    assert svc.task_exist_ids(("g", "p0", 0, None, None, None, [D[51], D[50]])) == [51, 50]
    with pytest.raises(ValueError):
        svc.task_exist_ids(("g", "p0", 1, [D[0], D[21], D[51]], None, None, [D[51], D[50]]))
    assert svc.winning_probability(715.0, 990) == pytest.approx(0.8611111111)


def test_generate_game_result(phe):
    R = phe.GamePlayerResult
    d = {0: R.LOSE, 1: R.LOSE}
    phe.GameEnv.generate_game_result({0}, d)
    assert d == {0: R.WIN, 1: R.LOSE}
    d = {0: R.LOSE, 1: R.LOSE, 2: R.LOSE}
    phe.GameEnv.generate_game_result({0, 2}, d)
    assert d == {0: R.EVEN, 1: R.LOSE, 2: R.EVEN}
    assert phe.GameEnv.get_winner_set_v2(None, None, None, {"solo": None}) == {"solo"}   # env/game.py:664-665


def test_shard_range_tiles(phe):
    for total in (0, 1, 7, 133784560, 10**9):
        for world in (1, 2, 3, 8):
            edges = [phe.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_equity_mc_rejects_malformed_numpy_states(phe):
    good = phe.pack_states([["As", "Ah"]], [["2c", "7d", "Tc"]])
    for bad in (np.array([[good[0, 0], 0]]), np.array([[good[0, 0] | (1 << 13), good[0, 1]]]),
                np.array([[good[0, 1], good[0, 0]]]), np.array([[good[0, 0] | 0xF, good[0, 1]]])):
        with pytest.raises(ValueError):
            phe.equity_mc(bad.astype(np.int64), 1, 10, seed=1)
    with pytest.raises(ValueError):
        phe.equity_mc(good, 9, 10, seed=1)
