"""Pins the CPU oracle: known-answer vectors K1-K9, Philox KATs, and agreement of the restated
phevaluator algorithm with the first-principles best-of-21 evaluator."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

# (cards, rank): K3 upstream README examples, K4 independent brute force, K5 wheel
KATS = [
    ("9c 4c 4s 9d 4h Qc 6c", 292), ("9c 4c 4s 9d 4h 2c 9h", 236),
    ("As Ks Qs Js Ts 2c 3d", 1), ("5s 4s 3s 2s As Kd Kc", 10), ("Ac Ad Ah As Kc Kd Kh", 11),
    ("2c 2d 2h 2s 3c 3d 3h", 166), ("Ah Kh Qh Jh 9h 2c 3d", 323), ("7h 5h 4h 3h 2h 8c 9d", 1599),
    ("Ac Kd Qh Js 9c 2d 3h", 6186), ("Ac Ad 2h 7s 9c Jd 4h", 3435), ("As Ah 2c 7d Tc 3h 9s", 3463),
    ("2c 3d 4h 5s 7c 8d 9h", 7414), ("Ad 2h 3c 4d 5h Ks 9c", 1609),
]


def test_class_count(oracle):
    assert oracle.lib().orc_num_classes() == 7462


@pytest.mark.parametrize("text,rank", KATS)
def test_kat_ranks(oracle, text, rank):
    ids = oracle.ids(text)
    assert oracle.rank7(ids) == rank
    assert oracle.rank7(ids, brute=True) == rank


def test_kat_ids_match_survey(oracle):
    assert oracle.ids("As Ks Qs Js Ts 2c 3d") == [51, 47, 43, 39, 35, 0, 5]
    assert oracle.ids("Ad 2h 3c 4d 5h Ks 9c") == [49, 2, 4, 9, 14, 47, 28]


def test_restated_equals_brute_force_random(oracle):
    rng = np.random.default_rng(11)
    hands = np.argsort(rng.random((100000, 52)), axis=1)[:, :7].astype(np.uint8)
    assert (oracle.rank7_batch(hands) == oracle.rank7_batch(hands, brute=True)).all()


def test_five_card_category_bounds(oracle):
    # K2: weakest member of each category sits exactly on the bound
    probes = {"2c 3c 4c 5c Ac": 10, "2c 2d 2h 2s 3c": 166, "2c 2d 2h 3c 3d": 322, "2c 3c 4c 5c 7c": 1599,
              "Ac 2d 3h 4s 5c": 1609, "2c 2d 2h 3c 4d": 2467, "2c 2d 3c 3d 4h": 3325, "2c 2d 3c 4d 5h": 6185,
              "2c 3d 4h 5s 7c": 7462}
    for text, want in probes.items():
        assert oracle.rank5_brute(oracle.ids(text)) == want, text


def test_exhaustive_histogram_k1(oracle):
    h9, h = oracle.enum7()
    assert h9.tolist() == oracle.K1_HIST9
    assert int(h.sum()) == oracle.NUM_HANDS_7 and h[0] == 0 and int(h[7415:].sum()) == 0
    gold = json.load(open(os.path.join(GOLDEN, "hist7463.json")))
    assert h.tolist() == gold["hist7463"]
    lo = 1
    for cat, hi in enumerate(oracle.K2_BOUNDS):
        assert int(h[lo:hi + 1].sum()) == oracle.K1_HIST9[cat]
        lo = hi + 1


def test_enum_shards_add_up(oracle):
    a9, a = oracle.enum7(0, 3_000_000, nthreads=2)
    b9, b = oracle.enum7(0, 1_234_567, nthreads=1)
    c9, c = oracle.enum7(1_234_567, 3_000_000, nthreads=3)
    assert (a == b + c).all() and (a9 == b9 + c9).all()


def test_philox_kats(oracle):
    # Random123 known-answer vectors for philox4x32-10
    assert oracle.philox4x32_10([0] * 4, [0] * 2).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_k9_exact_equities(oracle):
    assert oracle.equity_plan(oracle.ids("As Ah 2c 7d Tc 3h 9s"), oracle.DEFAULT_PLANS[3][0]) == (852, 1, 137, 990)
    assert oracle.equity_plan(oracle.ids("7h 8h 9h Th 2c 2d Ks"), oracle.DEFAULT_PLANS[3][0]) == (0, 201, 789, 990)
    assert oracle.equity_plan(oracle.ids("As Ah 2c 7d Tc 3h"), oracle.DEFAULT_PLANS[2][0]) == (39837, 44, 5659, 45540)


def test_k6_plan_counts(oracle):
    for rnd, text in enumerate(["As Ah", "As Ah 2c 7d Tc", "As Ah 2c 7d Tc 3h", "As Ah 2c 7d Tc 3h 9s"]):
        for plans in (oracle.DEFAULT_PLANS, oracle.DEBUG_PLANS):
            plan, n = plans[rnd]
            w, t, l, got = oracle.equity_plan(oracle.ids(text), plan, seed=3)
            assert got == n and w + t + l == n


def test_k8_aces_preflop(oracle):
    h = oracle.mask_of(oracle.ids("As Ah"))
    for fn in (lambda: oracle.equity_mc(h, h, 1, 400000, seed=1), lambda: oracle.equity_mc_indep(h, h, 1, 400000, seed=1)):
        w, t, l = fn()
        p = (w + t / 2) / (w + t + l)
        assert abs(p - 0.852) < 3 * (0.25 / 400000) ** 0.5 + 0.001


@pytest.mark.parametrize("need", [1, 4, 5])
def test_shared_stream_is_uniform(oracle, need):
    # every unknown card is dealt equally often, for even counts (pairs only) and odd ones (single + pairs)
    known_ids = oracle.ids("As Ah 2c 7d Tc")
    known = oracle.mask_of(known_ids)
    counts = np.zeros(52)
    n = 40000
    for s in range(n):
        row = oracle.deal(known, 99, s, 0, need)
        assert len(set(row.tolist())) == need
        for c in row:
            counts[c] += 1
    assert counts[known_ids].sum() == 0
    free = np.delete(counts, known_ids)
    exp = n * need / 47
    chi2 = ((free - exp) ** 2 / exp).sum()
    assert chi2 < 46 + 4 * (2 * 46) ** 0.5


def test_shared_stream_pairs_are_uniform(oracle):
    # the two cards of a pair slot are a uniform unordered pair of the free cards
    known = oracle.mask_of(oracle.ids("As Ah 2c 7d Tc 3h 9s"))
    seen = {}
    n = 60000
    for s in range(n):
        a, b = oracle.deal(known, 5, s, 1, 2).tolist()
        assert a < b
        seen[(a, b)] = seen.get((a, b), 0) + 1
    assert len(seen) == 990
    exp = n / 990
    chi2 = sum((v - exp) ** 2 / exp for v in seen.values())
    assert chi2 < 989 + 4 * (2 * 989) ** 0.5


def test_winner_mask_semantics(oracle):
    board = oracle.ids("2c 7d Tc 3h 9s")
    holes = [oracle.ids("As Ah"), oracle.ids("Kd Kc"), oracle.ids("Ad Ac")]
    assert oracle.winner_mask(board, holes) == 0b101            # two players tie with aces
    assert oracle.winner_mask(board, holes, live=0b010) == 0b010  # lone player wins unevaluated
    assert oracle.winner_mask(board, holes, live=0b110) == 0b100
