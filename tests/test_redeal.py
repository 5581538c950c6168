"""Row a8 -- GameEnv.new_random (env/game.py:81-165).  The reference's re-deal is an unseeded random.sample, so parity is
distributional: the golden file holds the empirical card frequencies of the reference's own new_random() at every street
(tests/golden/make_redeal_golden.py); the oracle stream and the CUDA deal must be uniform over the same unknown cards and
statistically indistinguishable from the reference's frequencies."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

GOLD = json.load(open(os.path.join(GOLDEN, "redeal_run.json")))


def chi2_uniform(counts, known):
    free = np.delete(np.asarray(counts, dtype=np.float64), known)
    exp = free.sum() / len(free)
    return ((free - exp) ** 2 / exp).sum(), len(free) - 1


def test_reference_redeal_is_uniform_over_unknown_cards():
    for g in GOLD:
        known = g["hero"] + g["board_known"]
        assert sum(g["opp_counts"][c] for c in known) == 0 and sum(g["board_counts"][c] for c in known) == 0
        assert sum(g["opp_counts"]) == g["n"] * 2 * (g["P"] - 1)
        assert sum(g["board_counts"]) == g["n"] * (5 - len(g["board_known"]))
        for counts in (g["opp_counts"], g["board_counts"]):
            if sum(counts):
                c2, dof = chi2_uniform(counts, known)
                assert c2 < dof + 4.5 * (2 * dof) ** 0.5


@pytest.mark.parametrize("gi", range(len(GOLD)))
def test_oracle_stream_matches_reference_distribution(oracle, gi):
    g = GOLD[gi]
    known_ids = g["hero"] + g["board_known"]
    known = oracle.mask_of(known_ids)
    n_opp, nbm = g["P"] - 1, 5 - len(g["board_known"])
    need = 2 * n_opp + nbm
    n = 6000
    opp, board = np.zeros(52), np.zeros(52)
    for s in range(n):
        ids = oracle.deal(known, 7, s, gi, need)
        odd = need & 1
        for c in ids[odd:odd + 2 * n_opp]:
            opp[c] += 1
        for c in list(ids[:odd]) + list(ids[odd + 2 * n_opp:]):
            board[c] += 1
    for mine, ref in ((opp, g["opp_counts"]), (board, g["board_counts"])):
        if mine.sum() == 0:
            continue
        c2, dof = chi2_uniform(mine, known_ids)
        assert c2 < dof + 4.5 * (2 * dof) ** 0.5
        # two-sample chi-square against the reference's frequencies
        a, b = np.delete(mine, known_ids), np.delete(np.asarray(ref, dtype=np.float64), known_ids)
        k1, k2 = (b.sum() / a.sum()) ** 0.5, (a.sum() / b.sum()) ** 0.5
        c2 = ((k1 * a - k2 * b) ** 2 / (a + b)).sum()
        assert c2 < dof + 4.5 * (2 * dof) ** 0.5


@pytest.mark.gpu
@pytest.mark.parametrize("gi", range(len(GOLD)))
def test_gpu_determinise_matches_reference_distribution(phe, oracle, gi):
    g = GOLD[gi]
    known_ids = g["hero"] + g["board_known"]
    n = 40000
    opp, board = phe.determinise(g["hero"], g["board_known"], g["P"], n, seed=11 + gi)
    assert opp.shape == (n, g["P"] - 1, 2) and board.shape == (n, 5)
    assert (board[:, :len(g["board_known"])] == np.array(g["board_known"], dtype=np.uint8)).all()       # revealed cards kept
    assert (opp[:, :, 0] < opp[:, :, 1]).all()                                                           # hands sorted like the reference
    allc = np.concatenate([opp.reshape(n, -1), board, np.tile(np.array(g["hero"], dtype=np.uint8), (n, 1))], axis=1)
    assert (np.sort(allc, axis=1)[:, 1:] != np.sort(allc, axis=1)[:, :-1]).all()                         # no card twice
    if len(g["board_known"]) == 0:
        assert (np.sort(board[:, :3], axis=1) == board[:, :3]).all()                                     # new flop sorted
    for mine, ref in ((np.bincount(opp.reshape(-1), minlength=52), g["opp_counts"]),
                      (np.bincount(board[:, len(g["board_known"]):].reshape(-1), minlength=52), g["board_counts"])):
        if mine.sum() == 0:
            continue
        c2, dof = chi2_uniform(mine, known_ids)
        assert c2 < dof + 4.5 * (2 * dof) ** 0.5
        a, b = np.delete(mine.astype(np.float64), known_ids), np.delete(np.asarray(ref, dtype=np.float64), known_ids)
        k1, k2 = (b.sum() / a.sum()) ** 0.5, (a.sum() / b.sum()) ** 0.5
        assert ((k1 * a - k2 * b) ** 2 / (a + b)).sum() < dof + 4.5 * (2 * dof) ** 0.5
    # the first rows are the oracle's deals for (seed, stream row, sample 0)
    known = oracle.mask_of(known_ids)
    need = 2 * (g["P"] - 1) + 5 - len(g["board_known"])
    raw = phe.deal(np.full(4, known, dtype=np.int64), need, 11 + gi)
    for i in range(4):
        assert (raw[i] == oracle.deal(known, 11 + gi, 0, i, need)).all()


@pytest.mark.gpu
def test_gpu_determinise_streams_are_disjoint_and_follow_the_oracle(phe, oracle):
    """Re-deal i of determinise(stream=s) is sample (offset + i) of Philox stream s (phe_deal_samples): equal to the oracle's
    deal for that (stream, sample), and two callers numbering their trees s and s + 1 share no (stream, sample) pair."""
    hero, shown = ["As", "Ah"], ["2c", "7d", "Tc"]
    known = phe.cards_to_mask(hero + shown)
    a = phe.deal(np.full(64, known, dtype=np.int64), 4, 5, sample_offset=10, stream_base=7, by_sample=True)
    b = phe.deal(np.full(64, known, dtype=np.int64), 4, 5, sample_offset=10, stream_base=8, by_sample=True)
    for i in range(0, 64, 7):
        assert (a[i] == oracle.deal(known, 5, 10 + i, 7, 4)).all()
        assert (b[i] == oracle.deal(known, 5, 10 + i, 8, 4)).all()
    assert (a != b).any(axis=1).sum() >= 60                       # adjacent streams: independent deals, not a shifted copy
    assert not any((a[i + 1] == b[i]).all() for i in range(63))  # (the old row = stream numbering made these identical)
    o1, b1 = phe.determinise(hero, shown, 2, 16, seed=5, sample_offset=10, stream=7)
    o2, b2 = phe.determinise(hero, shown, 2, 16, seed=5, sample_offset=26, stream=7)
    o3, b3 = phe.determinise(hero, shown, 2, 32, seed=5, sample_offset=10, stream=7)
    assert (np.concatenate([o1, o2]) == o3).all() and (np.concatenate([b1, b2]) == b3).all()       # waves tile the sample axis
