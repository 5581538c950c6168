"""Generates tests/golden/redeal_run.json: empirical distribution of the REFERENCE'S OWN GameEnv.new_random()
(env/game.py:81-165, unmodified; phevaluator shim as in make_golden.py) at each street of a heads-up and a
three-player hand: how often each card lands in an opponent's hand / on the part of the board that was still hidden,
plus the invariants (acting player's hand and the revealed board are kept).  Build container only."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")

from env.game import GameEnv  # noqa: E402
from env.constants import PlayerActions  # noqa: E402
from utils.pokerhandevaluator_utils import card_to_evaluator as cid  # noqa: E402

random.seed(12345)
N = 12000
out = []
for P in (2, 3):
    env = GameEnv(P, init_value=100000, small_blind=25, big_blind=50, settle_automatically=False)
    env.reset(seed=99 + P)
    for rnd in range(4):
        while env.current_round < rnd and not env.game_over:      # everybody calls until the street is reached
            env.step((PlayerActions.CHECK_CALL, env.current_round_min_value), 0)
        assert env.current_round == rnd and not env.game_over
        acting = env.acting_player_name
        hero = [cid(c) for c in env.info_sets[acting].player_hand_cards]
        board_known = ([cid(c) for c in env.flop_cards] if rnd >= 1 else []) + ([cid(c) for c in env.turn_cards] if rnd >= 2 else []) + \
                      ([cid(c) for c in env.river_cards] if rnd >= 3 else [])
        opp_counts, board_counts = [0] * 52, [0] * 52
        for _ in range(N):
            e2 = env.new_random()
            assert [cid(c) for c in e2.info_sets[acting].player_hand_cards] == hero
            full_board = [cid(c) for c in e2.flop_cards + e2.turn_cards + e2.river_cards]
            assert full_board[:len(board_known)] == board_known
            seen = set(hero) | set(full_board)
            for name, info in e2.info_sets.items():
                if name != acting:
                    h = [cid(c) for c in info.player_hand_cards]
                    assert h == sorted(h) and not (set(h) & seen)
                    seen |= set(h)
                    for c in h:
                        opp_counts[c] += 1
            assert len(seen) == 5 + 2 * P
            for c in full_board[len(board_known):]:
                board_counts[c] += 1
        out.append({"P": P, "round": rnd, "hero": hero, "board_known": board_known, "n": N, "opp_counts": opp_counts, "board_counts": board_counts})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "redeal_run.json"), "w"))
print(len(out), "situations x", N, "re-deals")
