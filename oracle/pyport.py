"""Pure-Python port of the reference's CPU hot path (TEST INFRASTRUCTURE / bench baseline ONLY).

What bench.py --impl reference times on the GPU box's host cores, where /root/reference does not exist:
the same algorithmic steps the reference executes per deal, in the reference's language --
  card -> id            utils/pokerhandevaluator_utils.py:3-4
  evaluate_cards        phevaluator (restated in oracle/shim/phevaluator)
  get_winner_set_v2     env/game.py:663-681
  generate_game_result  env/game.py:551-558
  cal_game_result       env/workflow.py:130-148
  plan recursion        env/workflow.py:150-199
Cards here are plain ids; the reference additionally pays for Card objects, Enum look-ups and sorted(), so this
port is a *faster* stand-in than the reference itself (noted in DESIGN.md).
"""
import os
import random
import sys

_here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_here, "shim"))
from phevaluator import evaluate_cards  # noqa: E402

LOSE, EVEN, WIN = -1.0, 0.0, 1.0


def get_winner_set_v2(flop, turn, river, player_cards):
    if len(player_cards) == 1:
        return set(player_cards)
    board = list(flop) + [turn[0], river[0]]
    lowest = sys.maxsize
    winners = set()
    for name, hand in player_cards.items():
        rank = evaluate_cards(*board, *hand)
        if rank <= lowest:
            if rank < lowest:
                lowest = rank
                winners.clear()
            winners.add(name)
    return winners


def generate_game_result(winner_set, result_dict):
    res = EVEN if len(winner_set) > 1 else WIN
    for name in winner_set:
        result_dict[name] = res


def cal_game_result(cards):
    hero, flop, turn, river, villain = sorted(cards[0:2]), sorted(cards[2:5]), cards[5:6], cards[6:7], sorted(cards[7:9])
    winners = get_winner_set_v2(flop, turn, river, {0: hero, 1: villain})
    result = {0: LOSE}
    generate_game_result(winners, result)
    return result[0]


def run_plan(plan, exist, start_idx=0, rng=random):
    """(game_result_sum, num_generated) for one task, restating generate_and_calculate_recurrent."""
    plan = list(plan)
    gen, nrand, seg_end = plan.pop(0)
    more = 0
    if not seg_end:
        for g, _, e in plan:
            more += g
            if e:
                break
    last = not plan
    exist_set = set(exist)
    total, n = 0.0, 0
    if gen == 1:
        for idx in range(start_idx, 52):
            if idx in exist_set:
                continue
            cards = exist + [idx]
            if last:
                total += cal_game_result(cards)
                n += 1
            elif more + idx < 52:
                s, k = run_plan(plan, cards, 0 if seg_end else idx + 1, rng)
                total += s
                n += k
    else:
        free = [i for i in range(52) if i not in exist_set]
        for _ in range(nrand):
            total += cal_game_result(exist + rng.sample(free, gen))
            n += 1
    return total, n


def eval_range(args):
    """Evaluate the colex combinations [begin, end); returns (count, 9-bin category histogram)."""
    begin, end = args
    from oracle import oracle as O
    c = [int(x) for x in O.colex_unrank7(begin)]
    bounds = O.K2_BOUNDS
    hist = [0] * 9
    for _ in range(end - begin):
        r = evaluate_cards(*c)
        k = 0
        while r > bounds[k]:
            k += 1
        hist[k] += 1
        i = 0
        while i < 6 and c[i] + 1 == c[i + 1]:
            c[i] = i
            i += 1
        c[i] += 1
    return end - begin, hist
