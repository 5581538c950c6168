"""Restatement of the reference's terminal settlement (TEST INFRASTRUCTURE ONLY).

  finish_game                 env/game.py:410-441   (card_result: pure hand comparison over ALL players)
  _share_value                env/game.py:560-632   (per betting round: all-in layers in ascending stake, then main pot)
  share_pot_for_winner        env/game.py:702-716
  share_value_to_winner       env/game.py:718-734   (floor to small-blind multiples, odd chips toward the button)

Inputs are plain arrays (seat order = the reference's players dict order):
  ranks[P]        7-card rank of every player (lower wins)
  actions[R][P]   last action of the round: 0 None, 1 FOLD, 2 anything else (blind, check/call, raise)
  values[R][P]    chips put in during the round
  onboard[P]      status == ONBOARD at the end (used only by the odd-chip rule)
Quirks kept on purpose: a player whose action in a round is None (all-in earlier -- or folded earlier, for more than
two players) still contends in that round's showdowns (the reference tests `!= FOLD`, env/game.py:599, 625); the game
result is the one written by the LAST showdown a player won (:553).  Two fixes, both outside what the reference can
compute (tables of more than two seats): (1) where its odd-chip loop would not terminate (:731-733) the chips go to
the first winner clockwise from the button; (2) a final pot that is left without a contender -- every seat still in
the hand was already settled as an all-in layer, the reference divides by an empty winner set (:721) -- goes to the
best hand among the seats that have not folded.
"""
from math import floor

LOSE, EVEN, WIN = -1.0, 0.0, 1.0


def winner_set(ranks, contenders):
    contenders = list(contenders)
    if len(contenders) == 1:
        return set(contenders)
    best = min(ranks[p] for p in contenders)
    return {p for p in contenders if ranks[p] == best}


def _mark(winners, result):
    res = EVEN if len(winners) > 1 else WIN
    for p in winners:
        result[p] = res


def _next_onboard(button, onboard):
    P = len(onboard)
    q = (button + 1) % P
    while not onboard[q]:
        q = (q + 1) % P
        if q == button:
            return None
    return q


def _pay(total, winners, won, small_blind, button, onboard):
    if total <= 0:
        return
    left = total
    share = floor((total / len(winners)) // small_blind) * small_blind
    for p in winners:
        won[p] = int(won[p] + share)
        left -= share
    if left > 0:
        who = button
        if who not in winners:
            who = _next_onboard(button, onboard)
            if who not in winners:          # the reference spins forever here; documented fix
                P = len(onboard)
                who = next((button + k) % P for k in range(1, P + 1) if (button + k) % P in winners)
        won[who] = int(won[who] + left)


def settle(ranks, actions, values, onboard, button, small_blind):
    P = len(ranks)
    players = range(P)
    result = [LOSE] * P
    won = [0] * P
    pot = 0
    all_shared = set()
    for act, val in zip(actions, values):
        round_shared = set()
        top = max(val)
        allin = sorted(((p, val[p]) for p in players if act[p] == 2 and val[p] < top), key=lambda t: t[1])
        spare = {p: val[p] for p in players}
        for a, _ in allin:
            if a not in spare or spare[a] <= 0:
                round_shared.add(a); all_shared.add(a)
                continue
            winners = winner_set(ranks, [p for p in players if act[p] != 1 and p not in round_shared])
            _mark(winners, result)
            if a in winners:
                stake = spare[a]
                total = pot
                for p in list(spare):
                    if spare[p] > stake:
                        spare[p] -= stake; total += stake
                    else:
                        total += spare.pop(p)
                _pay(total, winners, won, small_blind, button, onboard)
                pot = 0
            else:
                pot += spare.pop(a)
            round_shared.add(a); all_shared.add(a)
        pot += sum(spare.values())
    if pot > 0:
        last = actions[-1]
        contenders = [p for p in players if last[p] != 1 and p not in all_shared]
        if not contenders:                  # the reference divides by an empty winner set here (:721); documented fix 2
            contenders = [p for p in players if last[p] != 1]
        winners = winner_set(ranks, contenders)
        _mark(winners, result)
        _pay(pot, winners, won, small_blind, button, onboard)
    cw = winner_set(ranks, players) if P > 1 else {0}
    card = [(WIN if len(cw) == 1 else EVEN) if p in cw else LOSE for p in players]
    return won, result, card
