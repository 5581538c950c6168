"""Generates tests/golden/settle_run.json by playing random legal games in the REFERENCE'S OWN GameEnv
(/root/reference/env/game.py, unmodified; phevaluator shim as in make_golden.py) to the terminal state and
recording the inputs and outputs of finish_game() / _share_value() (env/game.py:410-441, 560-632, 683-734).

  python tests/golden/make_settle_golden.py       (build container only)

settle_run.json: two- and three-player tables; settle_run_multi.json.gz: four to six seats (BASELINE configs 3 and 5 are
six-max).  Stacks carry over between hands so that all-in layers (side pots) occur.  A hand whose settlement does not
return within one second is dropped: for more than two players the reference's odd-chip loop (env/game.py:731-733) does
not terminate when neither the button nor the next ONBOARD seat after it is a winner; so is a hand
on which it divides by an empty winner set (env/game.py:721).  Those terminals are covered by the property test of
tests/test_settlement.py against oracle/settle.py, which documents the two fixes."""
import json
import os
import random
import signal
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")

from env.game import GameEnv  # noqa: E402
from env.constants import PlayerActions, PlayerStatus, GET_PLAYER_NAME  # noqa: E402
from utils.pokerhandevaluator_utils import card_to_evaluator  # noqa: E402

rng = random.Random(7)
ACT = {None: 0, PlayerActions.FOLD: 1}


class Timeout(Exception):
    pass


def _alarm(signum, frame):
    raise Timeout()


signal.signal(signal.SIGALRM, _alarm)


def random_action(env):
    name = env.acting_player_name
    pl = env.players[name]
    hist = env.current_round_value_dict[name]
    to_call = env.current_round_min_value - hist
    sb = env.small_blind
    r = rng.random()
    if r < 0.12:
        return (PlayerActions.FOLD, 0)
    if pl.value_left <= to_call:
        return (PlayerActions.CHECK_CALL, pl.value_left + hist)                # short stack: all-in call (env/game.py:1010-1011)
    if r < 0.55:
        return (PlayerActions.CHECK_CALL, env.current_round_min_value)
    min_raise_to = env.current_round_min_value + env.current_round_raise_min_value
    max_to = hist + pl.value_left
    if min_raise_to >= max_to or r > 0.9:
        return (PlayerActions.RAISE, max_to)                                   # all-in
    steps = (max_to - min_raise_to) // sb
    target = min_raise_to + sb * min(steps, int(rng.expovariate(0.15)))
    return (PlayerActions.RAISE, int(target))


crashed = []


def play(tables):
    records = []
    dropped = 0
    for table in tables:
        P, n_hands, stack = table[:3]
        keep_full = len(table) > 3 and table[3]
        env = GameEnv(P, init_value=stack, small_blind=25, big_blind=50, settle_automatically=False)
        for h in range(n_hands):
            if keep_full and env.players is not None and (len(env.players) < P or any(pl.value_left <= 0 for pl in env.players.values())):
                env = GameEnv(P, init_value=stack, small_blind=25, big_blind=50, settle_automatically=False)      # busted seats left: new table
            env.reset(seed=rng.randrange(1 << 30))
            guard = 0
            while not env.game_over and guard < 200:
                env.step(random_action(env), 0)
                guard += 1
            if not env.game_over:
                continue
            names = list(env.players.keys())
            rec = {
                "P": len(names), "small_blind": env.small_blind, "button": names.index(env.button_player_name),
                "board": [card_to_evaluator(c) for c in env.flop_cards + env.turn_cards + env.river_cards],
                "holes": [[card_to_evaluator(c) for c in env.info_sets[n].player_hand_cards] for n in names],
                "actions": [[ACT.get(a.get(n), 2) for n in names] for a in env.historical_round_action_list],
                "values": [[int(v.get(n, 0)) for n in names] for v in env.historical_round_value_list],
                "onboard": [int(env.players[n].status == PlayerStatus.ONBOARD) for n in names],
                "bet": [int(env.players[n].value_bet) for n in names],
            }
            signal.setitimer(signal.ITIMER_REAL, 1.0)
            try:
                won = env.finish_game()
                signal.setitimer(signal.ITIMER_REAL, 0)
            except Timeout:
                dropped += 1
                env.reset(force_restart_game=True)
                continue
            except ZeroDivisionError:
                # a pot left without a contender (every seat still in the hand was settled as an all-in layer):
                # share_value_to_winner divides by an empty winner set (env/game.py:721); only seen at four seats and more
                signal.setitimer(signal.ITIMER_REAL, 0)
                dropped += 1
                crashed.append(rec)
                env = GameEnv(P, init_value=stack, small_blind=25, big_blind=50, settle_automatically=False)
                continue
            rec["won"] = [int(won.get(n, 0)) for n in names]
            rec["game_result"] = [env.players[n].game_result.value for n in names]
            rec["card_result"] = [env.players[n].card_result.value for n in names]
            assert sum(rec["won"]) == sum(rec["bet"]), "reference settlement lost chips"
            records.append(rec)
    return records, dropped


def layered(records):
    return sum(1 for r in records if any(sum(1 for a in acts if a == 2) >= 2 and len(set(v for a, v in zip(acts, vals) if a == 2)) > 1 for acts, vals in zip(r["actions"], r["values"])))


records, dropped = play(((2, 700, 2000), (3, 500, 1500), (2, 300, 100000)))
json.dump(records, open(os.path.join(ROOT, "tests", "golden", "settle_run.json"), "w"))
print(len(records), "terminal states;", layered(records), "with unequal live stakes in some round;", dropped, "dropped (reference did not terminate)")
# four to six seats (BASELINE configs 3 / 5 are six-max): the same recipe; terminals on which the reference's odd-chip loop
# does not terminate are dropped here and covered by the property test in tests/test_settlement.py instead
rng.seed(8)
# (seats, hands, stack, restart the table whenever a seat is busted); without restarts the table shrinks and stacks diverge: side pots
records, dropped = play(((4, 200, 1500, True), (5, 200, 1500, True), (6, 300, 1500, True), (4, 250, 1500), (5, 300, 1500), (6, 500, 1500), (6, 300, 600)))
import gzip  # noqa: E402
with gzip.GzipFile(os.path.join(ROOT, "tests", "golden", "settle_run_multi.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps(records, separators=(",", ":")).encode())
print(len(crashed), "terminals on which the reference raised ZeroDivisionError (a pot without a contender; dropped, see tests/test_settlement.py)")
print(len(records), "terminal states at 4-6 seats;", layered(records), "with unequal live stakes in some round;", dropped, "dropped (reference did not terminate)",
      {P: sum(1 for r in records if r["P"] == P) for P in (4, 5, 6)})
