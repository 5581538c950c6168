"""Generates tests/golden/obs_run.json.gz: observations produced by the REFERENCE'S OWN Env.get_obs
(/root/reference/env/env.py:416-581, unmodified) while random legal heads-up hands are played through its Env /
GameEnv, together with the raw infoset quantities the encoder reads (env/game.py:1053-1092).

  python tests/golden/make_obs_golden.py       (build container only)

Every decision point is recorded twice: as seen by the acting player and, with mcts_acting_player_name set to the
other seat, as seen inside an MCTS simulation where the acting player's hole cards are hidden (env/env.py:445-448).
Stacks carry over between hands so that the ratio features leave their first bins."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")

from env.env import Env  # noqa: E402
from env.constants import ALL_PLAYER_NAMES, GET_PLAYER_ID_BY_NAME  # noqa: E402
from utils.pokerhandevaluator_utils import card_to_evaluator  # noqa: E402
from utils.workflow_utils import map_action_bin_to_actual_action_and_value_v2  # noqa: E402

random.seed(20261020)
NUM_BINS, HIST = 12, 3


def ids(cards):
    return None if cards is None else [card_to_evaluator(c) for c in cards]


def record(env, infoset, hidden):
    names = ALL_PLAYER_NAMES
    me = infoset.player_name
    other = [n for n in names if n != me][0]
    obs = env.get_obs(infoset, mcts_acting_player_name=other if hidden else None)
    st = infoset.player_status_value_left_bet_dict
    return {
        "round": int(infoset.current_round), "player": GET_PLAYER_ID_BY_NAME(me), "button": GET_PLAYER_ID_BY_NAME(infoset.button_player_name),
        "num_players": int(infoset.num_players), "hidden": bool(hidden),
        "hand": ids(infoset.player_hand_cards), "flop": ids(infoset.flop_cards), "turn": ids(infoset.turn_cards), "river": ids(infoset.river_cards),
        "game_init": infoset.game_init_value, "pot": infoset.pot_value, "bins": list(infoset.bin_bet_value_list),
        "init": [infoset.player_value_init_dict[n] for n in names],
        "status": [st[n][0].value for n in names], "left": [st[n][1] for n in names], "bet": [st[n][2] for n in names],
        "history": [[list(r) for r in infoset.all_player_round_action_bins_dict.get(n, [])] for n in names],
        "obs": [int(v) for v in obs],
    }


records = []
for stack, hands in ((100_000, 60), (3_000, 120), (1_000, 120)):
    env = Env(None, NUM_BINS, HIST, init_value=stack, small_blind=25, big_blind=50, ignore_all_async_tasks=True)
    for h in range(hands):
        try:
            env.reset(h, seed=random.randrange(1 << 30))
        except Exception:
            env = Env(None, NUM_BINS, HIST, init_value=stack, small_blind=25, big_blind=50, ignore_all_async_tasks=True)
            env.reset(h, seed=random.randrange(1 << 30))
        guard = 0
        while not env._env.game_over and guard < 100:
            infoset = env._game_infoset
            records.append(record(env, infoset, False))
            records.append(record(env, infoset, True))
            mask, ranges, left, hist_value, dmin = env.get_valid_action_info_v2()
            legal = [i for i, m in enumerate(mask) if not m]
            w = [0.3 if i == 0 else (3.0 if i == 1 else 1.0) for i in legal]
            b = random.choices(legal, weights=w)[0]
            action = map_action_bin_to_actual_action_and_value_v2(b, ranges, left, hist_value, dmin, 25)
            env.step(action, b)
            guard += 1
print(len(records), "observations")
types = {k: sorted({type(r[k]).__name__ for r in records}) for k in ("game_init", "pot")}
print(types, sorted({type(v).__name__ for r in records for v in r["bins"]}), sorted({type(v).__name__ for r in records for v in r["left"] + r["bet"] + r["init"]}))
import gzip  # noqa: E402
path = os.path.join(ROOT, "tests", "golden", "obs_run.json.gz")
with gzip.GzipFile(path, "wb", mtime=0) as f:
    f.write(json.dumps(records, separators=(",", ":")).encode())
print("wrote", path, os.path.getsize(path), "bytes")
