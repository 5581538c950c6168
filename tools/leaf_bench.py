"""BASELINE config 4, the part of it on this path: what one lock-step wave of B tree simulations asks of the device.

  python tools/leaf_bench.py [B] [debug|default]

Per wave (everything stays on the device, one stream):
  1. re-deal       B determinisations (GameEnv.new_random, env/game.py:119-163)            -> phe_deal
  2. leaf encode   B decision points -> int32[B, 93] (Env.get_obs, env/env.py:416-581)     -> phe_encode_obs
  3. leaf predict  B observations -> action probabilities + value (rl/MCTS.py:464-473)     -> BatchedPredictor
  4. terminals     B finished hands -> winners, chips, results (env/game.py:410-441,560-734) -> phe_settle
Decision points and finished hands are the reference's own (tests/golden/obs_run.json.gz, settle_run.json), replicated
to B; results of a wave are checked against the goldens before timing.  The tree (selection / backup) and the betting
step are the caller's and are not part of this measurement."""
import gzip
import json
import os
import sys
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import texasholdempocker_rl_b200 as phe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
which = sys.argv[2] if len(sys.argv) > 2 else "debug"
pv = phe.policy_value
torch.cuda.set_device(0)
dev = torch.device("cuda:0")

# --- inputs ---
recs = json.load(gzip.open(os.path.join(ROOT, "tests", "golden", "obs_run.json.gz")))
names = ["player_0", "player_1"]


def infoset_of(r):
    return SimpleNamespace(
        player_name=names[r["player"]], button_player_name=names[r["button"]], num_players=2, current_round=r["round"],
        player_hand_cards=r["hand"], flop_cards=r["flop"], turn_cards=r["turn"], river_cards=r["river"], game_init_value=r["game_init"],
        pot_value=r["pot"], bin_bet_value_list=r["bins"], player_value_init_dict={n: r["init"][k] for k, n in enumerate(names)},
        player_status_value_left_bet_dict={n: (r["status"][k], r["left"][k], r["bet"][k]) for k, n in enumerate(names)},
        all_player_round_action_bins_dict={n: r["history"][k] for k, n in enumerate(names) if r["history"][k]})


enc = phe.ObsEncoder()
base = enc.pack([infoset_of(r) for r in recs], [names[1 - r["player"]] if r["hidden"] else None for r in recs])
want_obs = np.asarray([r["obs"] for r in recs], dtype=np.int32)
rep = (B + len(recs) - 1) // len(recs)
dbatch = {k: torch.from_numpy(np.tile(v, (rep,) + (1,) * (v.ndim - 1))[:B]).to(dev) for k, v in base.items()}

games = [g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "settle_run.json"))) if g["P"] == 2]
G = len(games)
boards = phe.ids_to_masks(np.array([g["board"] for g in games]))
holes = phe.ids_to_masks(np.array([g["holes"] for g in games]))
act = np.zeros((G, 4, 2), np.int8)
val = np.zeros((G, 4, 2), np.int32)
for i, g in enumerate(games):
    act[i, : len(g["actions"])] = g["actions"]
    val[i, : len(g["actions"])] = g["values"]
nr = np.array([len(g["actions"]) for g in games], np.int32)
bt = np.array([g["button"] for g in games], np.int32)
ob = np.array([sum(b << p for p, b in enumerate(g["onboard"])) for g in games], np.int32)
rg = (B + G - 1) // G
tile = lambda a: torch.from_numpy(np.tile(a, (rg,) + (1,) * (a.ndim - 1))[:B]).to(dev)  # noqa: E731
sargs = [tile(a) for a in (boards, holes, act, val, nr, bt, ob)]
want_won = np.tile(np.array([g["won"] for g in games], np.int32), (rg, 1))[:B]

hero_board = [np.array(r["hand"] + (r["flop"] or []) + (r["turn"] or []) + (r["river"] or []), dtype=np.int64) for r in recs]
# re-deals: heads-up pre-flop roots (need 2 + 5 = 7 cards), the most expensive determinisation
pre = [hb for hb in hero_board if len(hb) == 2]
kmask = np.array([sum(1 << (16 * (int(c) & 3) + (int(c) >> 2)) for c in hb) for hb in pre], dtype=np.int64)
dknown = torch.from_numpy(np.tile(kmask, (B + len(kmask) - 1) // len(kmask))[:B]).to(dev)

params = pv.DEBUG_MODEL_PARAMS if which == "debug" else pv.DEFAULT_MODEL_PARAMS
pred = pv.BatchedPredictor(pv.fill_parameters(pv.PolicyValueNet(**params), 1), dtype=torch.bfloat16, max_batch=min(B, 4096), buckets=[min(B, 4096)])
pred.warmup()


def wave(seed):
    ids = phe.deal(dknown, 7, seed)
    obs = enc.encode(dbatch)
    probs, value = pred.predict_batch(obs)
    won, res, card, paid = phe.settle(*sargs, 25)
    return ids, obs, probs, value, won


ids, obs, probs, value, won = wave(1)
torch.cuda.synchronize()
assert np.array_equal(obs.cpu().numpy(), np.tile(want_obs, (rep, 1))[:B]), "observations differ from the reference's"
assert np.array_equal(won.cpu().numpy(), want_won), "settlement differs from the reference's"
idn = ids.cpu().numpy()
assert all(len(set(row.tolist())) == 7 for row in idn[:256]) and float(probs.sum(1).sub(1).abs().max()) < 1e-3


def timed(fn, k=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


t_deal = timed(lambda: phe.deal(dknown, 7, 3))
t_enc = timed(lambda: enc.encode(dbatch))
t_pred = timed(lambda: pred.predict_batch(obs))
t_set = timed(lambda: phe.settle(*sargs, 25))
t_wave = timed(lambda: wave(5))
out = {"wave": B, "model": which, "ms": {"redeal": t_deal, "encode_obs": t_enc, "predict": t_pred, "settle": t_set, "wave": t_wave},
       "per_s": {"redeals": B / t_deal * 1e3, "observations_encoded": B / t_enc * 1e3, "nn_leaf_evaluations": B / t_pred * 1e3,
                 "terminal_settlements": B / t_set * 1e3, "waves_of_all_four": B / t_wave * 1e3},
       "checked": "observations and chips equal the reference's goldens; probabilities sum to 1; re-deals are 7 distinct cards"}
print(json.dumps(out))
