"""Generates tests/golden/records_run.json: the text the REFERENCE'S OWN AlphaGoZero.store_transition
(/root/reference/rl/AlphaGoZero.py:90-103) writes for seeded random transitions, and what the reference's own reader
(tools/data_loader.py:train_data_generator) parses back from it.   python tests/golden/make_records_golden.py"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
import numpy as np  # noqa: E402

from rl.AlphaGoZero import AlphaGoZero  # noqa: E402
from tools.data_loader import train_data_generator  # noqa: E402

path = os.path.join(tempfile.mkdtemp(), "train.txt")
m = AlphaGoZero(n_observation=93, n_actions=12, num_output_class=12, embedding_dim=64, positional_embedding_dim=32, num_layers=1,
                transformer_head_dim=16, historical_action_per_round=3, num_acting_player_fields=50, save_train_data=True,
                default_train_file_path=path)
rng = np.random.default_rng(20261020)
inputs = []
for i in range(40):
    obs = rng.integers(0, 14, 93).astype(np.int32)
    probs = rng.dirichlet(np.ones(12) * 0.3)
    masks = (rng.random(12) < 0.3).astype(int).tolist()
    reward = float(rng.normal() * (10.0 ** rng.integers(-6, 2)))
    card = float(rng.choice([-1.0, 0.0, 1.0, 0.5, 0.123456789]))
    pb, ob = int(rng.integers(0, 10)), int(rng.integers(0, 10))
    m.store_transition(obs, probs, masks, reward, card, pb, ob)
    inputs.append({"obs": obs.tolist(), "probs": probs.tolist(), "masks": masks, "reward": reward, "card": card, "pbin": pb, "obin": ob})
m.train_file.flush()
text = open(path).read()
parsed = [{"obs": f.tolist(), "probs": p, "masks": k, "reward": r, "card": c, "pbin": a, "obin": b} for f, p, k, r, c, a, b in train_data_generator(path, epoch=1)]
json.dump({"inputs": inputs, "text": text, "parsed": parsed}, open(os.path.join(ROOT, "tests", "golden", "records_run.json"), "w"))
print(len(inputs), "records,", len(text), "bytes of text,", len(parsed), "parsed back")
