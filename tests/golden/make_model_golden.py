"""Generates tests/golden/model_run.json: outputs of the REFERENCE'S OWN TransformerAlphaGoZeroModel
(/root/reference/rl/models.py:402-481, unmodified, CPU fp32, eval mode) on synthetic observations.

  python tests/golden/make_model_golden.py       (build container only)

Weights are not stored: both the reference model here and PolicyValueNet in the tests are filled by
policy_value.fill_parameters(seed) (one seeded CPU generator over the sorted state_dict keys, which are identical).
Also records SingleThreadMCTS.predict's post-processing (softmax of the action logits, rl/MCTS.py:464-473)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

from rl.models import TransformerAlphaGoZeroModel  # noqa: E402
import texasholdempocker_rl_b200 as phe  # noqa: E402
pv = phe.policy_value

torch.manual_seed(0)
out = {"cases": []}
for name, params, n_obs, seed in (("debug", pv.DEBUG_MODEL_PARAMS, 48, 11), ("default", pv.DEFAULT_MODEL_PARAMS, 6, 12)):
    ref = TransformerAlphaGoZeroModel(**{k: params[k] for k in params}).eval()
    mine = pv.PolicyValueNet(**params)
    assert sorted(ref.state_dict()) == sorted(mine.state_dict()), "state_dict keys differ"
    for k, v in ref.state_dict().items():
        assert tuple(v.shape) == tuple(mine.state_dict()[k].shape), k
        if not v.dtype.is_floating_point:
            assert torch.equal(v.to(torch.int64), mine.state_dict()[k].to(torch.int64)), k
    pv.fill_parameters(ref, seed)
    obs = pv.random_observations(n_obs, seed=seed, num_output_class=params["num_output_class"], num_bins=params["num_bins"],
                                 historical_action_per_round=params["historical_action_per_round"])
    with torch.no_grad():
        o = ref(torch.from_numpy(obs))
        probs = torch.softmax(o[0], dim=1)
    out["cases"].append({"name": name, "params": params, "seed": seed, "obs": obs.tolist(),
                         "outputs": [t.double().tolist() for t in o], "action_probs": probs.double().tolist()})
    print(name, [tuple(t.shape) for t in o])
path = os.path.join(ROOT, "tests", "golden", "model_run.json")
with open(path, "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes")
