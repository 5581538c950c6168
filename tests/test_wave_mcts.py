"""Tree logic of the lock-step MCTS driver (mcts_wave.WaveMCTS) against the reference's own SingleThreadMCTS
(rl/MCTS.py:85-150, 343-365, 410-473), on the CPU: the four batched stages are substituted by reference-backed ones, so
with wave=1 the search must reproduce the reference's visit counts and action probabilities draw for draw."""
import copy
import random

import numpy as np
import pytest


def _params(reference_tree):
    import os
    import yaml
    from tools.param_parser import update_concurrent
    cfg = os.path.join(reference_tree, "config")
    default = yaml.safe_load(open(os.path.join(cfg, "default.yml"), encoding="utf-8"))
    debug = yaml.safe_load(open(os.path.join(cfg, "debug.yml"), encoding="utf-8"))
    return update_concurrent(default.copy(), debug.copy())


class RefBackends:
    """new_random / Env.get_obs / the reference model / finish_game, one row at a time, in the order the reference calls them."""

    def __init__(self, model, env):
        self.model, self.env = model, env

    def deal(self, game_env, n, first_simulation):
        out = []
        for _ in range(n):
            g = game_env.new_random()
            hands = [info.player_hand_cards for name, info in g.info_sets.items() if name != game_env.acting_player_name]
            out.append((hands, g.flop_cards, g.turn_cards, g.river_cards))
        return out

    def observe(self, infosets, hidden_names):
        return np.stack([self.env.get_obs(i, mcts_acting_player_name=h) for i, h in zip(infosets, hidden_names)])

    def predict_batch(self, observations):
        import torch
        probs, values = [], []
        with torch.no_grad():
            for row in np.asarray(observations):
                x = torch.tensor(np.array([row]), dtype=torch.int32)
                logits, value, _, _, _ = self.model(x)
                probs.append(torch.softmax(logits, dim=1).numpy()[0])
                values.append(value.numpy()[0])
        return np.stack(probs), np.array(values, dtype=np.float32)

    def settle(self, game_envs):
        from texasholdempocker_rl_b200.winner import GamePlayerResult
        out = []
        for g in game_envs:
            h = copy.deepcopy(g)
            won = h.finish_game()
            out.append((won, {n: GamePlayerResult(p.game_result.value) for n, p in h.players.items()},
                        {n: GamePlayerResult(p.card_result.value) for n, p in h.players.items()}))
        return out


def _setup(reference_tree, seed):
    import torch
    from env.env import Env
    from rl.AlphaGoZero import AlphaGoZero
    params = _params(reference_tree)
    mp = dict(params["model_param_dict"], save_train_data=False)
    torch.manual_seed(seed)
    model = AlphaGoZero(**mp).model.eval()
    env = Env(None, mp["num_output_class"], mp["historical_action_per_round"], small_blind=25, big_blind=50, ignore_all_async_tasks=True)
    return params, mp, model, env


@pytest.mark.parametrize("seed,steps_before", [(3, 0), (11, 2), (29, 4)])
def test_wave1_equals_reference_search(reference_tree, phe, seed, steps_before):
    torch = pytest.importorskip("torch")
    torch.set_num_threads(1)
    from env.constants import ChoiceMethod
    from rl.MCTS import SingleThreadMCTS
    from utils.workflow_utils import map_action_bin_to_actual_action_and_value_v2
    from texasholdempocker_rl_b200.mcts_wave import WaveMCTS
    params, mp, model, env = _setup(reference_tree, seed)
    random.seed(seed)
    np.random.seed(seed)
    obs = env.reset(0, seed=seed)
    for _ in range(steps_before):                  # walk into the hand with passive legal actions (check / call)
        mask, ranges, left, hist, dmin = env.get_valid_action_info_v2()
        action = map_action_bin_to_actual_action_and_value_v2(1, ranges, left, hist, dmin, 25)
        obs, _, done, _ = env.step(action, 1)
        assert not done
    kw = dict(n_simulation=40, c_puct=params["mcts_c_puct"], tau=params["mcts_tau"], model_Q_epsilon=params["mcts_model_Q_epsilon"],
              choice_method=ChoiceMethod(params["mcts_choice_method"]))
    random.seed(seed + 1)
    np.random.seed(seed + 1)
    ref = SingleThreadMCTS(mp["num_output_class"], is_root=True, apply_dirichlet_noice=False, small_blind=25, model=model, log_to_file=False, pid=0, **kw)
    want = ref.simulate(obs, env)
    random.seed(seed + 1)
    np.random.seed(seed + 1)
    terminals = []
    wave = WaveMCTS(mp["num_output_class"], 25, RefBackends(model, env), wave=1, pid=0, on_terminal=terminals.append, **kw)
    got = wave.simulate(obs, env)
    assert np.array_equal(wave.root.children_n_array, ref.children_n_array)
    np.testing.assert_allclose(wave.root.children_w_array, ref.children_w_array, rtol=0, atol=1e-12)
    assert np.array_equal(got, want)
    assert len(terminals) == 40 and wave.stats["terminals"] == 40
    for g in terminals:                            # chips are conserved on every simulated terminal
        assert sum(p.value_win for p in g.players.values()) == sum(p.value_bet for p in g.players.values())


def test_wave_batches_and_conserves(reference_tree, phe):
    """wave > 1: every simulation is backed up exactly once, visit counts add up, virtual visits are all resolved."""
    torch = pytest.importorskip("torch")
    torch.set_num_threads(1)
    from env.constants import ChoiceMethod
    from texasholdempocker_rl_b200.mcts_wave import WaveMCTS
    params, mp, model, env = _setup(reference_tree, 5)
    random.seed(5)
    np.random.seed(5)
    obs = env.reset(0, seed=5)
    terminals = []
    wave = WaveMCTS(mp["num_output_class"], 25, RefBackends(model, env), n_simulation=70, c_puct=params["mcts_c_puct"], tau=1, model_Q_epsilon=0.2,
                    choice_method=ChoiceMethod.ARGMAX, wave=16, on_terminal=terminals.append)
    probs = wave.simulate(obs, env)
    assert abs(probs.sum() - 1) < 1e-9 and wave.root.children_n_array.sum() == 70 and len(terminals) == 70
    assert wave.stats["waves"] == 5 and wave.stats["nn_calls"] < wave.stats["nn_rows"]
    mask = env.get_valid_action_info_v2()[0]
    assert all(n == 0 for n, m in zip(wave.root.children_n_array, mask) if m)       # illegal actions are never visited

    def check(node):
        n = node.children_n_array
        q = np.divide(node.children_w_array, n, out=np.zeros_like(n), where=n > 0)
        np.testing.assert_allclose(node.children_q_array, q, atol=1e-12)
        for a, child in enumerate(node.children):
            if child is not None and child.children_n_array is not None:
                assert child.children_n_array.sum() <= n[a]          # terminal-at-child simulations do not descend
                check(child)
    check(wave.root)
