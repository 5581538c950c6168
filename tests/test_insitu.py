"""In-situ drop-in proof (BASELINE config 4): the reference's OWN code -- GameEnv / Env /
SingleThreadMCTS from the unmodified checkout in baseline/_ref -- runs with this library rebound into it by
``dropin.install``, in one process, and is compared with the same run on the stock path
(restated phevaluator, reference model on the CPU).

Integer stages (evaluate_cards, get_winner_set_v2, finish_game settlement, get_obs) are bit-exact, so with the same
seeds whole self-play trajectories must be identical.  With the fp32 device predictor in place as well, every logged
terminal is re-settled by the stock code and must agree; actions must be legal and chips conserved."""
import copy
import random

import numpy as np
import pytest

from test_wave_mcts import _setup

N_SIM = 16
HANDS = 6


def _logging_finish(log, inner):
    def finish_game(self):
        before = copy.deepcopy(self)
        out = inner(self)
        log.append((before, dict(out), {n: (p.value_win, p.value_left, p.game_result.value, p.card_result.value, p.status.name)
                                        for n, p in self.players.items()}))
        return out
    return finish_game


def _self_play(reference_tree, seed, hands, log, predictor_model):
    """HANDS hands of heads-up self-play driven by the reference's SingleThreadMCTS, as alpha_go_zero_player.py:56-72 drives it."""
    import env.game as game_mod
    from env.constants import ChoiceMethod
    from rl.MCTS import SingleThreadMCTS
    from env.env import Env
    params, mp, model, _ = predictor_model
    env = Env(None, mp["num_output_class"], mp["historical_action_per_round"], small_blind=25, big_blind=50, ignore_all_async_tasks=True)   # fresh stacks per run
    inner = game_mod.GameEnv.finish_game               # the stock method, or the drop-in when installed with settle=True
    game_mod.GameEnv.finish_game = _logging_finish(log, inner)
    trace = []
    try:
        random.seed(seed)
        np.random.seed(seed)
        for h in range(hands):
            obs = env.reset(h, seed=seed * 100 + h)
            for _ in range(40):
                mask = env.get_valid_action_info_v2()[0]
                mcts = SingleThreadMCTS(mp["num_output_class"], is_root=True, apply_dirichlet_noice=False, small_blind=25, model=model,
                                        n_simulation=N_SIM, c_puct=params["mcts_c_puct"], tau=params["mcts_tau"],
                                        model_Q_epsilon=params["mcts_model_Q_epsilon"], choice_method=ChoiceMethod(params["mcts_choice_method"]),
                                        log_to_file=False, pid=0)
                probs = mcts.simulate(obs, env)
                action_bin, action, _ = mcts.get_action(probs, env=env, choice_method=ChoiceMethod.PROBABILITY)
                assert not mask[action_bin], "MCTS chose a masked action"
                trace.append((h, int(action_bin), action[0].name, int(action[1]), [float(p) for p in probs], [int(v) for v in obs]))
                obs, reward, done, _ = env.step(action, action_bin)
                if done:
                    trace.append((h, "reward", {n: tuple(float(x) for x in r) for n, r in reward.items()}))
                    break
    finally:
        game_mod.GameEnv.finish_game = inner
    return trace


def _play(reference_tree, seed, setup, **install_kw):
    """One logged self-play run; with install_kw the drop-ins are installed around it."""
    from texasholdempocker_rl_b200 import dropin
    log = []
    if install_kw.pop("dropins", False):
        names = dropin.install(**install_kw)
        assert "env.game.evaluate_cards" in names
    try:
        return _self_play(reference_tree, seed, HANDS, log, setup), log
    finally:
        dropin.uninstall()


def _terminal_key(entry):
    before, won, players = entry
    return (sorted(won.items()), sorted(players.items()))


def test_install_rebinds_and_restores(reference_tree, phe):
    import env.game as game_mod
    import env.workflow as workflow_mod
    from texasholdempocker_rl_b200 import dropin
    stock = (game_mod.evaluate_cards, game_mod.GameEnv.__dict__["get_winner_set_v2"], game_mod.GameEnv.finish_game, workflow_mod.simulate_processes)
    names = dropin.install(settle=True)
    try:
        assert game_mod.evaluate_cards is phe.evaluate_cards and workflow_mod.simulate_processes is phe.simulate_processes
        assert game_mod.GameEnv.finish_game is not stock[2] and len(names) == 5 and workflow_mod.predict_batch_process is phe.predict_batch_process
        with pytest.raises(RuntimeError):
            dropin.install()
    finally:
        dropin.uninstall()
    assert (game_mod.evaluate_cards, game_mod.GameEnv.__dict__["get_winner_set_v2"], game_mod.GameEnv.finish_game,
            workflow_mod.simulate_processes) == stock and not dropin.installed()


@pytest.mark.gpu
def test_reference_selfplay_identical_on_integer_dropins(reference_tree, phe):
    """evaluate_cards + get_winner_set_v2 (+ finish_game -> phe_settle) (+ Env.get_obs -> phe_encode_obs) rebound inside the
    reference: same seeds => the same actions, MCTS probabilities, observations, rewards and terminal settlements."""
    import torch
    torch.set_num_threads(1)
    setup = _setup(reference_tree, 17)
    want, want_log = _play(reference_tree, 17, setup)
    assert len(want_log) >= HANDS * N_SIM            # every simulation and every hand ends in a logged finish_game
    for kw in (dict(), dict(settle=True), dict(settle=True, encoder=phe.ObsEncoder(setup[1]["num_output_class"], setup[1]["historical_action_per_round"]))):
        got, got_log = _play(reference_tree, 17, setup, dropins=True, **kw)
        assert got == want, f"trajectory differs with {sorted(kw)}"
        assert [_terminal_key(e) for e in got_log] == [_terminal_key(e) for e in want_log], f"terminals differ with {sorted(kw)}"


@pytest.mark.gpu
def test_reference_selfplay_on_all_dropins(reference_tree, phe):
    """All of INTEGRATION.md at once (device predictor included): legal actions, chip conservation, and every logged
    terminal re-settled by the stock reference code (shim evaluator) gives the same chips / results."""
    import torch
    torch.set_num_threads(1)
    setup = _setup(reference_tree, 23)
    params, mp, model, env = setup
    net = phe.PolicyValueNet(**mp)
    net.load_reference_state_dict(model.state_dict())
    predictor = phe.BatchedPredictor(net, dtype=torch.float32, max_batch=64)
    trace, log = _play(reference_tree, 23, setup, dropins=True, settle=True, predictor=predictor,
                       encoder=phe.ObsEncoder(mp["num_output_class"], mp["historical_action_per_round"]))
    assert predictor.replays >= HANDS * N_SIM and len(log) >= HANDS * N_SIM
    for before, won, players in log:
        stock = copy.deepcopy(before)
        stock_won = stock.finish_game()                      # unpatched again: the stock path on the restated phevaluator
        assert dict(stock_won) == won
        assert {n: (p.value_win, p.value_left, p.game_result.value, p.card_result.value, p.status.name) for n, p in stock.players.items()} == players
        assert sum(v[0] for v in players.values()) == sum(p.value_bet for p in before.players.values())      # chips conserved
    # the device predictor agrees with the reference's own model on the observations this run produced
    obs = np.array([t[5] for t in trace if t[1] != "reward"], dtype=np.int32)
    p_dev, v_dev = predictor.predict_batch(obs)
    with torch.no_grad():
        logits, value, _, _, _ = model(torch.from_numpy(obs))
    assert np.abs(p_dev - torch.softmax(logits, 1).numpy()).max() < 1e-3 and np.abs(v_dev - value.numpy()).max() < 2e-3


class _SnapshotBackends:
    def __init__(self, inner):
        self.inner, self.snaps, self.rows = inner, [], []

    def __getattr__(self, name):
        return getattr(self.inner, name)

    def settle(self, game_envs):
        self.snaps += [copy.deepcopy(g) for g in game_envs]
        rows = self.inner.settle(game_envs)
        self.rows += rows
        return rows


@pytest.mark.gpu
@pytest.mark.parametrize("steps_before", [0, 2, 4])
def test_wave_mcts_on_device_stages(reference_tree, phe, steps_before):
    """Lock-step search (wave = 32) with determinise / ObsEncoder / BatchedPredictor / settle_envs as the stages: visits land on
    legal actions only, each terminal's batched settlement equals the stock finish_game of the same position, re-deals respect
    the cards the acting player sees, and the search concentrates like the sequential reference search on the same position."""
    import torch
    from env.constants import ChoiceMethod
    from rl.MCTS import SingleThreadMCTS
    from utils.workflow_utils import map_action_bin_to_actual_action_and_value_v2
    from texasholdempocker_rl_b200.mcts_wave import GpuBackends, WaveMCTS
    torch.set_num_threads(1)
    params, mp, model, env = _setup(reference_tree, 31)
    random.seed(31)
    np.random.seed(31)
    obs = env.reset(0, seed=31)
    for _ in range(steps_before):
        mask, ranges, left, hist, dmin = env.get_valid_action_info_v2()
        obs, _, done, _ = env.step(map_action_bin_to_actual_action_and_value_v2(1, ranges, left, hist, dmin, 25), 1)
        assert not done
    net = phe.PolicyValueNet(**mp)
    net.load_reference_state_dict(model.state_dict())
    backends = _SnapshotBackends(GpuBackends(phe.BatchedPredictor(net, dtype=torch.float32, max_batch=64),
                                             phe.ObsEncoder(mp["num_output_class"], mp["historical_action_per_round"]), seed=99))
    kw = dict(n_simulation=160, c_puct=params["mcts_c_puct"], tau=1, model_Q_epsilon=params["mcts_model_Q_epsilon"], choice_method=ChoiceMethod.ARGMAX)
    wave = WaveMCTS(mp["num_output_class"], 25, backends, wave=32, **kw)
    probs = wave.simulate(obs, env)
    mask = env.get_valid_action_info_v2()[0]
    assert abs(probs.sum() - 1) < 1e-9 and wave.root.children_n_array.sum() == 160
    assert all(p == 0 for p, m in zip(probs, mask) if m)
    assert wave.stats["waves"] == 5 and wave.stats["nn_calls"] <= 1 + 5 * 12
    g0 = env._env
    hero = g0.info_sets[g0.acting_player_name].player_hand_cards
    assert len(backends.snaps) == 160
    for snap, (won, results, cards) in zip(backends.snaps, backends.rows):
        assert snap.info_sets[g0.acting_player_name].player_hand_cards == hero
        if g0.current_round >= 1:
            assert snap.flop_cards == g0.flop_cards
        seen = [c for i in snap.info_sets.values() for c in i.player_hand_cards] + snap.flop_cards + snap.turn_cards + snap.river_cards
        assert len(set(seen)) == 9
        stock = copy.deepcopy(snap)
        stock_won = stock.finish_game()
        assert dict(stock_won) == won
        assert {n: p.game_result.value for n, p in stock.players.items()} == {n: r.value for n, r in results.items()}
        assert {n: p.card_result.value for n, p in stock.players.items()} == {n: r.value for n, r in cards.items()}
    # same position through the reference's sequential search: the two visit distributions are close (both are noisy)
    ref = SingleThreadMCTS(mp["num_output_class"], is_root=True, apply_dirichlet_noice=False, small_blind=25, model=model, log_to_file=False, pid=0, **kw)
    want = ref.simulate(obs, env)
    assert np.abs(np.asarray(want) - probs).sum() < 0.5
