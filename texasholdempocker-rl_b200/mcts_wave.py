"""Lock-step MCTS leaf evaluation: the reference's search (rl/MCTS.py:85-150, 343-365) run W simulations at a time.

The reference walks one determinised simulation at a time: ``env.new_random()`` (about twenty ``deepcopy`` calls and a
``random.sample``), then per tree level one ``Env.get_obs`` and one single-row network call, then ``finish_game`` with
up to three Python showdowns.  ``WaveMCTS.simulate`` keeps the reference's tree, selection rule and betting engine --
the nodes ARE ``rl.MCTS.MCTS`` objects and every action is chosen by their own ``_choose_action`` and applied by the
reference's ``GameEnv.step`` -- and batches everything that is data parallel across the W simulations of a wave:

  re-deal        W x ``GameEnv.new_random``          -> one ``determinise`` call            (phe_deal, k_deal)
  observation    per level, all live simulations     -> one ``ObsEncoder.encode``           (phe_encode_obs)
  policy/value   per level, all live simulations     -> one ``BatchedPredictor.predict_batch`` (CUDA-graph replay)
  settlement     all W terminals                     -> one ``settle_envs``                 (phe_settle, k_settle)

The root position is serialised once per search and each simulation is one ``pickle.loads`` of it instead of the
per-field deep copies.  Simulations of one wave select with the statistics of the previous waves plus a virtual visit
(N += 1 at selection time) for every simulation already in flight, so with ``wave=1`` the search is the reference's
sequential algorithm, draw for draw (tests/test_wave_mcts.py checks that against ``SingleThreadMCTS`` itself).

The reference tree (``env.*``, ``rl.*``, ``utils.*``) must be importable: the driver works on the reference's own
``Env`` / ``GameEnv`` objects and does not restate the betting rules (out of scope).
"""
import importlib
import pickle
import time

import numpy as np

from .equity import determinise
from .settlement import settle_envs


class _SimEnv:
    """What ``MCTS._choose_action`` and ``Env._get_final_reward`` read of an ``Env`` (env/env.py:370-381, 413-414)."""
    __slots__ = ("_env", "path")

    def __init__(self, game_env):
        self._env = game_env
        self.path = []

    @property
    def acting_player_name(self):
        return self._env.acting_player_name

    def get_valid_action_info_v2(self):
        return self._env.get_valid_action_info_v2()


class GpuBackends:
    """The four batched stages on the device.  Tests substitute reference-backed stages to check the tree logic alone."""

    def __init__(self, predictor, encoder, seed=0):
        self.predictor, self.encoder, self.seed = predictor, encoder, seed
        self.stream = 0                                              # Philox stream of the current search: one per simulate() call
        self._deck = importlib.import_module("env.game").deck        # deck[i] is the card with id i (env/game.py:14-18)

    def new_search(self):
        """Called by WaveMCTS at the start of every search: two searches never replay the same re-deals, even when the
        acting player sees the same cards (two decisions on one street)."""
        self.stream = (self.stream + 1) & 0xFFFFFFFF

    def deal(self, game_env, n, first_simulation):
        """n re-deals of what the acting player cannot see -> list of (hands of the other seats in ``info_sets`` order,
        flop, turn, river) as reference ``Card`` lists, sorted the way ``new_random`` sorts them (env/game.py:141-163)."""
        acting = game_env.acting_player_name
        shown = []
        if game_env.current_round >= 1:
            shown += list(game_env.flop_cards)
        if game_env.current_round >= 2:
            shown += list(game_env.turn_cards)
        if game_env.current_round >= 3:
            shown += list(game_env.river_cards)
        opp, board = determinise(game_env.info_sets[acting].player_hand_cards, shown, game_env.num_players, n, self.seed,
                                 sample_offset=first_simulation, stream=self.stream)
        d = self._deck
        out = []
        for i in range(n):
            hands = [[d[a], d[b]] for a, b in opp[i].tolist()]
            b = board[i].tolist()
            out.append((hands, [d[b[0]], d[b[1]], d[b[2]]], [d[b[3]]], [d[b[4]]]))
        return out

    def observe(self, infosets, hidden_names):
        return self.encoder.encode(self.encoder.pack(infosets, hidden_names))

    def predict_batch(self, observations):
        return self.predictor.predict_batch(observations)

    def settle(self, game_envs):
        return settle_envs(game_envs)


def apply_settlement(game_env, settled):
    """The bookkeeping of ``finish_game`` (env/game.py:415-439) around one ``settle_envs`` row."""
    const = importlib.import_module("env.constants")
    won, results, card_results = settled
    for name, value in won.items():
        player = game_env.players[name]
        player.value_left += value
        player.value_win += value
    for name, player in game_env.players.items():
        player.game_result = const.GamePlayerResult(results[name].value)
        player.card_result = const.GamePlayerResult(card_results[name].value)
        player.action = None
        if player.value_left <= 0:
            player.status = const.PlayerStatus.BUSTED


class WaveMCTS:
    """``simulate(observation, env)`` / ``get_action(...)`` with the reference's signatures (rl/MCTS.py:85, 250)."""

    def __init__(self, n_actions, small_blind, backends, n_simulation=800, c_puct=1., tau=1, model_Q_epsilon=0.5,
                 init_root_n_simulation=4, choice_method=None, wave=64, apply_dirichlet_noice=False, dirichlet_noice_epsilon=0.25,
                 pid=None, on_terminal=None, stats=None):
        self._mcts = importlib.import_module("rl.MCTS")
        self._env_mod = importlib.import_module("env.env")
        self._game_mod = importlib.import_module("env.game")
        const = importlib.import_module("env.constants")
        self.n_actions, self.small_blind, self.backends = n_actions, small_blind, backends
        self.n_simulation, self.c_puct, self.tau = n_simulation, c_puct, tau
        self.model_Q_epsilon, self.init_root_n_simulation = model_Q_epsilon, init_root_n_simulation
        self.choice_method = const.ChoiceMethod(choice_method) if choice_method is not None else const.ChoiceMethod.ARGMAX
        self.wave = max(1, int(wave))
        self.apply_dirichlet_noice, self.dirichlet_noice_epsilon = apply_dirichlet_noice, dirichlet_noice_epsilon
        self.pid = pid
        self.on_terminal = on_terminal            # callback(game_env) on every settled terminal (tests log them)
        self.root = None
        self.stats = stats if stats else {"simulations": 0, "nn_rows": 0, "nn_calls": 0, "terminals": 0, "waves": 0,
                      "t_deal": 0.0, "t_clone": 0.0, "t_tree": 0.0, "t_step": 0.0, "t_observe": 0.0, "t_predict": 0.0, "t_settle": 0.0}

    # -- tree nodes are the reference's own MCTS objects (selection rule: rl/MCTS.py:263-341) --------------------------
    def _node(self, is_root):
        node = self._mcts.MCTS(self.n_actions, is_root, None, None, False, None, None, None, self.small_blind,
                               n_simulation=self.n_simulation, c_puct=self.c_puct, tau=self.tau,
                               dirichlet_noice_epsilon=self.dirichlet_noice_epsilon, model_Q_epsilon=self.model_Q_epsilon,
                               init_root_n_simulation=self.init_root_n_simulation, choice_method=self.choice_method, log_to_file=False,
                               pid=self.pid, thread_name=None)
        node.children = [None] * self.n_actions
        node.children_w_array = np.zeros(self.n_actions)
        node.children_n_array = np.zeros(self.n_actions)
        node.children_q_array = np.zeros(self.n_actions)
        return node

    def _clone(self, blob, deal, acting):
        """One determinised copy of the root position: env/game.py:81-165 with the cards of one batched re-deal."""
        g = object.__new__(self._game_mod.GameEnv)
        g.__dict__.update(pickle.loads(blob))
        g.settle_automatically = False            # terminals are settled together at the end of the wave
        g.ignore_all_async_tasks = True
        hands, flop, turn, river = deal
        k = 0
        for name, info in g.info_sets.items():
            if name != acting:
                info.player_hand_cards = hands[k]
                k += 1
        if g.current_round < 1:
            g.flop_cards = flop
        if g.current_round < 2:
            g.turn_cards = turn
        if g.current_round < 3:
            g.river_cards = river
        return g

    def simulate(self, observation, env):
        T = time.perf_counter
        st = self.stats
        game_env = env._env
        acting = game_env.acting_player_name
        root = self.root = self._node(True)
        if hasattr(self.backends, "new_search"):
            self.backends.new_search()
        probs, values = self.backends.predict_batch(np.asarray(observation, dtype=np.int32).reshape(1, -1))
        st["nn_calls"] += 1
        st["nn_rows"] += 1
        action_prob, root_value = probs[0], values[0]
        if self.apply_dirichlet_noice:              # rl/MCTS.py:103-105
            noise = np.random.dirichlet(action_prob)
            action_prob = (1 - self.dirichlet_noice_epsilon) * action_prob + self.dirichlet_noice_epsilon * noise
        state = dict(vars(game_env))
        state["game_infoset"] = None                # as new_random does (env/game.py:112-114)
        blob = pickle.dumps(state, protocol=pickle.HIGHEST_PROTOCOL)
        eps = self.model_Q_epsilon
        done = 0
        while done < self.n_simulation:
            w = min(self.wave, self.n_simulation - done)
            t0 = T()
            deals = self.backends.deal(game_env, w, done)
            t1 = T()
            st["t_deal"] += t1 - t0
            sims = []
            for j in range(w):
                t0 = T()
                sim = _SimEnv(self._clone(blob, deals[j], acting))
                t1 = T()
                action_bin, action = root._choose_action(action_prob, sim, num_simulation=done + j, do_log=False)
                root.children_n_array[action_bin] += 1            # virtual visit; the value is added at backup
                sim.path.append((root, action_bin, acting, root_value))
                t2 = T()
                sim._env.step(action, action_bin)
                t3 = T()
                st["t_clone"] += t1 - t0
                st["t_tree"] += t2 - t1
                st["t_step"] += t3 - t2
                sims.append(sim)
            live = [s for s in sims if not s._env.game_over]
            while live:
                t0 = T()
                obs = self.backends.observe([s._env.game_infoset for s in live], [acting] * len(live))
                t1 = T()
                probs, values = self.backends.predict_batch(obs)
                t2 = T()
                st["t_observe"] += t1 - t0
                st["t_predict"] += t2 - t1
                st["nn_calls"] += 1
                st["nn_rows"] += len(live)
                for k, sim in enumerate(live):
                    t0 = T()
                    parent, parent_bin = sim.path[-1][0], sim.path[-1][1]
                    node = parent.children[parent_bin]
                    if node is None:
                        node = parent.children[parent_bin] = self._node(False)
                    action_bin, action = node._choose_action(probs[k], sim)
                    node.children_n_array[action_bin] += 1
                    sim.path.append((node, action_bin, sim._env.acting_player_name, values[k]))
                    t1 = T()
                    sim._env.step(action, action_bin)
                    t2 = T()
                    st["t_tree"] += t1 - t0
                    st["t_step"] += t2 - t1
                live = [s for s in live if not s._env.game_over]
            t0 = T()
            settled = self.backends.settle([s._env for s in sims])
            st["t_settle"] += T() - t0
            t0 = T()
            for sim, row in zip(sims, settled):
                apply_settlement(sim._env, row)
                if self.on_terminal is not None:
                    self.on_terminal(sim._env)
                reward = self._env_mod.Env._get_final_reward(sim)                 # env/env.py:199-212
                for node, action_bin, player, value in sim.path:                   # rl/MCTS.py:130-132, 360-362
                    node.children_w_array[action_bin] += eps * value + (1 - eps) * reward[player][1]
                    node.children_q_array[action_bin] = node.children_w_array[action_bin] / node.children_n_array[action_bin]
            st["t_tree"] += T() - t0
            st["terminals"] += w
            st["waves"] += 1
            done += w
        st["simulations"] += self.n_simulation
        return root._cal_action_probs(self.tau)

    def get_action(self, action_probs, env, choice_method):
        return self.root.get_action(action_probs, env, choice_method)
