"""In-process installation of the drop-ins into the reference's own modules (INTEGRATION.md sections A and C).

The reference is a script tree without a plugin mechanism, so "drop-in" means rebinding the names its hot path
resolves at call time.  ``install()`` applies, to the reference modules already importable as ``env.game``,
``env.env``, ``env.workflow`` and ``rl.MCTS`` (an unmodified checkout on ``sys.path``), exactly the edits
INTEGRATION.md describes:

  env/game.py:10        ``evaluate_cards``                     -> ``evaluator.evaluate_cards``          (phe_eval7_host)
  env/game.py:663-681   ``GameEnv.get_winner_set_v2``          -> ``winner.GameEnv.get_winner_set_v2``   (phe_showdown_host)
  env/game.py:410-441   ``GameEnv.finish_game``  (settle=True) -> ``settlement.settle_env``              (phe_settle_host)
  env/workflow.py:102   ``simulate_processes``                 -> ``service.simulate_processes``         (phe_equity_plan)
  env/workflow.py:274   ``predict_batch_process``              -> ``predict_server.predict_batch_process`` (BatchedPredictor)
  env/env.py:416        ``Env.get_obs``          (encoder=...) -> ``ObsEncoder.get_obs``                 (phe_encode_obs)
  rl/MCTS.py:464-473    ``SingleThreadMCTS.predict`` (predictor=...) -> ``BatchedPredictor.predict``     (CUDA graph replay)

``rl/MCTS.py``, ``rl/AlphaGoZero.py`` and the betting engine run unchanged on top.  ``uninstall()`` restores every
name.  tests/test_insitu.py plays the reference's own SingleThreadMCTS self-play through this and compares every
logged terminal with the same run on the stock (shim) evaluator.
"""
import importlib

from . import evaluator, predict_server, service, settlement, winner

_saved = []


def _swap(owner, name, value):
    _saved.append((owner, name, owner.__dict__[name] if name in owner.__dict__ else getattr(owner, name)))
    setattr(owner, name, value)


def _finish_game(self):
    """GameEnv.finish_game (env/game.py:410-441) with the three showdown passes done by one phe_settle call: chips from
    ``_share_value``, the WIN / EVEN / LOSE result and the action-free card result; the bookkeeping around them is the
    reference's."""
    self.game_over = True
    const = importlib.import_module("env.constants")        # the reference's own enums (env/constants.py:38-68)
    won, results, card_results = settlement.settle_env(self)
    for name, value in won.items():
        player = self.players[name]
        player.value_left += value
        player.value_win += value
    for name, player in self.players.items():
        player.game_result = const.GamePlayerResult(results[name].value)
        player.card_result = const.GamePlayerResult(card_results[name].value)
    for player in self.players.values():
        player.action = None
        if player.value_left <= 0:
            player.status = const.PlayerStatus.BUSTED
    return won


def install(predictor=None, encoder=None, settle=False, equity_service=True):
    """Rebinds the reference's hot-path names to this library.  ``predictor``: a ``BatchedPredictor`` for
    ``SingleThreadMCTS.predict``; ``encoder``: an ``ObsEncoder`` for ``Env.get_obs``; ``settle``: also replace
    ``GameEnv.finish_game`` by the batched settlement kernel.  Returns the list of patched names."""
    if _saved:
        raise RuntimeError("drop-ins already installed; call uninstall() first")
    game = importlib.import_module("env.game")
    _swap(game, "evaluate_cards", evaluator.evaluate_cards)
    _swap(game.GameEnv, "get_winner_set_v2", staticmethod(winner.GameEnv.get_winner_set_v2))
    if settle:
        _swap(game.GameEnv, "finish_game", _finish_game)
    if equity_service:
        workflow = importlib.import_module("env.workflow")
        _swap(workflow, "simulate_processes", service.simulate_processes)
        _swap(workflow, "predict_batch_process", predict_server.predict_batch_process)
    if encoder is not None:
        env_mod = importlib.import_module("env.env")
        _swap(env_mod.Env, "get_obs", lambda self, infoset, mcts_acting_player_name=None: encoder.get_obs(infoset, mcts_acting_player_name))
    if predictor is not None:
        mcts = importlib.import_module("rl.MCTS")

        def predict(self, observation):
            return predictor.predict(observation)        # (float32[A], float32 scalar), as rl/MCTS.py:472-473 returns them

        _swap(mcts.SingleThreadMCTS, "predict", predict)
    return [f"{getattr(o, '__name__', o)}.{n}" for o, n, _ in _saved]


def uninstall():
    while _saved:
        owner, name, old = _saved.pop()
        setattr(owner, name, old)


def installed():
    return bool(_saved)
