"""B200-native hand evaluation and equity for the hot path of weipeilun/texasholdempocker-rl.

Python surface (mirrors the reference where one exists):
  card_to_evaluator, Card, CardFigure, CardDecor, DECK          utils/pokerhandevaluator_utils.py, env/cards.py
  evaluate_cards                                                phevaluator.evaluate_cards (env/game.py:675)
  GameEnv.get_winner_set_v2 / generate_game_result              env/game.py:663-681, 551-558
  simulate_processes / solve_tasks / winning_probability        env/workflow.py:102-270
  settle / settle_env (finish_game, _share_value)               env/game.py:410-441, 560-632, 683-734
  ObsEncoder.get_obs / encode                                   env/env.py:383-408, 416-581
  PolicyValueNet, BatchedPredictor.predict / predict_batch       rl/models.py:402-481, rl/MCTS.py:464-473
  predict_batch_process                                         env/workflow.py:274-593 (queue protocol of the prediction server)
  dropin.install / uninstall                                    rebinding of the reference's own modules
  WaveMCTS, GpuBackends                                         rl/MCTS.py:85-150, 343-365 run W simulations per wave
  eval7, enumerate_c52_7, equity_mc, equity_plan, showdown, deal   batched entry points (include/phe_b200.h)
"""
from . import _native
from .cardset import DECK, Card, CardDecor, CardFigure, card_id, card_to_evaluator, cards_to_mask, ids_to_masks, mask_to_ids
from .equity import deal, determinise, equity_mc, equity_plan, pack_states, plan_num_deals, random_states, showdown, win_probability
from .evaluator import enumerate_c52_7, eval7, evaluate_cards
from . import observation, policy_value, records
from .observation import ObsEncoder
from .policy_value import BatchedPredictor, PolicyValueNet
from .predict_server import predict_batch_process
from . import dropin, mcts_wave
from .mcts_wave import GpuBackends, WaveMCTS
from .service import simulate_processes, solve_tasks, winning_probability
from .settlement import settle, settle_env, settle_envs
from .sharded import FusedEnum7, shard_range, sharded_enum7, sharded_equity_mc
from .winner import GameEnv, GamePlayerResult, winner_sets

__all__ = [
    "DECK", "Card", "CardDecor", "CardFigure", "card_id", "card_to_evaluator", "cards_to_mask", "ids_to_masks", "mask_to_ids",
    "deal", "determinise", "equity_mc", "equity_plan", "pack_states", "plan_num_deals", "random_states", "showdown", "win_probability",
    "enumerate_c52_7", "eval7", "evaluate_cards", "simulate_processes", "solve_tasks", "winning_probability",
    "settle", "settle_env", "settle_envs", "FusedEnum7", "shard_range", "sharded_enum7", "sharded_equity_mc", "GameEnv", "GamePlayerResult", "winner_sets",
    "BatchedPredictor", "PolicyValueNet", "ObsEncoder", "predict_batch_process", "dropin", "mcts_wave", "WaveMCTS", "GpuBackends",
]
