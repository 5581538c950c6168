"""Showdown winner logic with the reference's Python call surface.

``GameEnv.get_winner_set_v2`` and ``GameEnv.generate_game_result`` keep the signatures and semantics of
env/game.py:663-681 and env/game.py:551-558; the hand ranks come from the B200 showdown kernel.
``winner_sets`` is the batched form used by the equity service and MCTS leaf evaluation.
"""
from enum import Enum

import numpy as np

from .cardset import cards_to_mask
from .equity import showdown


class GamePlayerResult(Enum):
    """env/constants.py:64-68"""
    DEFAULT = None
    LOSE = -1.
    EVEN = 0.
    WIN = 1.


def winner_sets(games):
    """games: list of (flop_cards, turn_cards, river_cards, player_cards_dict).  One kernel launch for all
    games that need evaluation; returns a list of winner-name sets."""
    out = [None] * len(games)
    rows = []
    for gi, (flop, turn, river, players) in enumerate(games):
        if len(players) == 1:  # env/game.py:664-665: a lone player wins without evaluation
            out[gi] = set(players.keys())
        else:
            rows.append(gi)
    if rows:
        P = max(len(games[gi][3]) for gi in rows)
        boards = np.zeros(len(rows), dtype=np.int64)
        holes = np.zeros((len(rows), P), dtype=np.int64)
        live = np.zeros(len(rows), dtype=np.uint32)
        names = []
        for r, gi in enumerate(rows):
            flop, turn, river, players = games[gi]
            boards[r] = cards_to_mask(list(flop) + [turn[0], river[0]])
            nm = list(players.keys())
            names.append(nm)
            for p, name in enumerate(nm):
                hole = cards_to_mask(players[name])
                if hole & boards[r]:
                    raise ValueError("hole card duplicates a board card")
                holes[r, p] = hole
            live[r] = (1 << len(nm)) - 1
        win = showdown(boards, holes, live)
        for r, gi in enumerate(rows):
            out[gi] = {name for p, name in enumerate(names[r]) if (int(win[r]) >> p) & 1}
    return out


class GameEnv:
    """The two static methods of the reference's GameEnv that lie on the hot path."""

    @staticmethod
    def get_winner_set_v2(flop_cards, turn_cards, river_cards, player_cards_dict):
        return winner_sets([(flop_cards, turn_cards, river_cards, player_cards_dict)])[0]

    @staticmethod
    def generate_game_result(current_match_winner_set, player_game_result_dict):
        # sole winner -> WIN, shared minimum rank -> EVEN for every winner; others untouched
        result = GamePlayerResult.EVEN if len(current_match_winner_set) > 1 else GamePlayerResult.WIN
        for name in current_match_winner_set:
            player_game_result_dict[name] = result
