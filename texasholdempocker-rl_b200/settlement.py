"""Terminal settlement on the B200: batched ``finish_game`` / ``_share_value`` (env/game.py:410-441, 560-632, 683-734).

``settle`` takes seat-indexed arrays for many finished games and returns chips won, game result and pure card result
per seat.  ``settle_env`` reads a reference-shaped ``GameEnv`` (the attributes ``_share_value`` reads) and returns the
reference's ``(winner_value_dict, player_game_result_dict)`` plus the card results of ``finish_game``.
"""
import ctypes as C

import numpy as np

from . import _native as N
from .cardset import cards_to_mask
from .evaluator import _current_device, _is_cuda_tensor, _np_ptr, _stream_ptr, _use_device, torch
from .winner import GamePlayerResult

ACT_NONE, ACT_FOLD, ACT_LIVE = 0, 1, 2
MAX_ROUNDS = 4
_RESULT = {-1: GamePlayerResult.LOSE, 0: GamePlayerResult.EVEN, 1: GamePlayerResult.WIN}


def settle(boards, holes, actions, values, n_rounds, button, onboard, small_blind):
    """boards int64[G]; holes int64[G,P]; actions int8[G,R,P] (0 None, 1 FOLD, 2 other); values int32[G,R,P];
    n_rounds int32[G]; button int32[G]; onboard int32[G] (bit p = seat p still ONBOARD).

    Returns (won int32[G,P], result int8[G,P], card_result int8[G,P], paid int32[G]); results are -1 / 0 / 1 for
    LOSE / EVEN / WIN.  CUDA tensors in -> CUDA tensors out on the current stream; numpy in -> numpy out."""
    L = N.lib()
    if _is_cuda_tensor(boards):
        b, h, a, v = boards.contiguous(), holes.contiguous(), actions.contiguous(), values.contiguous()
        nr, bt, ob = n_rounds.contiguous(), button.contiguous(), onboard.contiguous()
        G, P = h.shape
        R = a.shape[1] if a.dim() == 3 else -1
        if (a.shape != (G, R, P) or v.shape != (G, R, P) or not 1 <= R <= MAX_ROUNDS or b.shape != (G,) or nr.shape != (G,) or bt.shape != (G,)
                or ob.shape != (G,) or (b.dtype, h.dtype, a.dtype, v.dtype, nr.dtype, bt.dtype) !=
                (torch.int64, torch.int64, torch.int8, torch.int32, torch.int32, torch.int32) or ob.dtype not in (torch.int32, torch.uint32)):
            raise ValueError("boards int64[G], holes int64[G,P], actions int8[G,R<=4,P], values int32[G,R,P], n_rounds / button / onboard int32[G]")
        dev = b.device
        won = torch.empty(G, P, dtype=torch.int32, device=dev)
        res = torch.empty(G, P, dtype=torch.int8, device=dev)
        card = torch.empty(G, P, dtype=torch.int8, device=dev)
        paid = torch.empty(G, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _use_device(dev.index)
            N.check(L.phe_settle(*(C.c_void_p(t.data_ptr()) for t in (b, h, a, v, nr, bt, ob)), G, P, R, int(small_blind),
                                 *(C.c_void_p(t.data_ptr()) for t in (won, res, card, paid)), _stream_ptr()))
        return won, res, card, paid
    b = np.ascontiguousarray(boards, dtype=np.int64)
    h = np.ascontiguousarray(holes, dtype=np.int64)
    a = np.ascontiguousarray(actions, dtype=np.int8)
    v = np.ascontiguousarray(values, dtype=np.int32)
    nr = np.ascontiguousarray(n_rounds, dtype=np.int32)
    bt = np.ascontiguousarray(button, dtype=np.int32)
    ob = np.ascontiguousarray(onboard, dtype=np.uint32)
    G, P = h.shape
    if a.ndim != 3 or a.shape != (G, a.shape[1], P) or v.shape != a.shape or not 1 <= a.shape[1] <= MAX_ROUNDS:
        raise ValueError("actions / values must be [G, R<=4, P]")
    if b.shape != (G,) or nr.shape != (G,) or bt.shape != (G,) or ob.shape != (G,):
        raise ValueError("boards, n_rounds, button, onboard must be [G]")
    if G and (bt.min() < 0 or bt.max() >= P or nr.min() < 0 or nr.max() > a.shape[1]):
        raise ValueError("button must be a seat index in [0, P) and n_rounds within [0, R]")
    won = np.zeros((G, P), dtype=np.int32)
    res = np.zeros((G, P), dtype=np.int8)
    card = np.zeros((G, P), dtype=np.int8)
    paid = np.zeros(G, dtype=np.uint32)
    R = a.shape[1]
    _use_device(_current_device())
    N.check(L.phe_settle_host(*(_np_ptr(t) for t in (b, h, a, v, nr, bt, ob)), G, P, R, int(small_blind),
                              *(_np_ptr(t) for t in (won, res, card, paid))))
    return won, res, card, paid.astype(np.int32)


def _act_code(action):
    if action is None:
        return ACT_NONE
    return ACT_FOLD if getattr(action, "name", action) == "FOLD" else ACT_LIVE


def env_arrays(env):
    """Seat-indexed arrays of one reference-shaped GameEnv at its terminal state."""
    names = list(env.players.keys())
    P = len(names)
    board = cards_to_mask(list(env.flop_cards) + list(env.turn_cards) + list(env.river_cards))
    holes = np.array([cards_to_mask(env.info_sets[n].player_hand_cards) for n in names], dtype=np.int64)
    R = len(env.historical_round_action_list)
    act = np.zeros((MAX_ROUNDS, P), dtype=np.int8)
    val = np.zeros((MAX_ROUNDS, P), dtype=np.int32)
    for r, (ad, vd) in enumerate(zip(env.historical_round_action_list, env.historical_round_value_list)):
        for p, n in enumerate(names):
            act[r, p] = _act_code(ad.get(n))
            val[r, p] = int(vd.get(n, 0))
    onboard = sum(1 << p for p, n in enumerate(names) if getattr(env.players[n].status, "name", "") == "ONBOARD")
    return names, board, holes, act, val, R, names.index(env.button_player_name), onboard


def settle_envs(envs):
    """Settle many finished games in one launch; each result is (winner_value_dict, player_game_result_dict,
    card_result_dict) keyed by player name, as finish_game / _share_value produce them."""
    if not envs:
        return []
    rows = [env_arrays(e) for e in envs]
    P = len(rows[0][0])
    sb = envs[0].small_blind
    if any(len(r[0]) != P for r in rows) or any(e.small_blind != sb for e in envs):
        raise ValueError("one call settles games of one table size and one small blind")
    won, res, card, paid = settle(np.array([r[1] for r in rows], dtype=np.int64), np.stack([r[2] for r in rows]),
                                  np.stack([r[3] for r in rows]), np.stack([r[4] for r in rows]),
                                  np.array([r[5] for r in rows], dtype=np.int32), np.array([r[6] for r in rows], dtype=np.int32),
                                  np.array([r[7] for r in rows], dtype=np.uint32), sb)
    out = []
    for g, r in enumerate(rows):
        names = r[0]
        out.append(({n: int(won[g, p]) for p, n in enumerate(names) if int(paid[g]) >> p & 1},
                    {n: _RESULT[int(res[g, p])] for p, n in enumerate(names)},
                    {n: _RESULT[int(card[g, p])] for p, n in enumerate(names)}))
    return out


def settle_env(env):
    return settle_envs([env])[0]
