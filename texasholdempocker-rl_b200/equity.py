"""Equity estimation, showdown and determinisation on the B200.

  equity_mc    Monte-Carlo win/tie/loss counts for batches of game states (generalises the pure-MC plans of
               config/debug.yml:33-61 to any number of opponents).
  equity_plan  the reference's enumerate/sample plans (simulate_processes, env/workflow.py:150-199;
               config/default.yml:73-133), heads-up, bit-exact where the plan is exhaustive.
  showdown     batched get_winner_set_v2 (env/game.py:663-681).
  deal         uniform re-deal of hidden cards (GameEnv.new_random, env/game.py:119-163).

Counts are (win, tie, loss) per state; the reference's ``game_result_sum`` is win - loss and its
``winning_probability`` is (sum/n)*0.5+0.5 (env/workflow.py:247-251).
"""
import ctypes as C

import numpy as np

from . import _native as N
from .cardset import card_id, cards_to_mask, ids_to_masks
from .evaluator import _current_device, _is_cuda_tensor, _np_ptr, _stream_ptr, _use_device, torch

OPP_SHIFT = 61   # bits 61-63; bit 60 is the ace of spades


def pack_states(heroes, boards, n_opp=None):
    """heroes: [S][2] cards; boards: [S] lists of 0/3/4/5 cards (ids, text or Card objects).

    Returns int64[S,2] rows {known = hero | board, hero}; a per-state opponent count (1..7) may ride in
    bits 61-63 of the second word.
    """
    S = len(heroes)
    out = np.zeros((S, 2), dtype=np.int64)
    for i in range(S):
        h = cards_to_mask(heroes[i])
        if bin(h).count("1") != 2:
            raise ValueError("hero needs exactly two cards")
        b = cards_to_mask(boards[i]) if boards[i] is not None else 0
        if bin(b).count("1") not in (0, 3, 4, 5) or (h & b):
            raise ValueError("board must hold 0, 3, 4 or 5 cards distinct from the hero's")
        k = 0 if n_opp is None else int(n_opp[i] if hasattr(n_opp, "__len__") else n_opp)
        if not 0 <= k <= 7:
            raise ValueError("per-state opponent count must be 0..7")
        out[i, 0] = h | b
        out[i, 1] = np.uint64(h | (k << OPP_SHIFT)).astype(np.int64)
    return out


def random_states(n_states, seed, streets=(0, 3, 4, 5)):
    """Synthetic game states per SURVEY 8d config 3: street uniform over `streets`, hero + board drawn
    uniformly without replacement with torch.Generator().manual_seed(seed)."""
    g = torch.Generator().manual_seed(int(seed))
    perm = torch.rand(n_states, 52, generator=g).argsort(1).numpy()
    street = torch.randint(0, len(streets), (n_states,), generator=g).numpy()
    nb = np.asarray(streets)[street]
    hero = ids_to_masks(perm[:, :2])
    board = np.zeros(n_states, dtype=np.int64)
    for k in (3, 4, 5):
        sel = nb >= k
        board[sel] |= np.left_shift(np.int64(1), 16 * (perm[sel, 2 + k - 1] & 3) + (perm[sel, 2 + k - 1] >> 2))
    for k in (1, 2):  # flop cards 1 and 2 come with the third
        sel = nb >= 3
        board[sel] |= np.left_shift(np.int64(1), 16 * (perm[sel, 2 + k - 1] & 3) + (perm[sel, 2 + k - 1] >> 2))
    return np.stack([hero | board, hero], axis=1).astype(np.int64)


_CARD_BITS = np.uint64(0x1FFF1FFF1FFF1FFF)


def _validate_states(a, n_opp):
    """Host-side check of numpy states (device tensors are the caller's responsibility)."""
    u = a.view(np.uint64)
    known, hero = u[:, 0], u[:, 1] & _CARD_BITS
    over = (u[:, 1] >> np.uint64(61)).astype(np.int64)
    nk = np.bitwise_count(known)
    if ((known & ~_CARD_BITS) != 0).any() or not np.isin(nk, (2, 5, 6, 7)).all():
        raise ValueError("known must hold the hero's two cards plus 0, 3, 4 or 5 board cards")
    if (np.bitwise_count(hero) != 2).any() or ((hero & known) != hero).any():
        raise ValueError("hero must be two cards contained in known")
    k = np.where(over > 0, over, n_opp)
    if (k < 1).any() or (k > 8).any():
        raise ValueError("opponent count must be 1..8")


def equity_mc(states, n_opp, rollouts, seed, sample_offset=0, state_index_base=0, out=None):
    """states int64[S,2] (see pack_states) -> counts int64[S,3] = (win, tie, loss) over `rollouts` deals.

    The deal for (state s, sample i) depends only on (seed, state_index_base+s, sample_offset+i), so shards
    over the sample range add up to the single-call result exactly.  CUDA tensor in -> CUDA tensor out
    (current stream, accumulates into `out` if given); numpy in -> numpy out.
    """
    L = N.lib()
    if _is_cuda_tensor(states):
        t = states.contiguous()
        if t.dtype != torch.int64 or t.ndim != 2 or t.shape[1] != 2:
            raise ValueError("states must be int64[S,2]")
        if out is None:
            out = torch.zeros(t.shape[0], 3, dtype=torch.int64, device=t.device)
        with torch.cuda.device(t.device):
            _use_device(t.device.index)
            N.check(L.phe_equity_mc(C.c_void_p(t.data_ptr()), t.shape[0], int(n_opp), int(rollouts), int(sample_offset),
                                    int(seed) & (2**64 - 1), int(state_index_base), C.c_void_p(out.data_ptr()), _stream_ptr()))
        return out
    a = np.ascontiguousarray(np.asarray(states, dtype=np.int64))
    if a.ndim != 2 or a.shape[1] != 2:
        raise ValueError("states must be int64[S,2]")
    _validate_states(a, int(n_opp))
    res = np.zeros((a.shape[0], 3), dtype=np.uint64)
    _use_device(_current_device())
    N.check(L.phe_equity_mc_host(_np_ptr(a), a.shape[0], int(n_opp), int(rollouts), int(sample_offset),
                                 int(seed) & (2**64 - 1), int(state_index_base), _np_ptr(res)))
    return res.astype(np.int64)


def plan_array(plan):
    p = np.ascontiguousarray([[int(g), int(r), int(bool(e))] for g, r, e in plan], dtype=np.int32)
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError("plan rows are (gen_num, num_random, is_segment_end)")
    return p


def plan_num_deals(n_exist, plan):
    p = plan_array(plan)
    return int(N.lib().phe_plan_num_deals(int(n_exist), _np_ptr(p), p.shape[0]))


def equity_plan(exist, plan, seed=0, stream_base=0):
    """exist: uint8[T,n_exist] known card ids per task in deal order (hero 2, then revealed board);
    plan: rows (gen_num, num_random, is_segment_end) of simulation_recurrent_params.

    Returns counts int64[T,3]; num_estimation = row sum, game_result_sum = win - loss.  Raises ValueError for
    plans the reference would reject or that do not reach nine cards.
    """
    L = N.lib()
    p = plan_array(plan)
    cuda_in = _is_cuda_tensor(exist)
    e = exist if cuda_in else np.asarray(exist, dtype=np.uint8)
    if e.ndim != 2:
        raise ValueError("exist must be [T, n_exist]")
    T, n_exist = int(e.shape[0]), int(e.shape[1])
    if L.phe_plan_num_deals(n_exist, _np_ptr(p), p.shape[0]) == 0:
        raise ValueError(f"plan {plan!r} is not expressible for {n_exist} known cards")
    if cuda_in:
        padded = torch.zeros(T, 8, dtype=torch.uint8, device=e.device)
        padded[:, :n_exist] = e
        out = torch.zeros(T, 3, dtype=torch.int64, device=e.device)
        with torch.cuda.device(e.device):
            _use_device(e.device.index)
            N.check(L.phe_equity_plan(C.c_void_p(padded.data_ptr()), n_exist, T, _np_ptr(p), p.shape[0], int(seed) & (2**64 - 1),
                                      int(stream_base), C.c_void_p(out.data_ptr()), _stream_ptr()))
        return out
    padded = np.zeros((T, 8), dtype=np.uint8)
    padded[:, :n_exist] = e
    res = np.zeros((T, 3), dtype=np.uint64)
    _use_device(_current_device())
    N.check(L.phe_equity_plan_host(_np_ptr(padded), n_exist, T, _np_ptr(p), p.shape[0], int(seed) & (2**64 - 1), int(stream_base), _np_ptr(res)))
    return res.astype(np.int64)


def showdown(boards, holes, live=None, return_ranks=False):
    """boards int64[G] (5-card sets), holes int64[G,P] (2-card sets), live int32[G] bitmask of contenders.

    Returns winners int32[G] (bit p = player p holds the minimum rank among live players) and optionally the
    ranks int16[G,P].  Same tie and lone-player semantics as env/game.py:663-681.
    """
    L = N.lib()
    if _is_cuda_tensor(boards):
        b = boards.contiguous(); h = holes.contiguous()
        G, P = h.shape
        win = torch.empty(G, dtype=torch.int32, device=b.device)
        ranks = torch.empty(G, P, dtype=torch.int16, device=b.device) if return_ranks else None
        lv = live.contiguous() if live is not None else None
        with torch.cuda.device(b.device):
            _use_device(b.device.index)
            N.check(L.phe_showdown(C.c_void_p(b.data_ptr()), C.c_void_p(h.data_ptr()), C.c_void_p(lv.data_ptr()) if lv is not None else None,
                                   G, P, C.c_void_p(win.data_ptr()), C.c_void_p(ranks.data_ptr()) if return_ranks else None, _stream_ptr()))
        return (win, ranks) if return_ranks else win
    b = np.ascontiguousarray(boards, dtype=np.int64)
    h = np.ascontiguousarray(holes, dtype=np.int64)
    G, P = h.shape
    lv = np.ascontiguousarray(live, dtype=np.uint32) if live is not None else None
    win = np.zeros(G, dtype=np.uint32)
    ranks = np.zeros((G, P), dtype=np.uint16) if return_ranks else None
    _use_device(_current_device())
    N.check(L.phe_showdown_host(_np_ptr(b), _np_ptr(h), _np_ptr(lv) if lv is not None else None, G, P, _np_ptr(win),
                                _np_ptr(ranks) if return_ranks else None))
    win = win.astype(np.int32)
    return (win, ranks) if return_ranks else win


def deal(known, need, seed, sample_offset=0, stream_base=0, by_sample=False):
    """known int64[B] card sets -> uint8[B,need] ids drawn uniformly without replacement from the rest.

    Row i draws (stream stream_base + i, sample sample_offset) -- or, with by_sample=True, (stream stream_base,
    sample sample_offset + i): n re-deals of ONE logical stream (phe_deal_samples)."""
    L = N.lib()
    dev_fn, host_fn = (L.phe_deal_samples, L.phe_deal_samples_host) if by_sample else (L.phe_deal, L.phe_deal_host)
    if _is_cuda_tensor(known):
        k = known.contiguous()
        out = torch.empty(k.shape[0], need, dtype=torch.uint8, device=k.device)
        with torch.cuda.device(k.device):
            _use_device(k.device.index)
            N.check(dev_fn(C.c_void_p(k.data_ptr()), k.shape[0], int(need), int(seed) & (2**64 - 1), int(sample_offset), int(stream_base),
                           C.c_void_p(out.data_ptr()), _stream_ptr()))
        return out
    k = np.ascontiguousarray(known, dtype=np.int64)
    out = np.zeros((k.shape[0], need), dtype=np.uint8)
    _use_device(_current_device())
    N.check(host_fn(_np_ptr(k), k.shape[0], int(need), int(seed) & (2**64 - 1), int(sample_offset), int(stream_base), _np_ptr(out)))
    return out


def win_probability(wtl):
    """(win + tie/2) / n -- identical to env/workflow.py:249-251's ((wins-losses)/n)*0.5+0.5."""
    w = np.asarray(wtl.cpu() if _is_cuda_tensor(wtl) else wtl, dtype=np.float64)
    n = w.sum(-1)
    return ((w[..., 0] - w[..., 2]) / n) * 0.5 + 0.5


def determinise(hero_cards, board_cards, n_players, n, seed, sample_offset=0, stream=0):
    """Batched GameEnv.new_random() (env/game.py:119-163): `n` uniform re-deals of everything the acting player cannot see.

    hero_cards: the acting player's two cards; board_cards: the revealed board (0, 3, 4 or 5 cards); n_players: seats
    at the table.  Returns (opp_holes uint8[n, n_players-1, 2] with each hand sorted as the reference sorts it,
    board uint8[n, 5] = revealed cards followed by the newly dealt ones, the dealt flop part sorted).  Re-deal i is
    sample (sample_offset + i) of Philox stream `stream` (phe_deal_samples): a caller that numbers its games / trees with
    `stream` never sees the same deal under two of them; results are numpy arrays (host).
    """
    hero = [card_id(c) for c in hero_cards]
    shown = [card_id(c) for c in (board_cards or [])]
    if len(hero) != 2 or len(shown) not in (0, 3, 4, 5) or len(set(hero + shown)) != len(hero) + len(shown):
        raise ValueError("need two hole cards and 0, 3, 4 or 5 distinct board cards")
    n_opp = int(n_players) - 1
    nbm = 5 - len(shown)
    need = 2 * n_opp + nbm
    if not 1 <= n_opp <= 8 or need > 21:
        raise ValueError("1..8 opponents")
    known = np.full(int(n), cards_to_mask(hero + shown), dtype=np.int64)
    ids = deal(known, need, seed, sample_offset=sample_offset, stream_base=stream, by_sample=True)
    odd = need & 1
    # slot order of phe_deal: [single if need is odd][pairs ...]; opponents take the first pairs, the board the rest
    opp = np.sort(ids[:, odd:odd + 2 * n_opp].reshape(n, n_opp, 2), axis=2)
    rest = np.concatenate([ids[:, :odd], ids[:, odd + 2 * n_opp:]], axis=1)
    if nbm == 5:
        rest = np.concatenate([np.sort(rest[:, :3], axis=1), rest[:, 3:]], axis=1)   # the reference sorts a new flop
    board = np.concatenate([np.tile(np.array(shown, dtype=np.uint8), (n, 1)).reshape(n, len(shown)), rest.astype(np.uint8)], axis=1)
    return opp.astype(np.uint8), board
