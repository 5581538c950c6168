"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product package never does.
See oracle/phe_oracle.c for what each entry point restates (reference file:line).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

NUM_HANDS_7 = 133784560          # C(52,7)
# 7-card category histogram (SF, quads, FH, flush, straight, trips, two pair, pair, high)
K1_HIST9 = [41584, 224848, 3473184, 4047644, 6180020, 6461620, 31433400, 58627800, 23294460]
# last rank of each category
K2_BOUNDS = [10, 166, 322, 1599, 1609, 2467, 3325, 6185, 7462]


def build(force=False):
    src = os.path.join(_HERE, "phe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u16p, u32p, u64p, i32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_int32))
        L.orc_init.restype = C.c_int
        L.orc_num_classes.restype = C.c_int
        L.orc_rank7_brute.argtypes = [u8p]; L.orc_rank7_brute.restype = C.c_int
        L.orc_rank7.argtypes = [u8p]; L.orc_rank7.restype = C.c_int
        L.orc_rank5_brute.argtypes = [u8p]; L.orc_rank5_brute.restype = C.c_int
        L.orc_rank7_batch.argtypes = [u8p, C.c_int64, u16p, C.c_int]
        L.orc_rank7_masks.argtypes = [u64p, C.c_int64, u16p]
        L.orc_enum7.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p, C.c_int]
        L.orc_colex_unrank7.argtypes = [C.c_uint64, u8p]
        L.orc_winner_mask.argtypes = [u8p, u8p, C.c_int, C.c_uint32, u16p]; L.orc_winner_mask.restype = C.c_uint32
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_deal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, u8p]
        L.orc_equity_mc.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, u64p]
        L.orc_equity_mc_indep.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, u64p]
        L.orc_equity_plan.argtypes = [u8p, C.c_int, i32p, C.c_int, C.c_uint64, C.c_uint32, u64p]
        L.orc_equity_plan.restype = C.c_int
        L.orc_rank7_sum_mt.argtypes = [u8p, C.c_int64, C.c_int]; L.orc_rank7_sum_mt.restype = C.c_uint64
        L.orc_get_tables.argtypes = [u8p, u16p, u32p, u16p]
        assert L.orc_init() == 0, "oracle class table did not come out as 7462 classes"
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# ---- card helpers (ids per utils/pokerhandevaluator_utils.py:3-4) ----
_RANKS = "23456789TJQKA"
_SUITS = "cdhs"


def card_id(s):
    """'As' -> 51, '2c' -> 0."""
    return _RANKS.index(s[0].upper()) * 4 + _SUITS.index(s[1].lower())


def ids(text):
    return [card_id(t) for t in text.split()]


def id_to_bit(i):
    return 16 * (i & 3) + (i >> 2)


def mask_of(id_list):
    m = 0
    for i in id_list:
        m |= 1 << id_to_bit(int(i))
    return m


def masks_from_ids(arr):
    arr = np.asarray(arr, dtype=np.int64)
    bits = 16 * (arr & 3) + (arr >> 2)
    return np.bitwise_or.reduce(np.left_shift(np.uint64(1), bits.astype(np.uint64)), axis=-1)


# ---- evaluators ----
def rank7(cards, brute=False):
    a = np.ascontiguousarray(cards, dtype=np.uint8)
    assert a.shape == (7,)
    f = lib().orc_rank7_brute if brute else lib().orc_rank7
    return int(f(_p(a, C.c_uint8)))


def rank5_brute(cards):
    a = np.ascontiguousarray(cards, dtype=np.uint8)
    return int(lib().orc_rank5_brute(_p(a, C.c_uint8)))


def rank7_batch(cards, brute=False):
    a = np.ascontiguousarray(cards, dtype=np.uint8).reshape(-1, 7)
    out = np.empty(a.shape[0], dtype=np.uint16)
    lib().orc_rank7_batch(_p(a, C.c_uint8), a.shape[0], _p(out, C.c_uint16), int(brute))
    return out


def rank7_masks(masks):
    a = np.ascontiguousarray(masks, dtype=np.uint64).reshape(-1)
    out = np.empty(a.shape[0], dtype=np.uint16)
    lib().orc_rank7_masks(_p(a, C.c_uint64), a.shape[0], _p(out, C.c_uint16))
    return out


def enum7(begin=0, end=NUM_HANDS_7, nthreads=None):
    h9 = np.zeros(9, dtype=np.uint64)
    h = np.zeros(7463, dtype=np.uint64)
    lib().orc_enum7(begin, end, _p(h9, C.c_uint64), _p(h, C.c_uint64), nthreads or os.cpu_count() or 1)
    return h9.astype(np.int64), h.astype(np.int64)


def colex_unrank7(idx):
    out = np.zeros(7, dtype=np.uint8)
    lib().orc_colex_unrank7(idx, _p(out, C.c_uint8))
    return out


def winner_mask(board, holes, live=None, with_ranks=False):
    b = np.ascontiguousarray(board, dtype=np.uint8)
    h = np.ascontiguousarray(holes, dtype=np.uint8).reshape(-1, 2)
    P = h.shape[0]
    if live is None:
        live = (1 << P) - 1
    ranks = np.zeros(P, dtype=np.uint16)
    w = int(lib().orc_winner_mask(_p(b, C.c_uint8), _p(h, C.c_uint8), P, live, _p(ranks, C.c_uint16)))
    return (w, ranks) if with_ranks else w


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


def deal(known_mask, seed, sample, stream, need):
    out = np.zeros(need, dtype=np.uint8)
    lib().orc_deal(known_mask, seed, sample, stream, need, _p(out, C.c_uint8))
    return out


def equity_mc(known_mask, hero_mask, n_opp, rollouts, seed, sample_offset=0, state_idx=0):
    out = np.zeros(3, dtype=np.uint64)
    lib().orc_equity_mc(known_mask, hero_mask, n_opp, rollouts, sample_offset, seed, state_idx, _p(out, C.c_uint64))
    return out.astype(np.int64)


def equity_mc_indep(known_mask, hero_mask, n_opp, rollouts, seed):
    out = np.zeros(3, dtype=np.uint64)
    lib().orc_equity_mc_indep(known_mask, hero_mask, n_opp, rollouts, seed, _p(out, C.c_uint64))
    return out.astype(np.int64)


def equity_plan(exist_ids, plan, seed=0, stream=0):
    """exist_ids: hero(2) + known board ids in deal order; plan: [(gen_num, num_random, is_segment_end), ...]
    (config/default.yml:73-133).  Returns (wins, ties, losses, n)."""
    e = np.ascontiguousarray(exist_ids, dtype=np.uint8)
    p = np.ascontiguousarray([[int(a), int(b), int(bool(c))] for a, b, c in plan], dtype=np.int32)
    out = np.zeros(4, dtype=np.uint64)
    rc = lib().orc_equity_plan(_p(e, C.c_uint8), e.shape[0], _p(p, C.c_int32), p.shape[0], seed, stream, _p(out, C.c_uint64))
    if rc != 0:
        raise ValueError(f"bad plan for {e.shape[0]} known cards (rc={rc})")
    return tuple(int(x) for x in out)


def rank7_sum_mt(cards, nthreads):
    a = np.ascontiguousarray(cards, dtype=np.uint8).reshape(-1, 7)
    return int(lib().orc_rank7_sum_mt(_p(a, C.c_uint8), a.shape[0], nthreads))


def tables():
    suits = np.zeros(4609, dtype=np.uint8)
    flush = np.zeros(8192, dtype=np.uint16)
    dp = np.zeros((5, 14, 10), dtype=np.uint32)
    nof = np.zeros(49205, dtype=np.uint16)
    lib().orc_get_tables(_p(suits, C.c_uint8), _p(flush, C.c_uint16), _p(dp, C.c_uint32), _p(nof, C.c_uint16))
    return suits, flush, dp, nof


# reference default plans (config/default.yml:73-133) and debug plans (config/debug.yml:33-61)
DEFAULT_PLANS = {
    0: ([(1, 0, False), (1, 0, False), (1, 0, True), (4, 5, True)], 98000),
    1: ([(1, 0, True), (1, 0, True), (2, 45, True)], 97290),
    2: ([(1, 0, True), (1, 0, False), (1, 0, True)], 45540),
    3: ([(1, 0, False), (1, 0, True)], 990),
}
DEBUG_PLANS = {
    0: ([(7, 100, True)], 100),
    1: ([(4, 100, True)], 100),
    2: ([(3, 100, True)], 100),
    3: ([(2, 100, True)], 100),
}
