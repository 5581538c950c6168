"""Pure-Python stand-in for the PyPI package ``phevaluator`` (TEST INFRASTRUCTURE ONLY).

The reference imports ``from phevaluator import evaluate_cards`` (env/game.py:10) but the package is neither
vendored nor installable offline.  This shim restates the package's published 7-card algorithm in pure
Python -- suit hash -> SUITS; flush -> FLUSH[8192]; else base-5 rank counts -> hash_quinary (DP) ->
NO_FLUSH_7[49205] -- on tables filled by the oracle's first-principles evaluator (oracle/phe_oracle.c).
It exists so that (a) the reference's own env/game.py and env/workflow.py can be imported in the build
container to generate golden vectors (tests/golden/make_golden.py) and (b) bench.py --impl reference can time
a faithful pure-Python evaluator on the GPU box's host cores.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.abspath(os.path.join(_here, "..", "..", ".."))
if _repo not in sys.path:
    sys.path.insert(0, _repo)

from oracle import oracle as _O  # noqa: E402

_suits, _flush, _dp, _noflush = _O.tables()
SUITS = _suits.tolist()
FLUSH = _flush.tolist()
DP = _dp.tolist()
NO_FLUSH_7 = _noflush.tolist()
SUITBIT_BY_ID = [0x1, 0x8, 0x40, 0x200] * 13
BINARIES_BY_ID = [1 << (i >> 2) for i in range(52)]
_RANK = "23456789TJQKA"
_SUIT = "cdhs"


def _to_id(c):
    if isinstance(c, int):
        return c
    if isinstance(c, str):
        return _RANK.index(c[0].upper()) * 4 + _SUIT.index(c[1].lower())
    return c.id_


def hash_quinary(quinary, num_cards):
    total = 0
    length = len(quinary)
    for rank, cnt in enumerate(quinary):
        if cnt:
            total += DP[cnt][length - rank - 1][num_cards]
            num_cards -= cnt
            if num_cards <= 0:
                break
    return total


def evaluate_cards(*cards):
    ids = [_to_id(c) for c in cards]
    if len(ids) != 7:
        raise ValueError("this shim restates the 7-card path only (the reference passes 5 board + 2 hole cards)")
    suit_hash = 0
    for c in ids:
        suit_hash += SUITBIT_BY_ID[c]
    flush_suit = SUITS[suit_hash]
    if flush_suit:
        suit_binary = [0, 0, 0, 0]
        for c in ids:
            suit_binary[c & 3] |= BINARIES_BY_ID[c]
        return FLUSH[suit_binary[flush_suit - 1]]
    quinary = [0] * 13
    for c in ids:
        quinary[c >> 2] += 1
    return NO_FLUSH_7[hash_quinary(quinary, 7)]
