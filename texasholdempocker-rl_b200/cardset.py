"""Card ids and 64-bit card sets.

Mirrors the reference's card contract: ``card_to_evaluator(card) = figure*4 + decor``
(utils/pokerhandevaluator_utils.py:3-4), figures 0 ('2') .. 12 ('A') and decors 0 C, 1 D, 2 H, 3 S
(env/constants.py:14-36); ``DECK[i]`` has id ``i`` like the reference's ``deck`` (env/game.py:14-18).
A card *set* is a 64-bit word with bit ``16*decor + figure`` (four 16-bit suit lanes), the layout every
kernel consumes.
"""
from dataclasses import dataclass
from enum import IntEnum

import numpy as np

NUM_CARDS = 52
FIGURE_CHARS = "23456789TJQKA"
DECOR_CHARS = "CDHS"


class CardFigure(IntEnum):
    CARD_FIGURE_2 = 0
    CARD_FIGURE_3 = 1
    CARD_FIGURE_4 = 2
    CARD_FIGURE_5 = 3
    CARD_FIGURE_6 = 4
    CARD_FIGURE_7 = 5
    CARD_FIGURE_8 = 6
    CARD_FIGURE_9 = 7
    CARD_FIGURE_10 = 8
    CARD_FIGURE_JACK = 9
    CARD_FIGURE_QUEEN = 10
    CARD_FIGURE_KING = 11
    CARD_FIGURE_ACE = 12


class CardDecor(IntEnum):
    CLUB = 0
    DIAMOND = 1
    HEART = 2
    SPADE = 3


@dataclass(frozen=True, order=True)
class Card:
    """Ordered by (figure, decor) like the reference's Card.__cmp__ (env/cards.py:7-34)."""
    figure: CardFigure
    decor: CardDecor

    def __str__(self):
        return DECOR_CHARS[self.decor] + FIGURE_CHARS[self.figure]

    @staticmethod
    def from_id(i):
        return Card(CardFigure(i >> 2), CardDecor(i & 3))


DECK = [Card(f, d) for f in CardFigure for d in CardDecor]


def card_to_evaluator(card):
    """Same value as utils/pokerhandevaluator_utils.py:3-4 for any object with .figure/.decor enums."""
    return card.figure.value * 4 + card.decor.value


def card_id(c):
    """int id, 'As'-style text (phevaluator accepts both), or a Card-like object -> id in [0, 52).

    Card objects (ours or the reference's env/cards.py Card) get their id memoised on the object: the Enum
    ``.value`` descriptor look-ups of card_to_evaluator are the hot spot of the equity service's host side."""
    if type(c) is int:
        i = c
    elif hasattr(c, "_phe_id"):
        return c._phe_id
    elif isinstance(c, (int, np.integer)):
        i = int(c)
    elif isinstance(c, str):
        if len(c) != 2 or c[0].upper() not in FIGURE_CHARS or c[1].upper() not in DECOR_CHARS:
            raise ValueError(f"bad card text {c!r}")
        i = FIGURE_CHARS.index(c[0].upper()) * 4 + DECOR_CHARS.index(c[1].upper())
    else:
        i = card_to_evaluator(c)
        if 0 <= i < NUM_CARDS:
            try:
                object.__setattr__(c, "_phe_id", i)
            except Exception:  # noqa: BLE001  (objects with __slots__: just recompute next time)
                pass
    if not 0 <= i < NUM_CARDS:
        raise ValueError(f"card id {i} out of range")
    return i


def id_to_bit(i):
    return 16 * (i & 3) + (i >> 2)


def bit_to_id(b):
    return (b & 15) * 4 + (b >> 4)


def cards_to_mask(cards):
    """Iterable of cards -> card-set word; ValueError on duplicates (phevaluator's contract is distinct cards)."""
    m = 0
    for c in cards:
        bit = 1 << id_to_bit(card_id(c))
        if m & bit:
            raise ValueError("duplicate card in hand")
        m |= bit
    return m


def mask_to_ids(m):
    out = []
    m = int(m)
    while m:
        low = m & -m
        out.append(bit_to_id(low.bit_length() - 1))
        m ^= low
    return out


def ids_to_masks(ids):
    """int array [..., k] of card ids -> int64 array [...] of card-set words (numpy)."""
    a = np.asarray(ids, dtype=np.int64)
    if a.size and (a.min() < 0 or a.max() >= NUM_CARDS):
        raise ValueError("card id out of range")
    bits = np.left_shift(np.int64(1), 16 * (a & 3) + (a >> 2))
    m = np.bitwise_or.reduce(bits, axis=-1)
    if a.size and not (np.bitwise_count(m.astype(np.uint64)) == a.shape[-1]).all():
        raise ValueError("duplicate card in hand")
    return m
