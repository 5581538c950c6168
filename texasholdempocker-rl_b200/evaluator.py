"""Hand evaluation on the B200: drop-in ``evaluate_cards`` plus the batched entry points.

``evaluate_cards(*cards)`` returns what ``phevaluator.evaluate_cards`` returns for the reference's only
call (env/game.py:675): the rank of the best five of seven cards, 1 (royal flush) .. 7462, lower is
stronger.  Every call runs on the GPU through include/phe_b200.h; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import _native as N
from .cardset import card_id, cards_to_mask

try:  # torch is plumbing (device memory, streams); numpy callers do not need it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_cuda_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def _use_device(index):
    N.check(N.lib().phe_init(int(index)))


def _current_device():
    if torch is not None and torch.cuda.is_available():
        return torch.cuda.current_device()
    return 0


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


def evaluate_cards(*cards):
    """Rank of a 7-card hand; cards are ids (figure*4+decor), 'As'-style text, or Card objects.

    Raises ValueError for a hand that is not 7 distinct cards (phevaluator raises for sizes it does not
    support; the reference only ever passes 5 board + 2 hole cards).
    """
    if len(cards) != 7:
        raise ValueError(f"evaluate_cards on the B200 path takes exactly 7 cards, got {len(cards)}")
    m = np.array([cards_to_mask(cards)], dtype=np.int64)
    out = np.zeros(1, dtype=np.uint16)
    _use_device(_current_device())
    N.check(N.lib().phe_eval7_host(_np_ptr(m), N.LAYOUT_MASKS, 1, _np_ptr(out)))
    return int(out[0])


def _check_ids(a):
    if a.ndim != 2 or a.shape[1] != 7:
        raise ValueError("id layout wants shape [N, 7]")


def eval7(hands, out=None):
    """Batched 7-card evaluation.

    hands: uint8[N,7] card ids or int64[N] card-set words; a CUDA torch tensor (result: int16 CUDA tensor,
    launched on the current stream, not synchronised) or a numpy array / CPU tensor (result: numpy uint16,
    staged through the library's *_host entry point).
    """
    L = N.lib()
    if _is_cuda_tensor(hands):
        t = hands.contiguous()
        if t.dtype == torch.uint8:
            _check_ids(t)
            layout = N.LAYOUT_IDS
        elif t.dtype == torch.int64 and t.ndim == 1:
            layout = N.LAYOUT_MASKS
        else:
            raise ValueError("hands must be uint8[N,7] ids or int64[N] card sets")
        n = t.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.int16, device=t.device)
        with torch.cuda.device(t.device):
            _use_device(t.device.index)
            N.check(L.phe_eval7(C.c_void_p(t.data_ptr()), layout, n, C.c_void_p(out.data_ptr()), _stream_ptr()))
        return out
    a = hands.numpy() if (torch is not None and isinstance(hands, torch.Tensor)) else np.asarray(hands)
    if a.dtype == np.uint8:
        _check_ids(a)
        layout = N.LAYOUT_IDS
    elif a.dtype in (np.int64, np.uint64) and a.ndim == 1:
        layout = N.LAYOUT_MASKS
    else:
        raise ValueError("hands must be uint8[N,7] ids or int64[N] card sets")
    a = np.ascontiguousarray(a)
    res = np.empty(a.shape[0], dtype=np.uint16)
    _use_device(_current_device())
    N.check(L.phe_eval7_host(_np_ptr(a), layout, a.shape[0], _np_ptr(res)))
    return res


def enumerate_c52_7(comb_begin=0, comb_end=N.NUM_HANDS_7, device_out=False, part=0, nparts=1):
    """Exhaustive evaluation of the 7-card combinations with colex index in [comb_begin, comb_end).

    part / nparts select one of `nparts` interleaved, equally expensive parts of the range (device_out only); the
    parts' histograms add up to the whole range's.

    Returns (hist9 int64[9], hist7463 int64[7463]): category counts (SF, quads, full house, flush, straight,
    trips, two pair, pair, high card) and the per-rank histogram (bin = rank, bin 0 empty).  With
    device_out=True the histograms stay on the current CUDA device (torch int64) and nothing is synchronised.
    """
    L = N.lib()
    if device_out:
        buf = torch.zeros(9 + N.NUM_RANKS + 1, dtype=torch.int64, device="cuda")
        _use_device(buf.device.index)
        N.check(L.phe_enum7_part(C.c_void_p(buf.data_ptr()), C.c_void_p(buf.data_ptr() + 72), comb_begin, comb_end, int(part), int(nparts), _stream_ptr()))
        return buf[:9], buf[9:]
    if nparts != 1:
        raise ValueError("interleaved parts are a device_out feature")
    h9 = np.zeros(9, dtype=np.uint64)
    h = np.zeros(N.NUM_RANKS + 1, dtype=np.uint64)
    _use_device(_current_device())
    N.check(L.phe_enum7_host(_np_ptr(h9), _np_ptr(h), comb_begin, comb_end))
    return h9.astype(np.int64), h.astype(np.int64)


def hand_ids(cards):
    return [card_id(c) for c in cards]
