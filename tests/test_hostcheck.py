"""The arithmetic the sm_100a kernels execute (csrc/phe_core.cuh, table builder csrc/phe_tables.cpp),
instantiated on the host by the test-only libphe_hostcheck.so and compared bit-exactly with the oracle.
Covers every table entry, so a GPU/CPU difference can only come from the launch plumbing."""
import ctypes as C

import numpy as np
import pytest


def vp(a):
    return C.c_void_p(a.ctypes.data)


def test_table_image_every_entry(hostcheck, oracle):
    img = np.zeros(81920, dtype=np.uint16)
    hp = np.zeros(4, dtype=np.uint32)
    hostcheck.hc_get_image(vp(img), vp(hp))
    suits, flush, dp, noflush = oracle.tables()
    assert (img[73728:] == flush).all()                      # all 8192 flush lanes
    assert int((img[:65536] > 0).sum()) == 49205 == hostcheck.hc_num_keys()
    assert sorted(img[:65536][img[:65536] > 0].tolist()) == sorted(noflush.tolist())   # same multiset of ranks


def test_every_rank_multiset(hostcheck, oracle):
    # one representative hand per non-flush rank multiset (49205) and per flush lane (4719)
    hands = []

    def rec(q, pos, left):
        if pos == 13:
            if left == 0:
                c, n = [], 0
                for r in range(13):
                    for _ in range(q[r]):
                        c.append(r * 4 + (n & 3)); n += 1
                hands.append(c)
            return
        for d in range(min(4, left) + 1):
            q[pos] = d
            rec(q, pos + 1, left - d)
        q[pos] = 0
    rec([0] * 13, 0, 7)
    for m in range(8192):
        pc = bin(m).count("1")
        if 5 <= pc <= 7:
            c = [r * 4 + 3 for r in range(13) if m >> r & 1]
            for r in range(13):
                if len(c) < 7 and not (m >> r & 1):
                    c.append(r * 4 + (len(c) & 1))
            hands.append(c)
    ids = np.array(hands, dtype=np.uint8)
    masks = oracle.masks_from_ids(ids)
    out = np.zeros(len(masks), dtype=np.uint16)
    hostcheck.hc_eval7_masks(vp(masks), len(masks), vp(out))
    assert (out == oracle.rank7_batch(ids)).all()
    assert (out == oracle.rank7_batch(ids, brute=True)).all()


def test_random_hands(hostcheck, oracle):
    rng = np.random.default_rng(5)
    ids = np.argsort(rng.random((500000, 52)), axis=1)[:, :7].astype(np.uint8)
    masks = oracle.masks_from_ids(ids)
    out = np.zeros(len(masks), dtype=np.uint16)
    hostcheck.hc_eval7_masks(vp(masks), len(masks), vp(out))
    assert (out == oracle.rank7_batch(ids)).all()


def test_deal_stream(hostcheck, oracle):
    rng = np.random.default_rng(6)
    for i in range(500):
        k = int(rng.integers(0, 8))
        known = oracle.mask_of(rng.permutation(52)[:k])
        need = int(rng.integers(1, 22))
        got = np.zeros(need, dtype=np.uint8)
        hostcheck.hc_deal(known, 1000 + i, i * 977, i, need, vp(got))
        assert (got == oracle.deal(known, 1000 + i, i * 977, i, need)).all()


@pytest.mark.parametrize("text,n_opp", [("As Ah", 1), ("As Ah 2c 7d Tc", 5), ("7h 8h 9h Th 2c 2d", 3),
                                        ("Kd Kc 2c 7d Tc 3h 9s", 5), ("2c 7d", 8)])
def test_mc_rollouts(hostcheck, oracle, text, n_opp):
    ids = oracle.ids(text)
    hero, known = oracle.mask_of(ids[:2]), oracle.mask_of(ids)
    w = np.zeros(3, dtype=np.uint64)
    hostcheck.hc_equity_mc(known, hero, n_opp, 5000, 123, 99, 3, vp(w))
    assert (w.astype(np.int64) == oracle.equity_mc(known, hero, n_opp, 5000, 99, sample_offset=123, state_idx=3)).all()


def test_colex_unrank(hostcheck, oracle):
    rng = np.random.default_rng(7)
    for idx in [0, 1, 2, 133784559, 77777777] + rng.integers(0, 133784560, 300).tolist():
        c = np.zeros(7, dtype=np.int32)
        hostcheck.hc_colex_unrank7(int(idx), vp(c))
        assert (c == oracle.colex_unrank7(int(idx))).all()


@pytest.mark.parametrize("text,rnd", [("As Ah 2c 7d Tc 3h 9s", 3), ("As Ah 2c 7d Tc 3h", 2), ("As Ah 2c 7d Tc", 1),
                                      ("As Ah", 0), ("2c 3c", 0), ("Ks 2d 3c 4c 5c", 1), ("Kh Ks Ah Kd Kc 2c", 2)])
def test_plans(hostcheck, oracle, text, rnd):
    for plans in (oracle.DEFAULT_PLANS, oracle.DEBUG_PLANS):
        plan = plans[rnd][0]
        e = np.array(oracle.ids(text), dtype=np.uint8)
        p = np.array([[a, b, int(c)] for a, b, c in plan], dtype=np.int32)
        out = np.zeros(4, dtype=np.uint64)
        assert hostcheck.hc_equity_plan(vp(e), len(e), vp(p), len(p), 5, 2, vp(out)) == 0
        assert tuple(int(x) for x in out) == oracle.equity_plan(oracle.ids(text), plan, seed=5, stream=2)
        assert int(out[3]) == plans[rnd][1] == hostcheck.hc_plan_num_deals(len(e), vp(p), len(p))


@pytest.mark.parametrize("b,e", [(0, 1), (0, 7), (5, 6), (133784559, 133784560), (12345, 1234567), (133000000, 133784560)])
def test_enum_block_walk(hostcheck, oracle, b, e):
    """phe_enum.cuh: block setup / successor / per-hand rank, walked serially the way one warp walks its slice."""
    hostcheck.hc_enum7.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    h = np.zeros(7463, dtype=np.uint64)
    hostcheck.hc_enum7(b, e, vp(h))
    o9, oh = oracle.enum7(b, e, nthreads=2)
    assert (h.astype(np.int64) == oh).all()


def test_enum_block_walk_full(hostcheck):
    import json, os
    from conftest import GOLDEN
    hostcheck.hc_enum7.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p]
    h = np.zeros(7463, dtype=np.uint64)
    hostcheck.hc_enum7(0, 133784560, vp(h))
    assert h.tolist() == json.load(open(os.path.join(GOLDEN, "hist7463.json")))["hist7463"]


def test_philox_round_keys(hostcheck, oracle):
    """phe_core.cuh: Philox with precomputed round keys (the MC kernel's launch parameter) == Philox with the seed as key ==
    the oracle's, on the Random123 vectors and on random counters / seeds."""
    hostcheck.hc_philox.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    hostcheck.hc_philox.restype = None
    rng = np.random.default_rng(7)
    cases = [([0] * 4, 0), ([0xffffffff] * 4, 0xffffffffffffffff), ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], 0x299f31d0a4093822)]
    cases += [(rng.integers(0, 2**32, 4).tolist(), int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))) for _ in range(200)]
    for ctr, seed in cases:
        c = np.array(ctr, dtype=np.uint32)
        out = np.zeros(8, dtype=np.uint32)
        hostcheck.hc_philox(vp(c), seed, vp(out))
        want = oracle.philox4x32_10(ctr, [seed & 0xffffffff, seed >> 32]).tolist()
        assert out[:4].tolist() == want and out[4:].tolist() == want


def test_enum_work_units(hostcheck):
    """build_enum_units (phe_tables.cpp): the enumeration kernel's tiles.  Units partition [0, C(52,7)), hold at most 512 hands
    and at most 4 whole enumeration blocks (a block = one (c3..c6) suffix = C(c3,3) consecutive colex indices)."""
    from math import comb
    hostcheck.hc_enum_units.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int64]
    hostcheck.hc_enum_units.restype = C.c_int64
    buf = np.zeros(400000, dtype=np.uint32)
    starts = np.zeros(400000, dtype=np.uint64)
    n = hostcheck.hc_enum_units(512, 4, 128, vp(buf), vp(starts), len(buf))
    u = buf[:n].astype(np.int64)
    assert u[0] == 0 and u[-1] == 133784560 and (np.diff(u) > 0).all() and np.diff(u).max() <= 512
    sizes = np.array([comb(c3, 3) for c6 in range(6, 52) for c5 in range(5, c6) for c4 in range(4, c5) for c3 in range(3, c4)], dtype=np.int64)
    ends = np.cumsum(sizes)                                            # block ends on the colex axis
    starts_in = np.searchsorted(ends, u[1:], side="right") - np.searchsorted(ends, u[:-1], side="right")   # block ends inside each unit
    assert starts_in.max() <= 4
    assert n < 340000                                                  # ~1.3 MB + 2.6 MB tables
    # the packed start of every unit is the colex unranking of its first index
    st = starts[:n - 1].astype(np.int64)
    blk, tri = st & ((1 << 18) - 1), st >> 18
    block_start = np.concatenate([[0], ends[:-1]])                     # block id = position in colex order of the (c3..c6) suffixes
    assert blk.max() == len(sizes) - 1 == 211876 - 1 and (np.diff(blk) >= 0).all()
    assert (tri < sizes[blk]).all() and (block_start[blk] + tri == u[:-1]).all()
    # a unit that resumes a block it did not begin starts on a whole 128-hand trip of that block, unless the previous unit was
    # too short to reach one (then the cut could not advance)
    resumed = tri > 0
    assert (tri[resumed] % 128 == 0).mean() > 0.99


def test_multiply_shift_reductions_are_exact():
    """phe_core.cuh mod1326_u16 / mod52_u16: exact for every 16-bit input."""
    r = np.arange(65536, dtype=np.uint64)
    assert ((r - (((r >> 1) * 50611) >> 25) * 1326) == r % 1326).all() and (r.max() * 50611 < 2 ** 32)
    assert ((r - ((r * 40330) >> 21) * 52) == r % 52).all()
