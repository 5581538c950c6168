"""Parity tests proper: the CUDA path (through the C ABI) against the oracle, the golden fixtures and
size-independent properties.  Integer work -> bit-exact; Monte Carlo -> bit-exact under the shared seed stream
and within 3 sigma of an independent CPU estimate."""
import json
import os
import queue

import numpy as np
import pytest

from conftest import GOLDEN
from test_oracle_kat import KATS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device; there is no CPU fallback to fall back to"
    torch.cuda.set_device(0)
    return torch


def rand_hands(n, seed):
    rng = np.random.default_rng(seed)
    return np.argsort(rng.random((n, 52)), axis=1)[:, :7].astype(np.uint8)


# ---- evaluate_cards / eval7 -------------------------------------------------------------------------
@pytest.mark.parametrize("text,rank", KATS)
def test_evaluate_cards_kats(phe, oracle, torch, text, rank):
    assert phe.evaluate_cards(*text.split()) == rank
    assert phe.evaluate_cards(*oracle.ids(text)) == rank
    assert phe.evaluate_cards(*[phe.DECK[i] for i in oracle.ids(text)]) == rank


def test_eval7_large_batch_both_layouts(phe, oracle, torch):
    ids = rand_hands(1 << 20, 1)
    want = oracle.rank7_batch(ids)
    got_ids = phe.eval7(torch.from_numpy(ids).cuda()).cpu().numpy().astype(np.uint16)
    masks = phe.ids_to_masks(ids)
    got_masks = phe.eval7(torch.from_numpy(masks).cuda()).cpu().numpy().astype(np.uint16)
    assert (got_ids == want).all()
    assert (got_masks == want).all()
    assert (phe.eval7(masks) == want).all()          # numpy in -> *_host entry point
    assert (phe.eval7(ids) == want).all()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 255, 257, 65535, 65536, 65537, 100003])
def test_eval7_ragged_sizes(phe, oracle, torch, n):
    ids = rand_hands(max(n, 1), 2)[:n]
    want = oracle.rank7_batch(ids) if n else np.zeros(0, dtype=np.uint16)
    got = phe.eval7(torch.from_numpy(phe.ids_to_masks(ids) if n else np.zeros(0, dtype=np.int64)).cuda())
    assert (got.cpu().numpy().astype(np.uint16) == want).all()
    got = phe.eval7(torch.from_numpy(ids.reshape(n, 7)).cuda())
    assert (got.cpu().numpy().astype(np.uint16) == want).all()


def test_eval7_unaligned_views(phe, oracle, torch):
    ids = rand_hands(70001, 3)
    masks = torch.from_numpy(phe.ids_to_masks(ids)).cuda()
    want = oracle.rank7_batch(ids)
    got = phe.eval7(masks[1:])                        # 8-byte but not 16-byte aligned input
    assert (got.cpu().numpy().astype(np.uint16) == want[1:]).all()
    dids = torch.from_numpy(ids).cuda()
    for off in (1, 2, 3, 4):                          # id rows start 7 bytes apart: only every fourth view is word aligned
        got = phe.eval7(dids[off:])
        assert (got.cpu().numpy().astype(np.uint16) == want[off:]).all()


def test_eval7_every_rank_multiset(phe, oracle, torch):
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
    ids = np.array(hands, dtype=np.uint8)
    assert len(ids) == 49205
    got = phe.eval7(torch.from_numpy(ids).cuda()).cpu().numpy().astype(np.uint16)
    assert (got == oracle.rank7_batch(ids, brute=True)).all()


def test_eval7_at_bench_size_property(phe, oracle, torch):
    # BASELINE micro-config: 2^25 hands.  Size-independent checks: permutation invariance of the rank under
    # card order, histogram of a replicated batch = replication of the histogram, first 2^20 exact.
    base = rand_hands(1 << 20, 4)
    masks = torch.from_numpy(phe.ids_to_masks(base)).cuda().repeat(32)
    got = phe.eval7(masks)
    assert got.shape[0] == 1 << 25
    first = got[: 1 << 20]
    assert (first.cpu().numpy().astype(np.uint16) == oracle.rank7_batch(base)).all()
    assert bool((got.view(32, -1) == first.unsqueeze(0)).all())
    shuffled = np.ascontiguousarray(base[:, ::-1])
    assert bool((phe.eval7(torch.from_numpy(shuffled).cuda()) == first).all())


# ---- exhaustive enumeration ----------------------------------------------------------------------------
def test_enum7_full_bit_exact(phe, oracle, torch):
    h9, h = phe.enumerate_c52_7()
    assert h9.tolist() == oracle.K1_HIST9
    gold = json.load(open(os.path.join(GOLDEN, "hist7463.json")))
    assert h.tolist() == gold["hist7463"]
    assert h[0] == 0 and int(h[7415:].sum()) == 0 and int(h.sum()) == 133784560


@pytest.mark.parametrize("b,e", [(0, 1), (0, 7), (5, 6), (133784559, 133784560), (1000, 1000), (12345, 1234567),
                                 (133000000, 133784560)])
def test_enum7_ranges_match_oracle(phe, oracle, torch, b, e):
    h9, h = phe.enumerate_c52_7(b, e)
    o9, oh = oracle.enum7(b, e, nthreads=2)
    assert (h9 == o9).all() and (h == oh).all()


def test_enum7_shards_tile(phe, oracle, torch):
    full9, full = phe.enumerate_c52_7()
    acc9, acc = np.zeros(9, dtype=np.int64), np.zeros(7463, dtype=np.int64)
    for r in range(8):
        b, e = phe.shard_range(133784560, r, 8)
        a9, a = phe.enumerate_c52_7(b, e)
        acc9 += a9; acc += a
    assert (acc9 == full9).all() and (acc == full).all()


def test_enum7_device_out_accumulates(phe, oracle, torch):
    h9, h = phe.enumerate_c52_7(0, 2_000_000, device_out=True)
    torch.cuda.synchronize()
    o9, oh = oracle.enum7(0, 2_000_000, nthreads=2)
    assert (h9.cpu().numpy() == o9).all() and (h.cpu().numpy() == oh).all()


# ---- Monte-Carlo equity ------------------------------------------------------------------------------------
def test_equity_mc_bit_exact_vs_oracle(phe, oracle, torch):
    states = phe.random_states(24, 20261019)
    for n_opp, rollouts in [(1, 3000), (5, 1500), (8, 700)]:
        got = phe.equity_mc(torch.from_numpy(states).cuda(), n_opp, rollouts, seed=0x5EED20261019, sample_offset=17,
                            state_index_base=5).cpu().numpy()
        for s in range(len(states)):
            want = oracle.equity_mc(int(states[s, 0]), int(states[s, 1]), n_opp, rollouts, 0x5EED20261019,
                                    sample_offset=17, state_idx=5 + s)
            assert (got[s] == want).all(), (n_opp, s)
        assert (got.sum(1) == rollouts).all()


@pytest.mark.parametrize("n_opp,rollouts", [(8, 20011), (5, 24997), (1, 30000)])
def test_equity_mc_large_launch_bit_exact_vs_oracle(phe, oracle, torch, n_opp, rollouts):
    """>= 1e6 rollouts per launch: the persistent shared-memory kernel (k_equity_mc_smem; also what the -DPHE_MC_POOL
    experiment build runs here).  Every street, ragged rollout counts, a sample offset, per-state opponent overrides, items
    interleaved over the states -- counts equal the oracle's, state by state."""
    states = phe.random_states(56, 77)
    over = np.zeros(56, dtype=np.int64)
    over[5::7] = (n_opp % 7) + 1                                     # some states carry their own opponent count
    states[:, 1] = (states[:, 1].astype(np.uint64) | (over.astype(np.uint64) << np.uint64(61))).astype(np.int64)
    before = phe._native.launch_count()
    got = phe.equity_mc(torch.from_numpy(states).cuda(), n_opp, rollouts, seed=0x5EED20261019, sample_offset=12345, state_index_base=9).cpu().numpy()
    assert phe._native.launch_count() == before + 1 and 56 * rollouts >= 1_000_000
    assert (got.sum(1) == rollouts).all()
    hero_mask = np.uint64(0x1FFF1FFF1FFF1FFF)
    for s in range(56):
        k = int(over[s]) or n_opp
        want = oracle.equity_mc(int(states[s, 0]), int(np.uint64(states[s, 1]) & hero_mask), k, rollouts, 0x5EED20261019, sample_offset=12345, state_idx=9 + s)
        assert (got[s] == want).all(), (s, k)


def test_equity_mc_config1_single_state(phe, oracle, torch):
    # heads-up, hero As Ah, pre-flop and on flop 2c 7d Tc, 10k rollouts, seed 20261019
    st = phe.pack_states([["As", "Ah"], ["As", "Ah"]], [[], ["2c", "7d", "Tc"]])
    got = phe.equity_mc(st, 1, 10000, seed=20261019)          # numpy -> host entry point
    assert (got.sum(1) == 10000).all()
    for s in range(2):
        want = oracle.equity_mc(int(st[s, 0]), int(st[s, 1]), 1, 10000, 20261019, state_idx=s)
        assert (got[s] == want).all()
        ind = oracle.equity_mc_indep(int(st[s, 0]), int(st[s, 1]), 1, 400000, seed=5)
        p_ref = (ind[0] + ind[1] / 2) / ind.sum()
        p = (got[s, 0] + got[s, 1] / 2) / 10000
        assert abs(p - p_ref) < 3 * (0.25 / 10000) ** 0.5      # 3 sigma binomial


def test_equity_mc_shards_and_grid_independence(phe, oracle, torch):
    # the same sample range split three ways (different launch shapes and chunk sizes) adds up exactly
    states = torch.from_numpy(phe.random_states(64, 7)).cuda()
    whole = phe.equity_mc(states, 5, 60000, seed=11)
    parts = torch.zeros_like(whole)
    for b, e in [(0, 100), (100, 33333), (33333, 60000)]:
        phe.equity_mc(states, 5, e - b, seed=11, sample_offset=b, out=parts)
    assert bool((whole == parts).all())
    one = phe.equity_mc(states[3:4], 5, 60000, seed=11, state_index_base=3)
    assert bool((one[0] == whole[3]).all())


def test_equity_mc_river_states_exact_enumeration(phe, oracle, torch):
    # river + one opponent: Monte Carlo must agree with the exact 990-deal enumeration within 3 sigma
    states = phe.random_states(200, 3, streets=(5,))
    R = 200000
    got = phe.equity_mc(torch.from_numpy(states).cuda(), 1, R, seed=1).cpu().numpy()
    for s in range(0, 200, 10):
        ids = phe.mask_to_ids(states[s, 1]) + phe.mask_to_ids(states[s, 0] & ~states[s, 1])
        w, t, l, n = oracle.equity_plan(ids, oracle.DEFAULT_PLANS[3][0])
        for k, exact in enumerate((w, t, l)):
            p = exact / 990
            assert abs(got[s, k] / R - p) <= 3 * (p * (1 - p) / R) ** 0.5 + 1e-9


def test_equity_mc_config3_sixmax_statistics(phe, oracle, torch):
    # 6-max states from the config-3 generator: every state within 3 sigma of the independent CPU estimator
    states = phe.random_states(4096, 20261019)
    R = 100000
    got = phe.equity_mc(torch.from_numpy(states).cuda(), 5, R, seed=0x5EED20261019).cpu().numpy()
    assert (got.sum(1) == R).all()
    for s in range(0, 4096, 256):
        ind = oracle.equity_mc_indep(int(states[s, 0]), int(states[s, 1]), 5, R, seed=s + 1)
        for k in range(3):
            p = (got[s, k] + ind[k]) / (2 * R)
            assert abs(got[s, k] - ind[k]) / R <= 3 * (2 * p * (1 - p) / R) ** 0.5 + 2e-4


def test_equity_mc_per_state_opponent_override(phe, oracle, torch):
    st = phe.pack_states([["As", "Ah"], ["As", "Ah"], ["As", "Ah"]], [[], [], []], n_opp=[0, 2, 7])
    got = phe.equity_mc(torch.from_numpy(st).cuda(), 4, 4000, seed=3).cpu().numpy()
    hero = int(st[0, 1])
    for s, k in enumerate((4, 2, 7)):
        assert (got[s] == oracle.equity_mc(hero, hero, k, 4000, 3, state_idx=s)).all()


# ---- exact plans (simulate_processes) ------------------------------------------------------------------------
@pytest.mark.parametrize("rnd", [0, 1, 2, 3])
def test_equity_plan_matches_oracle(phe, oracle, torch, rnd):
    rng = np.random.default_rng(20 + rnd)
    n_exist = {0: 2, 1: 5, 2: 6, 3: 7}[rnd]
    T = 6
    exist = np.argsort(rng.random((T, 52)), axis=1)[:, :n_exist].astype(np.uint8)
    for plans in (oracle.DEFAULT_PLANS, oracle.DEBUG_PLANS):
        plan, n = plans[rnd]
        got = phe.equity_plan(exist, plan, seed=9, stream_base=100)
        got_dev = phe.equity_plan(torch.from_numpy(exist).cuda(), plan, seed=9, stream_base=100).cpu().numpy()
        for t in range(T):
            w, tt, l, nn = oracle.equity_plan(exist[t], plan, seed=9, stream=100 + t)
            assert tuple(got[t]) == (w, tt, l) == tuple(got_dev[t]) and nn == n == got[t].sum()


def test_equity_plan_k9(phe, oracle, torch):
    for text, rnd, want in [("As Ah 2c 7d Tc 3h 9s", 3, (852, 1, 137)), ("7h 8h 9h Th 2c 2d Ks", 3, (0, 201, 789)),
                            ("As Ah 2c 7d Tc 3h", 2, (39837, 44, 5659))]:
        got = phe.equity_plan(np.array([oracle.ids(text)], dtype=np.uint8), oracle.DEFAULT_PLANS[rnd][0])
        assert tuple(got[0]) == want


def test_equity_plan_large_batch_uses_smem_kernel(phe, oracle, torch):
    rng = np.random.default_rng(5)
    exist = np.argsort(rng.random((64, 52)), axis=1)[:, :6].astype(np.uint8)   # 64 turn tasks x 45,540 deals
    got = phe.equity_plan(exist, oracle.DEFAULT_PLANS[2][0])
    for t in (0, 31, 63):
        assert tuple(got[t]) == oracle.equity_plan(exist[t], oracle.DEFAULT_PLANS[2][0])[:3]


def test_equity_plan_rejects_bad_plans(phe, torch):
    with pytest.raises(ValueError):
        phe.equity_plan(np.array([[51, 50]], dtype=np.uint8), [(1, 0, True)])
    with pytest.raises(ValueError):
        phe.equity_plan(np.array([[51, 50, 0, 21, 32, 10, 9]], dtype=np.uint8), [(2, 5000, True)])


# ---- showdown / winner sets ------------------------------------------------------------------------------------
@pytest.mark.parametrize("G,P", [(1, 2), (1000, 2), (1000, 6), (40000, 6), (3, 9)])
def test_showdown_matches_oracle(phe, oracle, torch, G, P):
    rng = np.random.default_rng(G + P)
    deals = np.argsort(rng.random((G, 52)), axis=1)[:, : 5 + 2 * P].astype(np.uint8)
    boards = phe.ids_to_masks(deals[:, :5])
    holes = phe.ids_to_masks(deals[:, 5:].reshape(G, P, 2))
    live = rng.integers(1, 1 << P, size=G).astype(np.uint32)
    win, ranks = phe.showdown(boards, holes, live, return_ranks=True)
    win_d, ranks_d = phe.showdown(torch.from_numpy(boards).cuda(), torch.from_numpy(holes).cuda(),
                                  torch.from_numpy(live.astype(np.int32)).cuda(), return_ranks=True)
    assert (win_d.cpu().numpy() == win).all() and (ranks_d.cpu().numpy().astype(np.uint16) == ranks).all()
    step = max(1, G // 500)
    for g in range(0, G, step):
        w, r = oracle.winner_mask(deals[g, :5], deals[g, 5:], int(live[g]), with_ranks=True)
        assert int(win[g]) == w and (ranks[g] == r).all()


def test_get_winner_set_v2_surface(phe, oracle, torch):
    D = phe.DECK
    ids = oracle.ids
    flop, turn, river = [D[i] for i in ids("2c 7d Tc")], [D[ids("3h")[0]]], [D[ids("9s")[0]]]
    players = {"a": [D[i] for i in ids("As Ah")], "b": [D[i] for i in ids("Kd Kc")], 7: [D[i] for i in ids("Ad Ac")]}
    assert phe.GameEnv.get_winner_set_v2(flop, turn, river, players) == {"a", 7}
    assert phe.GameEnv.get_winner_set_v2(flop, turn, river, {"b": players["b"]}) == {"b"}
    res = {k: phe.GamePlayerResult.LOSE for k in players}
    phe.GameEnv.generate_game_result({"a", 7}, res)
    assert res["a"] == phe.GamePlayerResult.EVEN and res["b"] == phe.GamePlayerResult.LOSE


# ---- determinisation -----------------------------------------------------------------------------------------------
def test_deal_matches_oracle_and_is_valid(phe, oracle, torch):
    rng = np.random.default_rng(8)
    known_ids = np.argsort(rng.random((3000, 52)), axis=1)[:, :7]
    known = phe.ids_to_masks(known_ids)
    for need in (1, 2, 9, 21):
        got = phe.deal(torch.from_numpy(known).cuda(), need, seed=77, sample_offset=5, stream_base=10).cpu().numpy()
        for i in range(0, 3000, 100):
            assert (got[i] == oracle.deal(int(known[i]), 77, 5, 10 + i, need)).all()
        for i in range(3000):
            row = got[i].tolist()
            assert len(set(row)) == need and not (set(row) & set(known_ids[i].tolist())) and max(row) < 52


# ---- the equity service with the reference's tuples ---------------------------------------------------------------
def test_simulate_processes_service(phe, oracle, torch):
    D = phe.DECK
    ids = oracle.ids
    plans = {k: (v[0], v[1]) for k, v in oracle.DEFAULT_PLANS.items()}
    tasks = [
        ("g1", "p0", 3, [D[i] for i in ids("2c 7d Tc")], [D[ids("3h")[0]]], [D[ids("9s")[0]]], [D[i] for i in ids("As Ah")]),
        ("g1", "p1", 2, [D[i] for i in ids("2c 7d Tc")], [D[ids("3h")[0]]], None, [D[i] for i in ids("As Ah")]),
        ("g2", "p0", 1, [D[i] for i in ids("2c 7d Tc")], None, None, [D[i] for i in ids("Kd Kc")]),
        ("g2", "p1", 0, None, None, None, [D[i] for i in ids("7h 2d")]),
        ("g3", "p1", 3, [D[i] for i in ids("9h Th 2c")], [D[ids("2d")[0]]], [D[ids("Ks")[0]]], [D[i] for i in ids("7h 8h")]),
    ]
    q_in, q_out = queue.Queue(), queue.Queue()
    for t in tasks:
        q_in.put(t)
    q_in.put(None)
    served = phe.simulate_processes(q_in, q_out, plans, 0, "INFO")
    assert served == len(tasks)
    results = {}
    while not q_out.empty():
        r = q_out.get()
        results[(r[0], r[1], r[2])] = r
    assert results[("g1", "p0", 3)][3:] == (990, 715.0)                 # K9
    assert results[("g1", "p1", 2)][3:] == (45540, 34178.0)             # K9
    assert results[("g3", "p1", 3)][3:] == (990, -789.0)                # K9
    assert results[("g2", "p0", 1)][3] == 97290 and results[("g2", "p1", 0)][3] == 98000
    r = results[("g2", "p1", 0)]
    assert abs(phe.winning_probability(r[4], r[3]) - 0.35) < 0.03          # 7-2 offsuit heads-up is about 35 %


def test_enum7_interleaved_parts_add_up(phe, oracle, torch):
    full9, full = phe.enumerate_c52_7(device_out=True)
    for nparts in (2, 3, 8):
        acc9, acc = torch.zeros_like(full9), torch.zeros_like(full)
        for part in range(nparts):
            a9, a = phe.enumerate_c52_7(device_out=True, part=part, nparts=nparts)
            acc9 += a9; acc += a
        assert bool((acc9 == full9).all()) and bool((acc == full).all())
    a9, a = phe.enumerate_c52_7(1000, 5_000_000, device_out=True, part=1, nparts=2)
    b9, b = phe.enumerate_c52_7(1000, 5_000_000, device_out=True, part=0, nparts=2)
    o9, oh = oracle.enum7(1000, 5_000_000, nthreads=2)
    assert ((a + b).cpu().numpy() == oh).all() and ((a9 + b9).cpu().numpy() == o9).all()


def test_enum7_fused_allreduce_single_rank(phe, oracle, torch):
    """phe_enum7_allreduce with world = 1: the last-CTA epilogue (slot exchange, flags, epoch) on one GPU; replayed
    from a CUDA graph to check that the launch counter kept in device memory advances correctly."""
    gold = json.load(open(os.path.join(GOLDEN, "hist7463.json")))
    fe = phe.FusedEnum7()
    for _ in range(3):
        h9, h = fe.run()
        torch.cuda.synchronize()
        assert h9.tolist() == gold["hist9"] and h.tolist() == gold["hist7463"]
    h9, h = fe.run(5, 250001)
    torch.cuda.synchronize()
    o9, oh = oracle.enum7(5, 250001, nthreads=2)
    assert (h.cpu().numpy() == oh).all() and (h9.cpu().numpy() == o9).all()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fe.run()
    for _ in range(4):
        g.replay()
    torch.cuda.synchronize()
    assert fe.out[:9].tolist() == gold["hist9"] and fe.out[9:].tolist() == gold["hist7463"]
    assert fe.check() == 8          # launches completed (3 + 1 eager, 4 replays; the capture itself launches nothing); no exchange timed out


@pytest.mark.parametrize("nvls", ["0", "1"])
def test_enum7_fused_allreduce_two_ranks(torch, nvls):
    """The fused exchange between real peers (peer stores and NVLS multimem): tools/fused_enum_bench.py asserts the golden
    histograms on every rank for eager launches, a sub-range and graph replays.  Needs two GPUs on the box."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PHE_FUSED_NVLS=nvls)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29870 + int(nvls)), os.path.join(root, "tools", "fused_enum_bench.py")],
                       capture_output=True, text=True, timeout=300, env=env, cwd=root)
    if nvls == "1" and "no multicast mapping" in (r.stdout + r.stderr):
        pytest.skip("no NVLS multicast mapping on this system")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ("NVLS multimem" if nvls == "1" else "peer stores") in r.stdout


def test_concurrent_callers(phe, oracle, torch):
    """Many Python threads in one process (the reference runs its game loops as threads, env/workflow.py:848-878):
    ctypes drops the GIL around every call, the library keeps per-thread scratch and no mutable global state."""
    import threading
    ids = rand_hands(50000, 9)
    want = oracle.rank7_batch(ids)
    masks = phe.ids_to_masks(ids)
    st = phe.pack_states([["As", "Ah"]], [["2c", "7d", "Tc"]])
    want_mc = oracle.equity_mc(int(st[0, 0]), int(st[0, 1]), 1, 3000, 5)
    errors = []

    def worker(k):
        try:
            for _ in range(5):
                assert (phe.eval7(masks) == want).all()                                   # host entry point, per-thread scratch
                assert phe.evaluate_cards(*ids[k].tolist()) == int(want[k])
                assert (phe.equity_mc(st, 1, 3000, seed=5)[0] == want_mc).all()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):                                                   # device entry point on a private stream
                    got = phe.eval7(torch.from_numpy(masks).cuda())
                s.synchronize()
                assert (got.cpu().numpy().astype(np.uint16) == want).all()
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_eval7_suit_permutation_invariance(phe, torch):
    """Property at scale: relabelling the four suits (a permutation of the 16-bit lanes) never changes a rank."""
    ids = rand_hands(1 << 21, 12)
    base = phe.eval7(torch.from_numpy(ids).cuda())
    for perm in ([1, 2, 3, 0], [3, 2, 1, 0], [2, 0, 3, 1]):
        p = np.array(perm, dtype=np.uint8)
        relabelled = ((ids >> 2) << 2) | p[ids & 3]
        assert bool((phe.eval7(torch.from_numpy(relabelled.astype(np.uint8)).cuda()) == base).all())


def test_eval7_and_enum7_agree_on_a_colex_slice(phe, oracle, torch):
    """Two independent kernels and table sets: the generic evaluator on explicitly generated hands of a colex slice
    must reproduce the enumeration kernel's histogram of that slice."""
    b, n = 60_000_000, 400_000
    c = oracle.colex_unrank7(b).astype(np.int64)
    hands = np.empty((n, 7), dtype=np.uint8)
    for i in range(n):
        hands[i] = c
        j = 0
        while j < 6 and c[j] + 1 == c[j + 1]:
            c[j] = j
            j += 1
        c[j] += 1
    ranks = phe.eval7(torch.from_numpy(hands).cuda()).cpu().numpy().astype(np.int64)
    h9, h = phe.enumerate_c52_7(b, b + n)
    assert (np.bincount(ranks, minlength=7463) == h).all()
    assert int(h9.sum()) == n


def test_deal_on_garbage_input_terminates(phe, torch):
    """A `known` set that leaves fewer free cards than requested cannot be dealt; the kernel must give up, not spin."""
    full = int(0x1FFF1FFF1FFF1FFF)
    known = torch.tensor([full & ~0b111, full], dtype=torch.int64).cuda()        # 3 free cards / no free card
    out = phe.deal(known, 9, seed=1)
    torch.cuda.synchronize()
    assert out.shape == (2, 9)


def test_eval7_inside_a_cuda_graph(phe, oracle, torch):
    """The batched evaluator is launched with programmatic stream serialisation (PDL); it must also capture into and
    replay from a CUDA graph, after a copy node, after another kernel node and after itself."""
    ids = rand_hands(1 << 17, 21)
    want = oracle.rank7_batch(ids)
    src = torch.from_numpy(phe.ids_to_masks(ids)).cuda()
    masks = torch.empty_like(src)
    out = torch.empty(src.shape[0], dtype=torch.int16, device="cuda")
    out2 = torch.empty_like(out)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        masks.copy_(src)
        phe.eval7(masks, out=out)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        masks.copy_(src)                 # copy node -> evaluator
        phe.eval7(masks, out=out)
        masks.add_(0)                    # kernel node -> evaluator -> evaluator
        phe.eval7(masks, out=out2)
        phe.eval7(masks, out=out)
    out.zero_(); out2.zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert (out.cpu().numpy().astype(np.uint16) == want).all() and (out2.cpu().numpy().astype(np.uint16) == want).all()


@pytest.mark.gpu
def test_host_calls_do_not_wait_for_other_streams(phe, oracle, torch):
    """SURVEY 8b "Threading": the *_host entry points run on a non-blocking stream owned by the calling thread and synchronise
    only that stream.  A long kernel on torch's default (legacy NULL) stream -- what another game-loop thread or the predictor
    may have in flight -- must not delay them: on the NULL stream they would queue behind it."""
    import time
    ids = rand_hands(64, 21)
    want = oracle.rank7_batch(ids)
    st = phe.pack_states([["As", "Ah"]], [["2c", "7d", "Tc"]])
    board_and_hands = (["2c", "7d", "Tc"], ["3h"], ["9s"], {"a": ["As", "Ah"], "b": ["Kd", "Kc"]})
    phe.evaluate_cards(*ids[0].tolist())            # first launches load their kernels lazily, which waits for running work: warm up
    phe.eval7(ids)
    phe.equity_mc(st, 1, 2000, seed=5)
    phe.GameEnv.get_winner_set_v2(*board_and_hands)
    torch.cuda.synchronize()
    done = torch.cuda.Event()
    torch.cuda._sleep(int(0.4 * 1.9e9))                 # ~0.4 s spin kernel on the default stream (one thread: leaves the SMs free)
    done.record()
    t0 = time.perf_counter()
    for k in range(64):
        assert phe.evaluate_cards(*ids[k].tolist()) == int(want[k])
    assert (phe.eval7(ids) == want).all()
    got = phe.equity_mc(st, 1, 2000, seed=5)
    winners = phe.GameEnv.get_winner_set_v2(*board_and_hands)
    dt = time.perf_counter() - t0
    still_running = not done.query()
    torch.cuda.synchronize()
    assert winners == {"a"} and int(got.sum()) == 2000
    assert still_running and dt < 0.2, f"host calls took {dt * 1e3:.1f} ms behind a 400 ms kernel on another stream (still running: {still_running})"
