"""Observation encoder against observations produced by the reference's own Env.get_obs while it
played random heads-up hands (tests/golden/obs_run.json.gz, made by tests/golden/make_obs_golden.py).  Bit-exact."""
import gzip
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import GOLDEN

import texasholdempocker_rl_b200 as phe
from oracle import obs as obs_oracle

ob = phe.observation


@pytest.fixture(scope="module")
def records():
    return json.load(gzip.open(os.path.join(GOLDEN, "obs_run.json.gz")))


def infoset_of(r):
    """A stand-in with the attributes of the reference's InfoSet (env/game.py:1053-1092) built from a golden record."""
    names = [f"player_{k}" for k in range(r["num_players"])]
    return SimpleNamespace(
        player_name=names[r["player"]], button_player_name=names[r["button"]], num_players=r["num_players"], current_round=r["round"],
        player_hand_cards=r["hand"], flop_cards=r["flop"], turn_cards=r["turn"], river_cards=r["river"],
        game_init_value=r["game_init"], pot_value=r["pot"], bin_bet_value_list=r["bins"],
        player_value_init_dict={n: r["init"][k] for k, n in enumerate(names)},
        player_status_value_left_bet_dict={n: (r["status"][k], r["left"][k], r["bet"][k]) for k, n in enumerate(names)},
        all_player_round_action_bins_dict={n: r["history"][k] for k, n in enumerate(names) if r["history"][k]})


def pack_records(enc, recs):
    names = ["player_0", "player_1"]
    return enc.pack([infoset_of(r) for r in recs], [names[1 - r["player"]] if r["hidden"] else None for r in recs])


def test_thresholds_are_the_reference_floats():
    t = ob.cutter_thresholds()
    assert [len(x) for x in t] == [9, 9, 9, 9, 8, 8, 17, 8, 8, 17, 19, 8, 8, 8, 9, 10]
    assert t[2][2] == 0.30000000000000004 and t[2][6] == 0.7000000000000001           # np.arange(0.1, 1, 0.1), not 0.1 * i
    assert t[0][2] == 0.30000000000000004 * 2 and t[6][8] == 1.0 and t[6][-1] == 1 / 0.0125 and t[15][-1] == 2.0


def test_packing_and_oracle_match_reference_observations(records):
    enc = ob.ObsEncoder()
    assert enc.n_fields == 93
    batch = pack_records(enc, records)
    got = obs_oracle.encode(batch, ob.cutter_thresholds())
    want = np.asarray([r["obs"] for r in records], dtype=np.int32)
    assert got.shape == want.shape
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"first mismatch at record/field {bad[0].tolist()}: {got[tuple(bad[0])]} != {want[tuple(bad[0])]}"


def test_observations_index_inside_the_embedding_table(records):
    """Reference quirk: a cutter with n thresholds emits bins 0..n but its field has only n embedding rows
    (rl/models.py:39-67), so top bins alias the next field's first row.  The reference's observations do that
    routinely; what must hold is that every index stays inside the table (and PolicyValueNet accepts them)."""
    import torch
    starts, sizes, rows = phe.policy_value.field_layout(num_output_class=12, num_bins=10, historical_action_per_round=3)
    want = np.asarray([r["obs"] for r in records], dtype=np.int32)
    assert (want >= 0).all() and (want <= sizes[None, :]).all() and (want + starts[None, :] < rows).all()
    assert (want == sizes[None, :]).any()                      # the aliasing really occurs in reference play
    net = phe.policy_value.PolicyValueNet(**phe.policy_value.DEBUG_MODEL_PARAMS)
    net.check_observations(torch.from_numpy(want))


def test_pack_rejects_malformed(records):
    enc = ob.ObsEncoder()
    info = infoset_of(records[0])
    info.bin_bet_value_list = info.bin_bet_value_list[:-1]
    with pytest.raises(ValueError):
        enc.pack([info])
    info = infoset_of(records[0])
    info.player_hand_cards = info.player_hand_cards[:1]
    with pytest.raises(ValueError):
        enc.pack([info])
    with pytest.raises(ValueError):
        ob.ObsEncoder(num_players=3)


def test_encoder_needs_cuda(records):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    enc = ob.ObsEncoder()
    with pytest.raises(phe._native.PheError):
        enc.encode(pack_records(enc, records[:2]))


@pytest.mark.gpu
def test_kernel_matches_reference_observations_gpu(records):
    import torch
    enc = ob.ObsEncoder()
    batch = pack_records(enc, records)
    want = np.asarray([r["obs"] for r in records], dtype=np.int32)
    before = phe._native.launch_count()
    got = enc.encode(batch)
    assert phe._native.launch_count() == before + 1
    assert got.dtype == np.int32 and np.array_equal(got, want)
    # device-resident batch, drop-in single call, and the edge cases: empty batch
    dev = {k: torch.from_numpy(v).cuda() for k, v in batch.items()}
    dgot = enc.encode(dev)
    assert dgot.is_cuda and np.array_equal(dgot.cpu().numpy(), want)
    r = records[7]
    one = enc.get_obs(infoset_of(r), "player_%d" % (1 - r["player"]) if r["hidden"] else None)
    assert one.shape == (93,) and np.array_equal(one, want[7])
    assert enc.encode(enc.empty_batch(0)).shape == (0, 93)


@pytest.mark.gpu
def test_kernel_matches_oracle_on_random_chip_counts_gpu():
    """Property test beyond the golden games: random stacks / pots / bets including zeros (denominator 0 -> last bin),
    invalid bins (-1) and ratios that land exactly on thresholds; the kernel equals the CPU restatement bit for bit."""
    enc = ob.ObsEncoder()
    rng = np.random.default_rng(3)
    n = 20000
    b = enc.empty_batch(n)
    b["round_role"][:] = np.stack([rng.integers(0, 4, n), rng.integers(0, 2, n)], 1)
    b["cards"][:] = rng.integers(0, 54, (n, 7))
    unit = rng.choice([1, 25, 50, 1000], n)
    vals = rng.integers(0, 400, (n, 5)) * unit[:, None]
    vals[rng.random((n, 5)) < 0.05] = 0
    b["values"][:] = vals
    bins = rng.integers(0, 400, (n, 10)) * unit[:, None]
    bins[rng.random((n, 10)) < 0.2] = -1
    b["bins"][:] = bins
    oth = rng.integers(0, 400, (n, 1, 3)) * unit[:, None, None]
    oth[:, :, 0] = rng.integers(0, 4, (n, 1))
    b["others"][:] = oth
    b["history"][:] = rng.integers(0, 13, (n, 24))
    # exact-threshold ratios: bet / pot = k / 10
    k = rng.integers(0, 12, n)
    b["values"][: n // 4, 4] = k[: n // 4] * 7
    b["values"][: n // 4, 1] = 70
    got = enc.encode(b)
    want = obs_oracle.encode(b, ob.cutter_thresholds())
    assert np.array_equal(got, want)
