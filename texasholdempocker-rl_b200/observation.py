"""Batched observation encoder on the device.

Replaces ``Env.get_obs`` and its helpers -- ``get_all_historical_action``, ``get_all_player_current_status``,
``_map_obs_idx_to_model_index``, ``get_card_representations`` (env/env.py:383-408, 416-581, 702-723) with the cutters
of ``Env.__init__`` (env/env.py:52-81, tools/biner.py) -- by one kernel over a struct-of-arrays batch of decision
points (``phe_encode_obs``, csrc/phe_obs.cu).  Its output is exactly what ``PolicyValueNet`` / ``BatchedPredictor``
consume, so leaves of many simulations go infoset -> int32[B, 93] -> policy / value without leaving the device.

``ObsEncoder.get_obs(infoset, mcts_acting_player_name=None)`` is the drop-in for ``Env.get_obs``; ``pack`` turns
reference ``InfoSet`` objects (env/game.py:1053-1092) into the batch arrays; ``encode`` runs the kernel on numpy
arrays (host in, host out) or CUDA tensors (in place on the device).  There is no CPU path.

Heads-up only, like the reference (``MAX_PLAYER_NUMBER = 2``, env/constants.py:4; env/env.py:475 "todo").
"""
import ctypes as C

import numpy as np

from . import _native as N
from .cardset import card_id

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

CARD_NONE, CARD_HIDDEN = 52, 53          # PHE_OBS_CARD_NONE / PHE_OBS_CARD_HIDDEN
N_CUTTERS, MAX_THR = 16, 20              # PHE_OBS_CUTTERS / PHE_OBS_MAX_THR
MAX_PLAYER_NUMBER = 2                    # env/constants.py:4

# env/constants.py:93-111 -- the threshold lists, computed the way the reference computes them so that every float
# is bit-identical (np.arange(0.1, 1, 0.1) is not 0.1 * i)
_RAISE_RANGE_STARTS = [0, 0, 0., 0.0125, 0.025, 0.05, 0.1, 0.2, 0.4, 0.6, 0.8, 1.]       # ACTION_BINS_DICT range starts
CUTTER_BINS_LIST = _RAISE_RANGE_STARTS[3:-1]
CUTTER_SELF_BINS_LIST = CUTTER_BINS_LIST + [1.] + [1 / i for i in CUTTER_BINS_LIST[::-1]]
CUTTER_DEFAULT_LIST = np.arange(0.1, 1, 0.1).tolist()
CUTTER_SELF_DEFAULT_LIST = CUTTER_DEFAULT_LIST + [1.] + [1 / i for i in CUTTER_DEFAULT_LIST[::-1]]
N_ACTION_FEATURE_BINS = 10               # utils/workflow_utils.py:179-184


def cutter_thresholds(max_players=MAX_PLAYER_NUMBER):
    """The 16 threshold arrays in the order of include/phe_b200.h (env/env.py:52-81)."""
    d, b = np.array(CUTTER_DEFAULT_LIST), np.array(CUTTER_BINS_LIST)
    sb, sd = np.array(CUTTER_SELF_BINS_LIST), np.array(CUTTER_SELF_DEFAULT_LIST)
    return [d * max_players, d * max_players, d, d,                       # 0-3
            b * max_players, b, sb,                                       # 4-6  player_history_bet_to_assets
            b * max_players, b * max_players, sb,                         # 7-9  pot_to_assets
            sd,                                                           # 10   action_value_to_pot (+ invalid bin)
            b * max_players, b, b,                                        # 11-13 action_value_to_assets (+ invalid bin)
            d,                                                            # 14   player_value_left_to_init_value
            np.array(CUTTER_DEFAULT_LIST + [1.]) * max_players]           # 15   player_init_value_to_acting_player_init_value


def _cut_table_bytes(max_players):
    thr = np.zeros((N_CUTTERS, MAX_THR), dtype=np.float64)
    ln = np.zeros(N_CUTTERS, dtype=np.int32)
    for i, t in enumerate(cutter_thresholds(max_players)):
        thr[i, : len(t)] = t
        ln[i] = len(t)
    raw = thr.tobytes() + ln.tobytes()
    want = int(N.lib().phe_obs_cut_table_bytes())
    if len(raw) > want:
        raise N.PheError("cut table layout mismatch")
    return raw + bytes(want - len(raw))


def _player_index(name):
    return int(str(name).lstrip("player_"))          # GET_PLAYER_ID_BY_NAME, env/constants.py:118


def relative_position(player, base, num_players):
    """env/env.py:742-749."""
    r = _player_index(player) - _player_index(base)
    return r + num_players if r < 0 else r


class ObsEncoder:
    def __init__(self, num_action_bins=12, historical_action_per_round=3, num_players=MAX_PLAYER_NUMBER):
        if num_players != 2:
            raise ValueError("the observation layout is heads-up only, like the reference (env/constants.py:4)")
        self.num_action_bins = num_action_bins
        self.historical_action_per_round = historical_action_per_round
        self.num_players = num_players
        self.n_bins, self.n_other = N_ACTION_FEATURE_BINS, num_players - 1
        self.n_hist = 4 * historical_action_per_round * num_players
        self.n_fields = int(N.lib().phe_obs_num_fields(self.n_bins, self.n_other, self.n_hist))
        self._tab_bytes = None
        self._tab_dev = {}

    # -- packing ---------------------------------------------------------------------------------------------------
    def empty_batch(self, n):
        return {"round_role": np.zeros((n, 2), np.int32), "cards": np.full((n, 7), CARD_NONE, np.uint8),
                "values": np.zeros((n, 5), np.int64), "bins": np.full((n, self.n_bins), -1, np.int64),
                "others": np.zeros((n, self.n_other, 3), np.int64),
                "history": np.full((n, self.n_hist), self.num_action_bins, np.uint8)}

    def pack_into(self, batch, i, infoset, mcts_acting_player_name=None):
        """Writes decision point `infoset` (reference InfoSet or anything with its attributes) into row i."""
        me = infoset.player_name
        batch["round_role"][i] = (int(infoset.current_round), relative_position(me, infoset.button_player_name, infoset.num_players))
        hidden = mcts_acting_player_name is not None and me != mcts_acting_player_name           # env/env.py:445-448
        row = batch["cards"][i]
        row[:] = CARD_NONE
        for off, cards, cnt in ((0, infoset.player_hand_cards, 2), (2, infoset.flop_cards, 3), (5, infoset.turn_cards, 1), (6, infoset.river_cards, 1)):
            if cards is None:
                continue
            if len(cards) != cnt:
                raise ValueError("Length of list_cards does not equal to num_cards.")           # env/env.py:712
            for j, c in enumerate(cards):
                row[off + j] = CARD_HIDDEN if (hidden and off == 0) else card_id(c)
        st = infoset.player_status_value_left_bet_dict
        _, left, bet = st[me]
        a_init = infoset.player_value_init_dict[me]
        batch["values"][i] = (infoset.game_init_value, infoset.pot_value, a_init, left, bet)
        bins = list(infoset.bin_bet_value_list)
        if len(bins) != self.n_bins:
            raise ValueError(f"bin_bet_value_list has {len(bins)} entries, the model layout has {self.n_bins}")
        batch["bins"][i] = bins
        for name, init in infoset.player_value_init_dict.items():
            if name == me:
                continue
            status, o_left, _ = st[name]
            slot = (relative_position(name, infoset.button_player_name, infoset.num_players) - 1) % self.n_other
            batch["others"][i, slot] = (getattr(status, "value", status), o_left, init)
        hist = batch["history"][i]
        hist[:] = self.num_action_bins
        H = self.historical_action_per_round
        names = [f"player_{k}" for k in range(self.num_players)]
        opp = [n for n in names if n != me][0]                                                    # env/env.py:475-476
        for base, name in ((0, me), (4 * H, opp)):
            for r, round_bins in enumerate(infoset.all_player_round_action_bins_dict.get(name, ())):
                for j, b in enumerate(round_bins[:H]):
                    hist[base + r * H + j] = b
        return batch

    def pack(self, infosets, mcts_acting_player_names=None):
        batch = self.empty_batch(len(infosets))
        for i, info in enumerate(infosets):
            self.pack_into(batch, i, info, None if mcts_acting_player_names is None else mcts_acting_player_names[i])
        return batch

    # -- encoding ----------------------------------------------------------------------------------------------------
    def _table(self, device):
        t = self._tab_dev.get(device)
        if t is None:
            if self._tab_bytes is None:
                self._tab_bytes = _cut_table_bytes(self.num_players)
            t = torch.frombuffer(bytearray(self._tab_bytes), dtype=torch.uint8).to(device)
            self._tab_dev[device] = t
        return t

    def encode(self, batch):
        """batch of numpy arrays -> numpy int32[B, F]; batch of CUDA tensors -> CUDA tensor (current stream)."""
        if torch is None or not torch.cuda.is_available():
            raise N.PheError("the observation encoder needs a CUDA device (there is no CPU path)")
        on_device = isinstance(batch["values"], torch.Tensor) and batch["values"].is_cuda
        dev = batch["values"].device if on_device else torch.device("cuda", torch.cuda.current_device())
        want = {"round_role": torch.int32, "cards": torch.uint8, "values": torch.int64, "bins": torch.int64, "others": torch.int64,
                "history": torch.uint8}
        t = {}
        for k, dt in want.items():
            v = batch[k]
            v = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
            if v.dtype != dt:
                raise ValueError(f"{k} must be {dt}, got {v.dtype}")
            t[k] = v.to(dev).contiguous()
        n = t["values"].shape[0]
        shapes = {"round_role": (n, 2), "cards": (n, 7), "values": (n, 5), "bins": (n, self.n_bins), "others": (n, self.n_other, 3),
                  "history": (n, self.n_hist)}
        for k, s in shapes.items():
            if tuple(t[k].shape) != s:
                raise ValueError(f"{k} must have shape {s}, got {tuple(t[k].shape)}")
        out = torch.empty((n, self.n_fields), dtype=torch.int32, device=dev)
        p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
        with torch.cuda.device(dev):
            N.check(N.lib().phe_init(dev.index if dev.index is not None else torch.cuda.current_device()))
            N.check(N.lib().phe_encode_obs(p(t["round_role"]), p(t["cards"]), p(t["values"]), p(t["bins"]), p(t["others"]), p(t["history"]),
                                           p(self._table(dev)), n, self.n_bins, self.n_other, self.n_hist, p(out),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out if on_device else out.cpu().numpy()

    def get_obs(self, infoset, mcts_acting_player_name=None):
        """Drop-in for ``Env.get_obs`` (env/env.py:416-465): np.int32[F] for one decision point."""
        return self.encode(self.pack([infoset], [mcts_acting_player_name]))[0]
