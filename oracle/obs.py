"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's observation encoder for the parity tests.

Follows Env.get_obs (env/env.py:416-465), get_all_player_current_status (:479-581), _map_obs_idx_to_model_index
(:383-408), get_card_representations (:702-723) and CutByThreshold.cut (tools/biner.py:8-27), on the packed batch
arrays of texasholdempocker-rl_b200/observation.py.  Pinned by tests/golden/obs_run.json.gz (outputs of the reference's
own Env.get_obs).  Never imported by the product path."""
import numpy as np


def cut(thr, num, den, marker=None):
    """tools/biner.py:8-27; marker is given only for the cutters built with include_invalid_bin."""
    n = len(thr)
    if marker is not None and marker < 0:
        return n + 1
    if den == 0:
        return n
    v = int(num) / int(den)
    for i, t in enumerate(thr):
        if v < t:
            return i
    return n


def encode(batch, thresholds):
    """batch: dict of numpy arrays (round_role, cards, values, bins, others, history); thresholds: the 16 lists."""
    T = [list(map(float, t)) for t in thresholds]
    n = batch["values"].shape[0]
    out = []
    for i in range(n):
        o = [int(batch["round_role"][i, 0]), int(batch["round_role"][i, 1])]
        fig, dec = [], []
        for c in batch["cards"][i]:
            c = int(c)
            if c < 52:
                fig.append(c >> 2), dec.append(c & 3)            # env/constants.py:14-36 via card id = figure*4 + decor
            elif c == 53:
                fig.append(13), dec.append(4)                    # env/env.py:717-719
            else:
                fig.append(14), dec.append(5)                    # env/env.py:709
        o += fig + dec
        game_init, pot, a_init, a_left, a_bet = (int(v) for v in batch["values"][i])
        assets = [game_init, a_init, a_left]
        bins = [int(v) for v in batch["bins"][i]]
        o += [cut(T[0], a_init, game_init), cut(T[1], a_left, game_init), cut(T[2], a_left, a_init), cut(T[3], a_bet, pot)]
        o += [cut(T[4 + a], a_bet, assets[a]) for a in range(3)]
        o += [cut(T[7 + a], pot, assets[a]) for a in range(3)]
        o += [cut(T[10], v, pot, v) for v in bins]
        for a in range(3):
            o += [cut(T[11 + a], v, assets[a], v) for v in bins]
        oth = batch["others"][i]
        for f in range(3):                                       # env/env.py:399-401: field-major over the other players
            for j in range(oth.shape[0]):
                status, left, init = (int(v) for v in oth[j])
                o.append(status if f == 0 else (cut(T[14], left, init) if f == 1 else cut(T[15], init, a_init)))
        o += [int(v) for v in batch["history"][i]]
        out.append(o)
    return np.asarray(out, dtype=np.int32)
