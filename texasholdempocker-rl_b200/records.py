"""Replay record text format: what ``AlphaGoZero.store_transition`` appends to the training file
(rl/AlphaGoZero.py:90-103) and ``tools/data_loader.py:6-39`` reads back -- eight lines per decision point:

    observation           comma separated ints
    action probabilities  comma separated '%.5f'
    action mask           comma separated '%d'
    reward value          '%.8f'
    card result value     '%f'
    player winning-probability bin      '%d'
    opponent winning-probability bin    '%d'
    (blank line)

``format_records`` writes a whole batch (arrays as they come back from the device) at once; ``parse_records`` is the
batched reader.  Host-side text handling only: no device work, nothing to accelerate."""
import io

import numpy as np


def _host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def format_records(observations, action_probs, action_masks, reward_values, card_result_values, player_bins, opponent_bins):
    """Arrays with leading dimension B -> the text store_transition would have written for the B transitions in order."""
    obs, probs, masks = _host(observations), _host(action_probs), _host(action_masks)
    rew, card, pb, ob = (_host(x).reshape(-1) for x in (reward_values, card_result_values, player_bins, opponent_bins))
    n = obs.shape[0]
    if not (probs.shape[0] == masks.shape[0] == rew.shape[0] == card.shape[0] == pb.shape[0] == ob.shape[0] == n):
        raise ValueError("all record arrays need the same leading dimension")
    out = io.StringIO()
    for i in range(n):
        out.write(",".join(str(int(v)) for v in obs[i]))
        out.write("\n")
        out.write(",".join("%.5f" % v for v in probs[i]))
        out.write("\n")
        out.write(",".join("%d" % v for v in masks[i]))
        out.write("\n%.8f\n%f\n%d\n%d\n\n" % (rew[i], card[i], pb[i], ob[i]))
    return out.getvalue()


def parse_records(text):
    """Inverse of format_records / batched form of tools/data_loader.py:train_data_generator (one pass, no epochs).
    Returns (obs int32[B,F], probs float64[B,A], masks int64[B,A], reward float64[B], card float64[B], pbin int64[B], obin int64[B]).
    A truncated trailing record is dropped, as the reference's reader stops at the first empty line group."""
    lines = text.split("\n")
    recs = []
    for k in range(0, len(lines), 8):
        if k + 7 >= len(lines):          # the seventh line is only complete if a newline (hence another element) follows it
            break
        grp = lines[k: k + 7]
        if any(g == "" for g in grp):
            break
        recs.append(grp)
    if not recs:
        e = np.empty
        return e((0, 0), np.int32), e((0, 0)), e((0, 0), np.int64), e(0), e(0), e(0, np.int64), e(0, np.int64)
    obs = np.asarray([[int(v) for v in g[0].rstrip().split(",")] for g in recs], dtype=np.int32)
    probs = np.asarray([[float(v) for v in g[1].split(",")] for g in recs], dtype=np.float64)
    masks = np.asarray([[int(v) for v in g[2].split(",")] for g in recs], dtype=np.int64)
    return (obs, probs, masks, np.asarray([float(g[3]) for g in recs]), np.asarray([float(g[4]) for g in recs]),
            np.asarray([int(g[5]) for g in recs], dtype=np.int64), np.asarray([int(g[6]) for g in recs], dtype=np.int64))
