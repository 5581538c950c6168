"""Replay record text format against the reference's own writer (AlphaGoZero.store_transition) and reader
(tools/data_loader.py), tests/golden/records_run.json."""
import json
import os

import numpy as np

from conftest import GOLDEN

import texasholdempocker_rl_b200 as phe


def test_format_and_parse_match_reference():
    g = json.load(open(os.path.join(GOLDEN, "records_run.json")))
    ins = g["inputs"]
    text = phe.records.format_records(np.asarray([r["obs"] for r in ins]), np.asarray([r["probs"] for r in ins]), np.asarray([r["masks"] for r in ins]),
                                      [r["reward"] for r in ins], [r["card"] for r in ins], [r["pbin"] for r in ins], [r["obin"] for r in ins])
    assert text == g["text"]                                        # byte-identical to the reference's file
    obs, probs, masks, rew, card, pb, ob = phe.records.parse_records(g["text"])
    want = g["parsed"]
    assert obs.dtype == np.int32 and obs.tolist() == [r["obs"] for r in want]
    assert probs.tolist() == [r["probs"] for r in want] and masks.tolist() == [r["masks"] for r in want]
    assert rew.tolist() == [r["reward"] for r in want] and card.tolist() == [r["card"] for r in want]
    assert pb.tolist() == [r["pbin"] for r in want] and ob.tolist() == [r["obin"] for r in want]


def test_truncated_and_empty():
    g = json.load(open(os.path.join(GOLDEN, "records_run.json")))
    cut = g["text"][: len(g["text"]) // 2]
    obs = phe.records.parse_records(cut)[0]
    assert 0 < obs.shape[0] < len(g["parsed"]) and obs.tolist() == [r["obs"] for r in g["parsed"][: obs.shape[0]]]
    for frac in (0.11, 0.37, 0.5, 0.83, 0.999):                      # cuts inside any of the eight lines
        part = phe.records.parse_records(g["text"][: int(len(g["text"]) * frac)])[0]
        assert part.tolist() == [r["obs"] for r in g["parsed"][: part.shape[0]]]
    assert phe.records.parse_records("")[0].shape[0] == 0
    assert phe.records.format_records(np.zeros((0, 93)), np.zeros((0, 12)), np.zeros((0, 12)), [], [], [], []) == ""
