"""CPU: the restatement of the temporal association (oracle/assoc.py) against the reference's own functions --
their recorded outputs (tests/golden/assoc_golden.json) everywhere, the functions themselves where /root/reference is mounted."""
import importlib.util
import json
import os

import pytest

from i2vsgg_b200 import synth
from oracle import assoc

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    spec = importlib.util.spec_from_file_location("make_assoc_golden", os.path.join(HERE, "golden", "make_assoc_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CASES


@pytest.mark.parametrize("name", ["plain", "gaps", "long"])
def test_association_oracle_equals_reference_output(name):
    golden = json.load(open(os.path.join(HERE, "golden", "assoc_golden.json")))[name]
    rec, cnt = synth.clip_records(**cases()[name])
    got = assoc.association({"vid": synth.records_to_frame_relations(rec, cnt)}).get("vid", [])
    assert len(got) == len(golden) and len(got) > 0
    for a, b in zip(got, golden):
        assert a == b           # triplet, score (bit-equal float64), duration, both trajectories, rel_idex


@pytest.mark.needs_reference
def test_association_oracle_equals_reference_live():
    ref = assoc.reference_functions()
    rec, cnt = synth.clip_records(seed=9, frames=50, tracks=15, clutter=12, empty=(3, 4, 30))
    fr = synth.records_to_frame_relations(rec, cnt, frame_numbers=range(7, 57))
    want = ref["association"]({"v": [[f, list(p)] for f, p in fr]})["v"]
    got = assoc.association({"v": fr})["v"]
    assert got == want and len(got) > 0
    assert assoc._iou([0, 0, 10, 10], [5, 5, 15, 15]) == ref["_iou"]([0, 0, 10, 10], [5, 5, 15, 15])


@pytest.mark.needs_reference
def test_association_oracle_equals_reference_live_on_long_relations():
    """Relations of several hundred frames (tracks that never miss a frame): their mean confidence, a sort key, is numpy's
    pairwise summation above 128 terms -- the case the device kernel once got wrong by overrunning its stack."""
    ref = assoc.reference_functions()
    rec, cnt = synth.clip_records(seed=21, frames=640, tracks=14, clutter=6, dropout=0.0)
    fr = synth.records_to_frame_relations(rec, cnt)
    want = ref["association"]({"v": [[f, list(p)] for f, p in fr]})["v"]
    got = assoc.association({"v": fr})["v"]
    assert got == want and max(r["duration"][1] - r["duration"][0] for r in got) > 256


def test_conv_layer_takes_parity_planes_only_where_the_kernel_applies():
    """`Conv2d.takes_split`: stride 2, odd kernel with "same" padding, even map, output positions dividing a 128-row tile."""
    from i2vsgg_b200.model.faster_rcnn.utils import Conv2d
    assert Conv2d(96, 128, 5, same_padding=True, stride=2).takes_split(16, 16, 96)       # conv_lo's second layer
    assert not Conv2d(96, 128, 5, same_padding=True, stride=1).takes_split(16, 16, 96)   # stride 1
    assert not Conv2d(96, 128, 5, same_padding=False, stride=2).takes_split(16, 16, 96)  # no padding
    assert not Conv2d(96, 128, 5, same_padding=True, stride=2).takes_split(15, 16, 96)   # odd map
    assert not Conv2d(96, 128, 5, same_padding=True, stride=2).takes_split(16, 16, 64)   # channel count of another layer
    assert not Conv2d(96, 256, 5, same_padding=True, stride=2).takes_split(16, 16, 96)   # more than one tile column
    assert not Conv2d(96, 128, 5, same_padding=True, stride=2).takes_split(20, 20, 96)   # 100 positions do not divide 128


def test_fill_empty_frames_rules():
    mk = lambda pattern: [[i, [1] if c == "x" else []] for i, c in enumerate(pattern)]
    assert assoc.fill_empty_frames(mk("x.x")) == [-1, 0, -1]                 # tie: the earlier neighbour
    assert assoc.fill_empty_frames(mk("x..x")) == [-1, 0, 3, -1]
    assert assoc.fill_empty_frames(mk(".x")) == [1, -1]
    lonely = assoc.fill_empty_frames(mk("x" + "." * 12 + "x"))
    assert lonely[6] == -2 and lonely[1] == 0 and lonely[12] == 13             # +-4 window all empty -> stays empty
