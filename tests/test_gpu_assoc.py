"""GPU parity test (B200 box): `sgg.association` (greedy association kernel + host assembly) against the oracle, which is
pinned to the reference's own functions by tests/test_oracle_assoc.py.  Everything must be equal, floats bit for bit."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from i2vsgg_b200 import synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    spec = importlib.util.spec_from_file_location("make_assoc_golden", os.path.join(HERE, "golden", "make_assoc_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CASES


@pytest.mark.parametrize("name", ["plain", "gaps", "long"])
def test_association_equals_oracle(name):
    from i2vsgg_b200 import sgg
    from oracle import assoc
    rec, cnt = synth.clip_records(**cases()[name])
    want = assoc.association({"vid": synth.records_to_frame_relations(rec, cnt)}).get("vid", [])
    got = sgg.association(torch.from_numpy(rec).cuda(), torch.from_numpy(cnt).cuda())
    assert len(got) == len(want) > 0
    for a, b in zip(got, want):
        assert a == b


def test_association_of_relations_hundreds_of_frames_long():
    """Tracks that never miss a frame: relations of 300-600 frames, whose mean confidence goes through the halving levels
    of numpy's pairwise summation (above 128 terms) -- the path that a clip of a real video takes and short tracks do not."""
    from i2vsgg_b200 import sgg
    from oracle import assoc
    rec, cnt = synth.clip_records(seed=21, frames=640, tracks=14, clutter=6, dropout=0.0)
    want = assoc.association({"vid": synth.records_to_frame_relations(rec, cnt)}).get("vid", [])
    got = sgg.association(torch.from_numpy(rec).cuda(), torch.from_numpy(cnt).cuda())
    assert len(got) == len(want) > 0
    assert max(r["duration"][1] - r["duration"][0] for r in got) > 256
    for a, b in zip(got, want):
        assert a == b


def test_association_frame_numbers_names_and_empty_clip():
    from i2vsgg_b200 import sgg
    from oracle import assoc
    rec, cnt = synth.clip_records(seed=11, frames=45, tracks=10, clutter=20, empty=(5, 6))
    fnos = list(range(100, 145))
    want = assoc.association({"v": synth.records_to_frame_relations(rec, cnt, fnos)})["v"]
    names_o, names_p = ["bg", "person", "dog", "ball"], ["ride", "chase", "hold"]
    got = sgg.association(torch.from_numpy(rec).cuda(), cnt, fnos, objects=names_o, predicates=names_p)
    assert len(got) == len(want) > 0
    for a, b in zip(got, want):
        assert a["triplet"] == [names_o[b["triplet"][0]], names_p[b["triplet"][1]], names_o[b["triplet"][2]]]
        assert {k: a[k] for k in a if k != "triplet"} == {k: b[k] for k in b if k != "triplet"}
    assert sgg.association(torch.zeros((4, 100, 13)).cuda(), np.zeros(4, np.int32)) == []
    # a gap in the frame numbers breaks every track (lib/utils.py:166 `fstart == r.fend`)
    gap = [f if f < 20 else f + 1 for f in range(45)]
    got_gap = sgg.association(torch.from_numpy(rec).cuda(), cnt, gap)
    want_gap = assoc.association({"v": synth.records_to_frame_relations(rec, cnt, gap)})["v"]
    assert got_gap == want_gap
