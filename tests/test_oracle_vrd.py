"""CPU: the numpy restatement of `vrd.forward` (oracle.vrd_forward) against the outputs of the UNMODIFIED reference
module (tests/golden/vrd_golden.npz, written by tests/golden/make_vrd_golden.py in the build container)."""
import importlib.util
import os

import numpy as np
import pytest

from i2vsgg_b200 import synth
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _gen():
    spec = importlib.util.spec_from_file_location("make_vrd_golden", os.path.join(HERE, "golden", "make_vrd_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def vrd_golden():
    return np.load(os.path.join(HERE, "golden", "vrd_golden.npz"))


@pytest.mark.parametrize("tag,kw", [("full", {}), ("novis_loc1", dict(use_obj_visual=False, spatial_type=1))])
def test_vrd_oracle_matches_the_reference_module(vrd_golden, tag, kw):
    gen = _gen()
    args = synth.VrdArgs(**kw)
    params = synth.vrd_params(gen.PARAM_SEED, args)
    prd = synth.prd_vectors(gen.PRD_SEED, args.num_relations)
    fmap, boxes, rel, masks, classes, ixs, ixo = gen.inputs()
    spatial = masks if args.spatial_type == 2 else np.random.default_rng(3).standard_normal((len(ixs), 8), dtype=np.float32)
    scores, feat = oracle.vrd_forward(params, prd, fmap, boxes, rel, spatial, ixs, ixo, args.use_obj_visual,
                                      args.spatial_type, nthreads=oracle.default_threads())
    want_s, want_f = vrd_golden[f"{tag}_scores"], vrd_golden[f"{tag}_feat"]
    # fp32 GEMMs with a different summation order (numpy/OpenBLAS vs torch/MKL) over K = 50176
    np.testing.assert_allclose(feat, want_f, rtol=0, atol=2e-4 * float(np.abs(want_f).max()))
    np.testing.assert_allclose(scores, want_s, rtol=1e-4, atol=1e-7)
    # the `rows` shortcut used by the full-size GPU spot check computes the same thing
    s2, f2 = oracle.vrd_forward(params, prd, fmap, boxes, rel, spatial, ixs, ixo, args.use_obj_visual,
                                args.spatial_type, rows=[5, 0, 11], nthreads=oracle.default_threads())
    np.testing.assert_allclose(f2, feat[[5, 0, 11]], rtol=0, atol=1e-5 * float(np.abs(feat).max()))
    np.testing.assert_allclose(s2, scores[[5, 0, 11]], rtol=1e-5, atol=1e-8)


def test_vrd_oracle_training_mode_matches_the_reference_module(vrd_golden):
    """The reference module executed in TRAINING mode (dropout with the supplied keep masks, raw cosine similarities:
    resnet_SGG_emb.py:148-151,215) pins the oracle's training-mode restatement."""
    gen = _gen()
    args = synth.VrdArgs()
    params = synth.vrd_params(gen.PARAM_SEED, args)
    prd = synth.prd_vectors(gen.PRD_SEED, args.num_relations)
    fmap, boxes, rel, masks, classes, ixs, ixo = gen.inputs()
    scores, feat = oracle.vrd_forward(params, prd, fmap, boxes, rel, masks, ixs, ixo, nthreads=oracle.default_threads(),
                                      dropout_masks=gen.train_masks())
    want_s, want_f = vrd_golden["train_scores"], vrd_golden["train_feat"]
    np.testing.assert_allclose(feat, want_f, rtol=0, atol=2e-4 * float(np.abs(want_f).max()))
    np.testing.assert_allclose(scores, want_s, rtol=0, atol=1e-4)
    assert float(np.abs(want_s).max()) <= 1.0 and abs(float(want_s.sum(1)[0]) - 1.0) > 1e-3      # cosines, not probabilities
