"""GPU parity tests (run on the B200 box): RoIAlign / RoIPool forward+backward through the C ABI against the CPU
oracle, and -- when oracle/_ref/libref_cuda.so travelled with the snapshot -- against the reference's own .cu
kernels compiled unmodified for sm_100a."""
import numpy as np
import pytest
import torch

from i2vsgg_b200 import synth

pytestmark = pytest.mark.gpu
SCALE = 1.0 / 16
# north_star: RoIAlign features and gradients within 1e-5 relative in fp32
RTOL = 1e-5


def close(got, want, rtol=RTOL):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    atol = rtol * float(np.abs(want).max()) if want.size else 0.0
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops():
    from i2vsgg_b200 import ops
    return ops


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


CASES = [  # (batch, channels, H, W, num_rois)
    (1, 32, 38, 63, 60),       # config-1 geometry, fewer channels so the oracle is quick
    (3, 16, 38, 63, 97),
    (2, 48, 20, 31, 50),       # runtime-width code path
]
FWD_CASES = CASES + [(2, 64, 63, 38, 120)]      # portrait map (lane-per-channel kernel: 33 + 31 row slabs)


def _fwd_combos():
    # every (case, impl, pool) the kernels take: the slab / even-pitch kernels have no max pool, the slab kernel needs
    # H * W = 2 (mod 4)
    for case in FWD_CASES:
        for impl in ("gather", "plane", "slab", "even", "chan", "auto"):
            for pool in ("none", "avg", "max"):
                if impl in ("slab", "even", "chan") and pool == "max":
                    continue
                if impl == "chan" and case[1] % 32:
                    continue
                if impl == "even" and case[2] > 40:
                    continue
                if impl == "slab" and (case[2] * case[3]) % 4 != 2:
                    continue
                yield case, impl, pool


@pytest.mark.parametrize("case,impl,pool", list(_fwd_combos()))
def test_roi_align_forward(ops, orc, case, impl, pool):
    B, C, H, W, N = case
    feat = synth.feature_map(100 + B, B, C, H, W)
    rois = synth.rois(200 + N, N, batch=B)
    if W != 63:
        rois[:, 1:] *= W / 63.0
    want = orc.roi_align_pooled_forward(feat, rois, 7, 7, SCALE, pool, nthreads=8)
    got = ops.roi_align_forward(cuda(feat), cuda(rois), 7, 7, SCALE, pool, impl)
    if impl == "gather" and pool in ("none", "avg", "max"):
        # the gather kernel evaluates the weights in double like roi_align_kernel.cu:64-67: bit-exact
        assert np.array_equal(got.cpu().numpy(), want)
    close(got, want)


@pytest.mark.parametrize("pool,p", [("none", 7), ("none", 4), ("avg", 7), ("max", 7), ("avg", 3)])
def test_roi_align_forward_other_sizes_and_bad_batch(ops, orc, pool, p):
    feat = synth.feature_map(7, 2, 8, 25, 40)
    rois = synth.rois(8, 33, batch=2)
    rois[:, 1:] *= 0.6
    want = orc.roi_align_pooled_forward(feat, rois, p, p, SCALE, pool)
    rois_bad = rois.copy()
    rois_bad[5, 0] = 7      # frame index outside the batch -> zero rows by contract
    rois_bad[9, 0] = -1
    want[5] = 0
    want[9] = 0
    got = ops.roi_align_forward(cuda(feat), cuda(rois_bad), p, p, SCALE, pool)
    assert np.array_equal(got.cpu().numpy(), want)


def test_roi_align_plane_zero_fills_bad_batch(ops, orc):
    feat = synth.feature_map(9, 2, 16, 38, 63)
    rois = synth.rois(10, 40, batch=2)
    want = orc.roi_align_pooled_forward(feat, rois, 7, 7, SCALE, "avg")
    rois[3, 0] = 2
    want[3] = 0
    got = ops.roi_align_forward(cuda(feat), cuda(rois), 7, 7, SCALE, "avg", "plane")
    close(got, want)
    assert float(got[3].abs().max()) == 0.0


def test_roi_align_empty(ops):
    feat = cuda(synth.feature_map(1, 1, 16, 38, 63))
    out = ops.roi_align_forward(feat, torch.zeros((0, 5), device="cuda"), 7, 7, SCALE, "avg")
    assert out.shape == (0, 16, 7, 7)
    g = ops.roi_align_backward(torch.zeros((0, 16, 7, 7), device="cuda"), None, torch.zeros((0, 5), device="cuda"),
                               (1, 16, 38, 63), 7, 7, SCALE, "avg")
    assert float(g.abs().max()) == 0.0


def _bwd_combos():
    # the max pool's arg-max routing needs the features: gather kernel (and `auto`, which takes it) only; the band-owner
    # kernel takes 32 channels per CTA
    for case in CASES + [(2, 64, 38, 63, 80)]:
        for impl in ("gather", "plane", "rows", "phase", "band", "auto"):
            for pool in ("none", "avg", "max"):
                if impl in ("plane", "rows", "phase", "band") and pool == "max":
                    continue
                if impl == "band" and case[1] % 32:
                    continue
                yield case, impl, pool


@pytest.mark.parametrize("case,impl,pool", list(_bwd_combos()))
def test_roi_align_backward(ops, orc, case, impl, pool):
    B, C, H, W, N = case
    feat = synth.feature_map(100 + B, B, C, H, W)
    rois = synth.rois(200 + N, N, batch=B)
    if W != 63:
        rois[:, 1:] *= W / 63.0
    g = np.random.default_rng(5).standard_normal((N, C, 7, 7)).astype(np.float32)
    want = orc.roi_align_pooled_backward(g, feat, rois, 7, 7, SCALE, pool, nthreads=8)
    got = ops.roi_align_backward(cuda(g), cuda(feat), cuda(rois), feat.shape, 7, 7, SCALE, pool, impl)
    close(got, want)


def test_roi_align_backward_small_and_repeated_cells(ops, orc):
    # tiny RoIs (several lattice points per cell in both axes), RoIs hugging the borders and many RoIs stacked on the
    # same cells: exercises the run-position serialisation and the column merge of the plane kernel
    rng = np.random.default_rng(17)
    B, C, H, W, N = 2, 32, 38, 63, 120
    feat_shape = (B, C, H, W)
    x1 = rng.uniform(0, 980, N); y1 = rng.uniform(0, 580, N)
    w = rng.choice([2.0, 9.0, 20.0, 45.0, 70.0, 130.0], N); h = rng.choice([2.0, 9.0, 20.0, 45.0, 70.0, 130.0], N)
    rois = np.stack([rng.integers(0, B, N), x1, y1, np.minimum(x1 + w, 999), np.minimum(y1 + h, 599)], 1).astype(np.float32)
    rois[:10] = rois[0]                     # identical RoIs
    rois[10:14] = [[0, 0, 0, 999, 599], [1, 990, 590, 999, 599], [0, 0, 560, 30, 599], [1, 960, 0, 999, 20]]
    g = rng.standard_normal((N, C, 7, 7)).astype(np.float32)
    for pool in ("avg", "none"):
        want = orc.roi_align_pooled_backward(g, None if pool != "max" else None, rois, 7, 7, SCALE, pool) \
            if False else orc.roi_align_pooled_backward(g, np.zeros(feat_shape, np.float32), rois, 7, 7, SCALE, pool)
        for impl in ("plane", "rows", "phase", "band"):
            got = ops.roi_align_backward(cuda(g), None, cuda(rois), feat_shape, 7, 7, SCALE, pool, impl)
            close(got, want)
            again = ops.roi_align_backward(cuda(g), None, cuda(rois), feat_shape, 7, 7, SCALE, pool, impl)
            assert torch.equal(got, again)      # no atomics: bit-reproducible


def test_roi_align_backward_frames_without_rois_are_zeroed(ops, orc):
    B, C, H, W = 3, 16, 38, 63
    rois = synth.rois(5, 20, batch=1)       # every RoI in frame 0
    g = np.random.default_rng(2).standard_normal((20, C, 7, 7)).astype(np.float32)
    gin = torch.full((B, C, H, W), 3.0, device="cuda")
    from i2vsgg_b200 import _lib
    import ctypes
    lib = _lib.load()
    r, gg = cuda(rois), cuda(g)
    ws = torch.empty(lib.i2v_roi_align_workspace_bytes(B, 20), dtype=torch.uint8, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.i2v_roi_align_backward(P(gg), None, P(r), P(gin), B, C, H, W, 20, 7, 7, SCALE, 1, 2, P(ws), ws.numel(),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    want = orc.roi_align_pooled_backward(g, np.zeros((B, C, H, W), np.float32), rois, 7, 7, SCALE, "avg")
    close(gin, want)
    assert float(gin[1:].abs().max()) == 0.0


def test_roi_align_autograd_module(ops, orc):
    import i2vsgg_b200
    i2vsgg_b200.install_as_model()
    from model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    feat = synth.feature_map(3, 2, 16, 38, 63)
    rois = synth.rois(4, 30, batch=2)
    g = np.random.default_rng(6).standard_normal((30, 16, 7, 7)).astype(np.float32)
    for mod, pool in ((RoIAlign, "none"), (RoIAlignAvg, "avg"), (RoIAlignMax, "max")):
        f = cuda(feat).requires_grad_(True)
        out = mod(7, 7, SCALE)(f, cuda(rois))
        out.backward(cuda(g))
        close(out, orc.roi_align_pooled_forward(feat, rois, 7, 7, SCALE, pool))
        close(f.grad, orc.roi_align_pooled_backward(g, feat, rois, 7, 7, SCALE, pool))


@pytest.mark.parametrize("mode", ["flat", "plane"])
def test_roi_pool_forward_backward(ops, orc, mode):
    B, C, H, W, N = 2, 24, 38, 63, 70
    feat = synth.feature_map(11, B, C, H, W)
    rois = synth.rois(12, N, batch=B)          # includes degenerate and past-the-edge RoIs
    g = np.random.default_rng(7).standard_normal((N, C, 7, 7)).astype(np.float32)
    if mode == "flat":
        want, warg = orc.roi_pool_forward(feat, rois, 7, 7, SCALE, nthreads=8)
        wgrad = orc.roi_pool_backward(g, rois, warg, feat.shape, 7, 7, SCALE)
        am = ops.ARGMAX_FLAT
    else:
        want, warg = orc.c_roi_pool_forward(feat, rois, 7, 7, SCALE, nthreads=8)
        wgrad = orc.c_roi_pool_backward(g, rois, warg, feat.shape, 7, 7)
        am = ops.ARGMAX_PLANE
    out, arg = ops.roi_pool_forward(cuda(feat), cuda(rois), 7, 7, SCALE, am)
    assert np.array_equal(out.cpu().numpy(), want)          # max and index: bit-exact
    assert np.array_equal(arg.cpu().numpy(), warg)
    grad = ops.roi_pool_backward(cuda(g), cuda(rois), arg, feat.shape, 7, 7, SCALE, am)
    close(grad, wgrad)


def test_roi_pool_modules(ops, orc):
    import i2vsgg_b200
    i2vsgg_b200.install_as_model()
    from model.roi_pooling.modules.roi_pool import _RoIPooling
    from model.roi_layers import ROIAlign, ROIPool
    feat = synth.feature_map(13, 2, 8, 38, 63)
    rois = synth.rois(14, 25, batch=2, degenerate=0)
    g = np.random.default_rng(8).standard_normal((25, 8, 7, 7)).astype(np.float32)
    f = cuda(feat).requires_grad_(True)
    out = _RoIPooling(7, 7, SCALE)(f, cuda(rois))
    out.backward(cuda(g))
    want, warg = orc.roi_pool_forward(feat, rois, 7, 7, SCALE)
    assert np.array_equal(out.detach().cpu().numpy(), want)
    close(f.grad, orc.roi_pool_backward(g, rois, warg, feat.shape, 7, 7, SCALE))
    f = cuda(feat).requires_grad_(True)
    out = ROIPool((7, 7), SCALE)(f, cuda(rois))
    out.backward(cuda(g))
    want, warg = orc.c_roi_pool_forward(feat, rois, 7, 7, SCALE)
    assert np.array_equal(out.detach().cpu().numpy(), want)
    close(f.grad, orc.c_roi_pool_backward(g, rois, warg, feat.shape, 7, 7))
    for sr in (0, 2):
        f = cuda(feat).requires_grad_(True)
        out = ROIAlign((7, 7), SCALE, sr)(f, cuda(rois))
        out.backward(cuda(g))
        close(out, orc.c_roi_align_forward(feat, rois, 7, 7, SCALE, sr), rtol=1e-5)
        close(f.grad, orc.c_roi_align_backward(g, rois, feat.shape, 7, 7, SCALE, sr), rtol=1e-5)


def test_c_ops_match_torchvision_cuda(ops):
    tv = pytest.importorskip("torchvision.ops")
    feat = cuda(synth.feature_map(15, 2, 16, 38, 63))
    rois = cuda(synth.rois(16, 40, batch=2, degenerate=0))
    want = tv.roi_align(feat, rois, (7, 7), SCALE, 0, aligned=False)
    got = ops.c_roi_align_forward(feat, rois, 7, 7, SCALE, 0)
    close(got, want.cpu().numpy(), rtol=1e-5)
    want = tv.roi_pool(feat, rois, (7, 7), SCALE)
    got, _ = ops.roi_pool_forward(feat, rois, 7, 7, SCALE, ops.ARGMAX_PLANE)
    assert torch.equal(got, want)


# ---------------------------------------------------------------- against the reference's own CUDA kernels
def _ref():
    from oracle import ref
    if not ref.have_cuda_ref():
        pytest.skip("oracle/_ref/libref_cuda.so did not travel with the snapshot")
    return ref


def test_reference_kernels_pin_the_oracle_and_us(ops, orc):
    ref = _ref()
    feat = synth.feature_map(21, 2, 16, 38, 63)
    rois = synth.rois(22, 64, batch=2)
    f, r = cuda(feat), cuda(rois)
    # forward: reference kernel vs oracle vs ours
    ref_out = ref.cuda_roi_align_forward(f, r, 8, 8, SCALE).cpu().numpy()
    want = orc.roi_align_forward(feat, rois, 8, 8, SCALE)
    # the reference's GPU build contracts `ph * bin + start` (roi_align_kernel.cu:44-45) into an FMA, its CPU twin
    # (roi_align.c:106-107, which the oracle pins bit for bit) does not: sample positions differ in the last bit,
    # i.e. by ~2e-6 cells, which moves values by up to ~2e-5.  That is the reference's own CPU/GPU spread.
    close(ref_out, want, rtol=1e-4)
    ours = ops.roi_align_forward(f, r, 8, 8, SCALE, "none", "gather")
    close(ours, ref_out, rtol=1e-4)
    # backward
    g = np.random.default_rng(9).standard_normal((64, 16, 8, 8)).astype(np.float32)
    ref_g = ref.cuda_roi_align_backward(cuda(g), r, feat.shape, 8, 8, SCALE).cpu().numpy()
    close(ref_g, orc.roi_align_backward(g, rois, feat.shape, 8, 8, SCALE), rtol=1e-4)
    close(ops.roi_align_backward(cuda(g), None, r, feat.shape, 8, 8, SCALE, "none"), ref_g, rtol=1e-4)
    # RoIPool fwd/bwd
    ref_p, ref_a = ref.cuda_roi_pool_forward(f, r, 7, 7, SCALE)
    wp, wa = orc.roi_pool_forward(feat, rois, 7, 7, SCALE)
    assert np.array_equal(ref_p.cpu().numpy(), wp) and np.array_equal(ref_a.cpu().numpy(), wa)
    g7 = np.random.default_rng(10).standard_normal((64, 16, 7, 7)).astype(np.float32)
    ref_pg = ref.cuda_roi_pool_backward(cuda(g7), r, ref_a, feat.shape, 7, 7, SCALE).cpu().numpy()
    close(ref_pg, orc.roi_pool_backward(g7, rois, wa, feat.shape, 7, 7, SCALE))
    close(ops.roi_pool_backward(cuda(g7), r, ref_a, feat.shape, 7, 7, SCALE, ops.ARGMAX_FLAT), ref_pg)


def test_legacy_launchers(ops, orc):
    import ctypes
    from i2vsgg_b200 import _lib
    lib = _lib.load()
    feat = synth.feature_map(31, 2, 8, 38, 63)
    rois = synth.rois(32, 20, batch=2)
    f, r = cuda(feat), cuda(rois)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    out = torch.empty((20, 8, 8, 8), device="cuda")
    assert lib.ROIAlignForwardLaucher(P(f), SCALE, 20, 38, 63, 8, 8, 8, P(r), P(out), s) == 1
    assert np.array_equal(out.cpu().numpy(), orc.roi_align_forward(feat, rois, 8, 8, SCALE))
    g = np.random.default_rng(11).standard_normal((20, 8, 8, 8)).astype(np.float32)
    gin = torch.zeros((2, 8, 38, 63), device="cuda")
    assert lib.ROIAlignBackwardLaucher(P(cuda(g)), SCALE, 2, 20, 38, 63, 8, 8, 8, P(r), P(gin), s) == 1
    close(gin, orc.roi_align_backward(g, rois, feat.shape, 8, 8, SCALE))
    po = torch.empty((20, 8, 7, 7), device="cuda")
    pa = torch.empty((20, 8, 7, 7), device="cuda", dtype=torch.int32)
    assert lib.ROIPoolForwardLaucher(P(f), SCALE, 20, 38, 63, 8, 7, 7, P(r), P(po), P(pa), s) == 1
    wp, wa = orc.roi_pool_forward(feat, rois, 7, 7, SCALE)
    assert np.array_equal(po.cpu().numpy(), wp) and np.array_equal(pa.cpu().numpy(), wa)
    g7 = np.random.default_rng(12).standard_normal((20, 8, 7, 7)).astype(np.float32)
    pg = torch.full((2, 8, 38, 63), 7.0, device="cuda")
    assert lib.ROIPoolBackwardLaucher(P(cuda(g7)), SCALE, 2, 20, 38, 63, 8, 7, 7, P(r), P(pg), P(pa), s) == 1
    close(pg, orc.roi_pool_backward(g7, rois, wa, feat.shape, 7, 7, SCALE))


def test_legacy_launchers_take_the_fast_kernels(ops):
    """roi_align_kernel.h:13-27 at config-2 width (1024 channels, 7x7 lattice): the launchers must run the plane-resident
    kernels -- same numbers as the i2v_* entry points (the backward ADDS into the caller's buffer like the reference's
    atomicAdd scatter) and within 1.3x of their time.  The forward launcher pays one 4-byte read-back per call for the
    frame count its signature lacks."""
    import ctypes
    from i2vsgg_b200 import _lib
    lib = _lib.load()
    B, C, H, W, N = 8, 1024, 38, 63, 2400
    gen = torch.Generator(device="cuda").manual_seed(5)
    feat = torch.randn((B, C, H, W), device="cuda", generator=gen)
    rois = cuda(synth.rois(77, N, batch=B, sort_by_batch=False))
    grad = torch.randn((N, C, 7, 7), device="cuda", generator=gen)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    out = torch.empty((N, C, 7, 7), device="cuda")
    fwd_legacy = lambda: lib.ROIAlignForwardLaucher(P(feat), SCALE, N, H, W, C, 7, 7, P(rois), P(out), s)
    assert fwd_legacy() == 1
    want = ops.roi_align_forward(feat, rois, 7, 7, SCALE, "none", "auto")
    assert torch.equal(out, want)
    t_legacy, t_entry = timed(fwd_legacy), timed(lambda: ops.roi_align_forward(feat, rois, 7, 7, SCALE, "none", "auto"))
    assert t_legacy <= 1.3 * t_entry, (t_legacy, t_entry)

    gin = torch.zeros((B, C, H, W), device="cuda")
    bwd_legacy = lambda: lib.ROIAlignBackwardLaucher(P(grad), SCALE, B, N, H, W, C, 7, 7, P(rois), P(gin), s)
    assert bwd_legacy() == 1
    gwant = ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "none", "auto")
    assert torch.equal(gin, gwant)
    assert bwd_legacy() == 1                                   # a second call adds the same planes again
    assert float((gin - 2 * gwant).abs().max()) <= 1e-6 * float(gwant.abs().max())
    t_legacy = timed(bwd_legacy)
    t_entry = timed(lambda: ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "none", "auto"))
    assert t_legacy <= 1.3 * t_entry, (t_legacy, t_entry)


def _portrait_rois(seed, n, batch):
    r = synth.rois(seed, n, batch=batch)
    return np.ascontiguousarray(r[:, [0, 2, 1, 4, 3]])          # swap x and y: boxes of a 1000 x 600 (h x w) frame


@pytest.mark.parametrize("pool", ["none", "avg", "max"])
def test_portrait_maps_stay_on_the_plane_kernels(ops, orc, pool):
    """A 1000 x 600 frame gives a 63 x 38 map: the plane kernels take it with the 40-cell row pitch (forward) and the
    orientation-independent plane size (backward) instead of falling back to the gather kernels."""
    B, C, H, W, N = 2, 32, 63, 38, 80
    feat = synth.feature_map(300, B, C, H, W)
    rois = _portrait_rois(301, N, B)
    want = orc.roi_align_pooled_forward(feat, rois, 7, 7, SCALE, pool, nthreads=8)
    got = ops.roi_align_forward(cuda(feat), cuda(rois), 7, 7, SCALE, pool, "plane")
    close(got, want)
    if pool != "max":
        g = np.random.default_rng(7).standard_normal((N, C, 7, 7)).astype(np.float32)
        wantg = orc.roi_align_pooled_backward(g, feat, rois, 7, 7, SCALE, pool, nthreads=8)
        gotg = ops.roi_align_backward(cuda(g), None, cuda(rois), feat.shape, 7, 7, SCALE, pool, "plane")
        close(gotg, wantg)


def test_portrait_roi_pool_rows(ops):
    from i2vsgg_b200._lib import ARGMAX_PLANE
    feat = cuda(synth.feature_map(302, 1, 32, 63, 38))
    rois = cuda(_portrait_rois(303, 60, 1))
    want, _ = ops.roi_pool_forward(feat, rois, 7, 7, SCALE, ARGMAX_PLANE)
    got = ops.roi_pool_rows(feat, rois, 7, 7, SCALE, dtype=torch.float32)
    assert torch.equal(got, want.reshape(60, -1))


def test_roi_align_properties_at_config2_size(ops, orc):
    """BASELINE.json configs[1] at full size (32 frames x 1024 channels, 9600 RoIs from the proposal layer): the oracle
    cannot run this in seconds, so parity is carried by properties that hold at any size:
      * sampled RoIs of the plane forward equal the gather kernel on those RoIs (itself bit-exact against roi_align.c)
        within 1e-5, and a few of them equal the CPU oracle;
      * <forward(F), G> == <F, backward(G)> (the backward is the adjoint of the forward), which ties the two directions
        together over all 9600 x 1024 x 49 outputs;
      * both backward kernels agree, and each is bit-reproducible."""
    B, C, H, W, post = 32, 1024, 38, 63, 300
    g = torch.Generator(device="cuda").manual_seed(11)
    feat = torch.randn((B, C, H, W), device="cuda", generator=g)
    cls, reg = synth.rpn_outputs(500, batch=B)
    rois = ops.proposal_forward(cuda(cls), cuda(reg), cuda(synth.im_info(B)), cuda(synth.BASE_ANCHORS), 16, 12000, post,
                                0.7).reshape(-1, 5)
    N = rois.size(0)
    out = ops.roi_align_forward(feat, rois, 7, 7, SCALE, "avg", "plane")
    assert out.shape == (N, C, 7, 7) and bool(torch.isfinite(out).all())
    pick = torch.from_numpy(np.random.default_rng(1).choice(N, 48, replace=False)).cuda()
    ref = ops.roi_align_forward(feat, rois[pick], 7, 7, SCALE, "avg", "gather")
    scale = float(ref.abs().max())
    assert float((out[pick] - ref).abs().max()) <= 1e-5 * scale
    few = pick[:3].cpu().numpy()
    r3 = rois[pick[:3]].cpu().numpy()
    b3 = r3[:, 0].astype(int)
    sub = feat[torch.from_numpy(np.unique(b3)).cuda()].cpu().numpy()
    remap = {b: i for i, b in enumerate(np.unique(b3))}
    r3[:, 0] = [remap[b] for b in b3]
    want = orc.roi_align_pooled_forward(sub, r3, 7, 7, SCALE, "avg", nthreads=8)
    close(out[torch.from_numpy(few).cuda()], want)

    # the kernel bench.py times (`auto` -> the slab kernel on a 38x63 map), held to the same bar
    for impl in ("auto", "even", "slab", "chan"):
        fast = ops.roi_align_forward(feat, rois, 7, 7, SCALE, "avg", impl)
        assert float((fast[pick] - ref).abs().max()) <= 1e-5 * scale, impl
        close(fast[torch.from_numpy(few).cuda()], want)
        assert float((fast - out).abs().max()) <= 1e-5 * scale           # all 9600 x 1024 x 49 outputs against the plane kernel
        del fast

    grad = torch.randn((N, C, 7, 7), device="cuda", generator=g)
    gin = ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", "plane")
    lhs = float((out.double() * grad.double()).sum())
    rhs = float((feat.double() * gin.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), float((out.double().abs() * grad.double().abs()).sum()) * 1e-2)
    again = ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", "plane")
    assert torch.equal(gin, again)
    rows = ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", "rows")
    assert float((rows - gin).abs().max()) <= 1e-5 * float(gin.abs().max())
    assert torch.equal(rows, ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", "rows"))
    del rows, again
    phase = ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", "phase")
    assert float((phase - gin).abs().max()) <= 1e-5 * float(gin.abs().max())
    assert torch.equal(phase, ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", "phase"))
    del phase
    # the kernel `auto` picks (band-owner): against the plane kernel, bit-reproducible, and the adjoint identity again
    for impl in ("band", "auto"):
        band = ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", impl)
        assert float((band - gin).abs().max()) <= 1e-5 * float(gin.abs().max())
        assert torch.equal(band, ops.roi_align_backward(grad, None, rois, (B, C, H, W), 7, 7, SCALE, "avg", impl))
        rhs = float((feat.double() * band.double()).sum())
        assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), float((out.double().abs() * grad.double().abs()).sum()) * 1e-2)
        del band


def test_roi_align_backward_at_config4_shape(ops, orc):
    """BASELINE.json configs[3]: the style-discriminator training step's RoIAlign backward, 8 images x 256 RoIs
    (TRAIN.BATCH_SIZE, cfgs/vgg16.yml:9) over 1024 x 38 x 63, upstream gradient [2048,1024,7,7].  The oracle is run on
    sampled channel groups (all 2048 RoIs, 3 x 16 channels); every kernel `auto` could take must match it, agree with
    the others over the whole tensor and be bit-reproducible."""
    B, C, H, W, per = 8, 1024, 38, 63, 256
    rois = synth.rois(401, B * per, batch=B, sort_by_batch=True)
    N = rois.shape[0]
    gen = torch.Generator(device="cuda").manual_seed(41)
    grad = torch.randn((N, C, 7, 7), device="cuda", generator=gen)
    r = cuda(rois)
    outs = {impl: ops.roi_align_backward(grad, None, r, (B, C, H, W), 7, 7, SCALE, "avg", impl)
            for impl in ("auto", "phase", "band")}
    for c0 in (0, 496, 1008):
        g = grad[:, c0:c0 + 16].contiguous().cpu().numpy()
        want = orc.roi_align_pooled_backward(g, np.zeros((B, 16, H, W), np.float32), rois, 7, 7, SCALE, "avg", nthreads=8)
        for impl, got in outs.items():
            close(got[:, c0:c0 + 16], want)
    scale = float(outs["auto"].abs().max())
    assert float((outs["band"] - outs["phase"]).abs().max()) <= 1e-5 * scale
    for impl, got in outs.items():
        assert torch.equal(got, ops.roi_align_backward(grad, None, r, (B, C, H, W), 7, 7, SCALE, "avg", impl))


@pytest.mark.parametrize("shape", [(1, 16, 60, 80), (2, 24, 38, 63)])   # planes too large for shared memory; C % 16 != 0
def test_roi_align_backward_auto_falls_back_when_the_plane_kernels_cannot_run(ops, orc, shape):
    """`auto` must pick a kernel that supports the shape (here: the gather kernel) and asking for a plane-resident kernel
    explicitly must fail loudly instead of computing something else."""
    from i2vsgg_b200._lib import I2VError
    B, C, H, W = shape
    feat = synth.feature_map(7, B, C, H, W)
    rois = synth.rois(8, 40, batch=B)
    rois[:, 1:] *= np.float32([W / 63.0, H / 38.0, W / 63.0, H / 38.0])
    g = np.random.default_rng(6).standard_normal((40, C, 7, 7)).astype(np.float32)
    want = orc.roi_align_pooled_backward(g, feat, rois, 7, 7, SCALE, "avg", nthreads=8)
    got = ops.roi_align_backward(cuda(g), None, cuda(rois), feat.shape, 7, 7, SCALE, "avg", "auto")
    close(got, want)
    for impl in ("phase", "plane", "rows"):
        with pytest.raises(I2VError):
            ops.roi_align_backward(cuda(g), None, cuda(rois), feat.shape, 7, 7, SCALE, "avg", impl)


def test_roi_align_backward_large_map_stays_plane_resident(ops, orc):
    """A 60x80 map (16 planes = 307 KB) does not fit the phased kernel; with C % 32 == 0 `auto` takes the band-owner
    kernel, which cuts the map into slabs of rows (here four) -- against the oracle, and bit-reproducible."""
    B, C, H, W, N = 2, 32, 60, 80, 70
    feat_shape = (B, C, H, W)
    rois = synth.rois(31, N, batch=B)
    rois[:, 1:] *= np.float32([W / 63.0, H / 38.0, W / 63.0, H / 38.0])
    g = np.random.default_rng(8).standard_normal((N, C, 7, 7)).astype(np.float32)
    for pool in ("avg", "none"):
        want = orc.roi_align_pooled_backward(g, np.zeros(feat_shape, np.float32), rois, 7, 7, SCALE, pool, nthreads=8)
        for impl in ("band", "auto"):
            got = ops.roi_align_backward(cuda(g), None, cuda(rois), feat_shape, 7, 7, SCALE, pool, impl)
            close(got, want)
            assert torch.equal(got, ops.roi_align_backward(cuda(g), None, cuda(rois), feat_shape, 7, 7, SCALE, pool, impl))


def test_roi_align_kernels_agree_on_random_shapes(ops):
    """Forty random (frames, channels, map, RoI) configurations -- tiny and huge boxes, boxes over the border, stray frame
    indices, maps of every aspect ratio that fits shared memory: every plane-resident backward kernel a shape supports must
    agree with the gather kernel (the reference's atomic scatter) to 1e-5 of scale (north_star's bar) and be
    bit-reproducible, and the forwards that `auto` / `slab` / `plane` pick must agree with the gather forward to 1e-5."""
    from i2vsgg_b200._lib import I2VError
    rng = np.random.default_rng(2024)
    checked = 0
    for trial in range(40):
        B = int(rng.integers(1, 4)); C = 16 * int(rng.integers(1, 4))
        H = int(rng.integers(2, 48)); W = int(rng.integers(2, 70))
        if H * W * 64 + 80000 > 232448:
            continue
        N = int(rng.integers(1, 120))
        iw, ih = W * 16.0, H * 16.0
        x1 = rng.uniform(-30, iw, N); y1 = rng.uniform(-30, ih, N)
        w = rng.choice([1.0, 6.0, 20.0, 60.0, 200.0, 700.0], N); h = rng.choice([1.0, 6.0, 20.0, 60.0, 200.0, 700.0], N)
        rois = cuda(np.stack([rng.integers(-1, B + 1, N), x1, y1, x1 + w, y1 + h], 1).astype(np.float32))
        gen = torch.Generator(device="cuda").manual_seed(trial)
        g = torch.randn((N, C, 7, 7), device="cuda", generator=gen)
        feat = torch.randn((B, C, H, W), device="cuda", generator=gen)
        for pool in ("avg", "none"):
            ref = ops.roi_align_backward(g, None, rois, (B, C, H, W), 7, 7, SCALE, pool, "gather")
            scale = max(float(ref.abs().max()), 1e-6)
            for impl in ("phase", "plane", "rows", "band"):
                try:
                    out = ops.roi_align_backward(g, None, rois, (B, C, H, W), 7, 7, SCALE, pool, impl)
                except I2VError:
                    continue                     # the shape is outside what this kernel takes
                assert float((out - ref).abs().max()) <= 1e-5 * scale, (trial, B, C, H, W, N, pool, impl)
                assert torch.equal(out, ops.roi_align_backward(g, None, rois, (B, C, H, W), 7, 7, SCALE, pool, impl))
                checked += 1
            fref = ops.roi_align_forward(feat, rois, 7, 7, SCALE, pool, "gather")
            for fimpl in ("auto", "even", "slab", "chan", "plane"):
                try:
                    fout = ops.roi_align_forward(feat, rois, 7, 7, SCALE, pool, fimpl)
                except I2VError:
                    continue
                assert float((fout - fref).abs().max()) <= 1e-5 * max(float(fref.abs().max()), 1e-6), (trial, H, W, pool, fimpl)
    assert checked >= 100


@pytest.mark.parametrize("mode", ["flat", "plane"])
@pytest.mark.parametrize("shape", [(2, 32, 38, 63, 90), (1, 16, 63, 38, 40), (3, 48, 20, 31, 60)])
def test_roi_pool_plane_resident_kernels_equal_the_per_element_ones(ops, orc, mode, shape):
    """The opt-in plane-resident forward (values + arg-max as two TMA stores) and owner-warp backward (no atomics; both
    I2V_POOL_PLANE=1) against the default per-element kernels and the oracle: values and arg-max bit for bit -- ties (duplicated feature
    values), RoIs over the border, tiny RoIs (bins under one cell: the bin-by-bin path of the backward), stray frame
    indices -- and the backward bit-reproducible."""
    import os
    from i2vsgg_b200._lib import ARGMAX_FLAT, ARGMAX_PLANE
    B, C, H, W, N = shape
    m = ARGMAX_FLAT if mode == "flat" else ARGMAX_PLANE
    rng = np.random.default_rng(H + N)
    feat = np.round(rng.standard_normal((B, C, H, W)) * 2).astype(np.float32) / 2       # many exact ties
    iw, ih = W * 16.0, H * 16.0
    x1 = rng.uniform(-40, iw - 10, N); y1 = rng.uniform(-40, ih - 10, N)
    w = rng.choice([3.0, 30.0, 90.0, 250.0, 600.0], N); h = rng.choice([3.0, 30.0, 90.0, 250.0, 600.0], N)
    rois = np.stack([rng.integers(0, B, N), x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    rois[4, 0] = B + 2
    g = rng.standard_normal((N, C, 7, 7)).astype(np.float32)
    f, r, go = cuda(feat), cuda(rois), cuda(g)
    ref_out, ref_arg = ops.roi_pool_forward(f, r, 7, 7, SCALE, m)              # the default per-element kernels
    ref_gin = ops.roi_pool_backward(go, r, ref_arg, feat.shape, 7, 7, SCALE, m)
    os.environ["I2V_POOL_PLANE"] = "1"                                          # read per call
    try:
        out, arg = ops.roi_pool_forward(f, r, 7, 7, SCALE, m)
        gin = ops.roi_pool_backward(go, r, arg, feat.shape, 7, 7, SCALE, m)
        again = ops.roi_pool_backward(go, r, arg, feat.shape, 7, 7, SCALE, m)
    finally:
        del os.environ["I2V_POOL_PLANE"]
    assert torch.equal(out, ref_out) and torch.equal(arg, ref_arg) and torch.equal(gin, again)
    assert float((gin - ref_gin).abs().max()) <= 1e-5 * float(ref_gin.abs().max())
    ok = rois[:, 0] < B
    if mode == "flat":
        wo, wa = orc.roi_pool_forward(feat, rois[ok], 7, 7, SCALE)
        wg = orc.roi_pool_backward(g[ok], rois[ok], wa, feat.shape, 7, 7, SCALE)
    else:
        wo, wa = orc.c_roi_pool_forward(feat, rois[ok], 7, 7, SCALE)
        wg = orc.c_roi_pool_backward(g[ok], rois[ok], wa, feat.shape, 7, 7)
    okd = torch.from_numpy(ok).cuda()
    assert np.array_equal(out[okd].cpu().numpy(), wo) and np.array_equal(arg[okd].cpu().numpy(), wa)
    assert float(out[~okd].abs().max()) == 0.0 and int(arg[~okd].max()) == -1
    close(gin, wg)
