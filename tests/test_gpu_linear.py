"""GPU parity tests (B200 box): the tcgen05 FC kernel, the bf16 cast, pooled rows and the relation scores.

The FC kernel is compared with a float64 product of the SAME operands (bf16 values are exact in float64, tf32 operands
are truncated to 19 bits like the tensor core does), so the only difference left is the fp32 accumulation order:
tolerance 2e-5 of the output scale.  Against the unrounded fp32 operands the error is the documented quantisation of
the tensor-core input format (bf16: 2^-9 per operand)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from i2vsgg_b200 import ops
    return ops


def _tf32_trunc(t: torch.Tensor) -> torch.Tensor:
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def _ref(x, w, bias, relu):
    y = x.double() @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    return torch.relu(y) if relu else y


SHAPES = [(128, 256, 64), (128, 256, 512), (1, 8, 8), (300, 300, 600), (257, 132, 304), (4096, 512, 3136),
          (64, 4096, 1024), (520, 768, 72)]


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("relu", [False, True])
def test_linear_bf16_matches_float64_of_same_operands(ops, m, n, k, relu):
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n * 3 + k)
    x = torch.randn((m, k), device="cuda", generator=g).bfloat16()
    w = (torch.randn((n, k), device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn((n,), device="cuda", generator=g)
    y = ops.linear(x, w, b, relu=relu)
    want = _ref(x, w, b, relu)
    scale = float(want.abs().max()) + 1e-30
    assert y.dtype == torch.float32 and y.shape == (m, n)
    assert float((y.double() - want).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("m,n,k", [(128, 256, 32), (300, 300, 600), (1000, 260, 1024)])
def test_linear_tf32_matches_float64_of_truncated_operands(ops, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    x = torch.randn((m, k), device="cuda", generator=g)
    w = torch.randn((n, k), device="cuda", generator=g) * 0.05
    y = ops.linear(x, w, None, relu=False)
    want = _ref(_tf32_trunc(x), _tf32_trunc(w), None, False)
    scale = float(want.abs().max())
    # the tensor core may round instead of truncate the low 13 bits: allow the tf32 quantisation itself
    full = _ref(x, w, None, False)
    err_trunc = float((y.double() - want).abs().max())
    err_full = float((y.double() - full).abs().max())
    assert min(err_trunc, err_full) <= 2e-5 * scale or err_full <= 2e-3 * scale, (err_trunc, err_full, scale)


def test_linear_bf16_output_and_column_slices(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((515, 320), device="cuda", generator=g).bfloat16()
    w1 = (torch.randn((256, 320), device="cuda", generator=g) * 0.05).bfloat16()
    w2 = (torch.randn((256, 320), device="cuda", generator=g) * 0.05).bfloat16()
    cat = torch.full((515, 768), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.linear(x, w1, None, relu=True, out=cat[:, 0:256])
    ops.linear(x, w2, None, relu=True, out=cat[:, 512:768])
    want1 = torch.relu(x.double() @ w1.double().t())
    want2 = torch.relu(x.double() @ w2.double().t())
    assert torch.all(cat[:, 256:512] == 7.0)                       # the untouched slice stays untouched
    for got, want in ((cat[:, 0:256], want1), (cat[:, 512:768], want2)):
        assert float((got.double() - want).abs().max()) <= 2 ** -8 * float(want.abs().max())   # one bf16 rounding
    # strided input rows: a column slice of a wider activation matrix
    wide = torch.randn((200, 640), device="cuda", generator=g).bfloat16()
    y = ops.linear(wide[:, 320:640], w1, None)
    assert float((y.double() - wide[:, 320:640].double() @ w1.double().t()).abs().max()) <= 2e-5 * float(y.abs().max())


def test_linear_rejects_bad_arguments(ops):
    from i2vsgg_b200._lib import I2VError
    x = torch.zeros((4, 10), device="cuda", dtype=torch.bfloat16)      # 20-byte rows: not TMA-addressable
    w = torch.zeros((8, 10), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(I2VError):
        ops.linear(x, w)
    with pytest.raises(I2VError):
        ops.linear(torch.zeros((4, 16), device="cuda"), torch.zeros((8, 16), device="cuda", dtype=torch.bfloat16))
    with pytest.raises(I2VError):
        ops.linear(torch.zeros((4, 16)), torch.zeros((8, 16)))       # CPU tensors: there is no CPU path


def test_cast_bf16_is_round_to_nearest_even(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn((37, 1000), device="cuda", generator=g) * 100
    assert torch.equal(ops.cast_bf16(a), a.bfloat16())
    wide = torch.randn((16, 300), device="cuda", generator=g)
    assert torch.equal(ops.cast_bf16(wide[:, 100:228]), wide[:, 100:228].bfloat16())


# the tiled kernel (>= 64 rows, E % 4 == 0, R <= 160) with full and ragged row tiles, and the row-per-warp kernel
@pytest.mark.parametrize("p,r,e", [(4032, 132, 300), (4035, 132, 300), (100, 40, 64), (65, 160, 8), (7, 5, 33), (1, 132, 300),
                                   (70, 161, 12)])
def test_rel_scores_match_torch(ops, p, r, e):
    g = torch.Generator(device="cuda").manual_seed(p + r)
    x = torch.randn((p, e), device="cuda", generator=g)
    prd = torch.randn((r, e), device="cuda", generator=g)
    x[0] = 0                                                        # F.normalize's eps branch
    got = ops.rel_scores(x, prd, softmax=True)
    sim = torch.nn.functional.normalize(x.double(), dim=1) @ torch.nn.functional.normalize(prd.double(), dim=1).t()
    want = torch.softmax(sim, dim=1)
    assert float((got.double() - want).abs().max()) <= 1e-6
    got_raw = ops.rel_scores(x, prd, softmax=False)
    assert float((got_raw.double() - sim).abs().max()) <= 1e-6


@pytest.mark.parametrize("channels,batch", [(64, 1), (24, 1), (32, 3)])   # plane kernel, gather fallback, several frames
def test_roi_pool_rows_equal_the_roi_pool_op(ops, channels, batch):
    from i2vsgg_b200 import synth
    from i2vsgg_b200._lib import ARGMAX_PLANE
    feat = torch.from_numpy(synth.feature_map(3, batch, channels)).cuda()
    r = synth.rois(4, 50, batch)
    if batch > 1:
        r[7, 0], r[19, 0] = -1, batch + 2            # stray frame indices: zero rows
    rois = torch.from_numpy(r).cuda()
    want, _ = ops.roi_pool_forward(feat, rois, 7, 7, 1 / 16, ARGMAX_PLANE)
    rows32 = ops.roi_pool_rows(feat, rois, 7, 7, 1 / 16, dtype=torch.float32)
    assert torch.equal(rows32, want.reshape(50, -1))
    big = torch.zeros((60, channels * 49), device="cuda", dtype=torch.bfloat16)
    ops.roi_pool_rows(feat, rois, 7, 7, 1 / 16, out=big[10:60])
    assert torch.equal(big[10:60], want.reshape(50, -1).bfloat16()) and torch.all(big[:10] == 0)


@pytest.mark.parametrize("layout,c,h,k,stride,pad,dtype", [("nchw", 2, 32, 5, 2, 2, torch.float32),
                                                            ("nhwc", 96, 16, 5, 2, 2, torch.bfloat16),
                                                            ("nhwc", 12, 9, 3, 1, 0, torch.bfloat16),
                                                            ("nhwc", 128, 8, 8, 1, 0, torch.bfloat16)])
def test_im2col_matches_unfold(ops, layout, c, h, k, stride, pad, dtype):
    g = torch.Generator(device="cuda").manual_seed(c + h)
    n = 5
    x = torch.randn((n, c, h, h), device="cuda", generator=g).to(dtype)
    xin = x if layout == "nchw" else x.permute(0, 2, 3, 1).contiguous()
    kk = k * k * c
    ld = (kk + 7) // 8 * 8
    got, (n2, oh, ow) = ops.im2col_bf16(xin, k, stride, pad, layout, ld=ld)
    # torch's unfold orders a patch (c, ky, kx); ours is (ky, kx, c)
    want = torch.nn.functional.unfold(x.float(), k, padding=pad, stride=stride)          # [n, c*k*k, oh*ow]
    want = want.view(n, c, k, k, oh * ow).permute(0, 4, 2, 3, 1).reshape(n * oh * ow, kk).bfloat16()
    assert (n2, got.shape[0]) == (n, n * oh * ow)
    assert torch.equal(got[:, :kk], want) and torch.all(got[:, kk:] == 0)
    got1, _ = ops.im2col_bf16(xin, k, stride, pad, layout, ld=kk + 1 if (kk + 1) % 8 else kk + 3)   # scalar-store path
    assert torch.equal(got1[:, :kk], want)


@pytest.mark.parametrize("c,o,h,k,stride,pad", [(96, 128, 16, 5, 2, 2), (64, 32, 16, 3, 2, 1), (8, 128, 8, 3, 1, 1)])
def test_conv2d_nhwc_implicit_gemm_matches_torch(ops, c, o, h, k, stride, pad):
    """The 4-D TMA map (element strides = conv stride, zero-filled borders and channel padding) against F.conv2d in
    float64 on the same bf16 operands."""
    g = torch.Generator(device="cuda").manual_seed(c + o)
    n = 37
    x = torch.randn((n, c, h, h), device="cuda", generator=g).bfloat16()
    w = (torch.randn((o, c, k, k), device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn((o,), device="cuda", generator=g)
    cp = (c + 63) // 64 * 64
    taps = torch.zeros((o, k * k, cp), device="cuda", dtype=torch.bfloat16)
    taps[:, :, :c] = w.permute(0, 2, 3, 1).reshape(o, k * k, c)
    y = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous(), taps.view(o, -1), b, k, stride, pad, relu=True,
                        out_dtype=torch.float32)
    want = torch.relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad))
    assert y.shape == (n, want.shape[2], want.shape[3], o)
    assert float((y.permute(0, 3, 1, 2).double() - want).abs().max()) <= 2e-5 * float(want.abs().max())


@pytest.mark.parametrize("c,o,h,k", [(96, 128, 16, 5), (64, 32, 16, 3), (8, 128, 8, 3), (72, 100, 4, 7)])
def test_conv2d_on_parity_planes_equals_the_strided_call(ops, c, o, h, k):
    """Stride-2 convolution on the parity-split activation ([N,2,2,H/2,W/2,C], dense 5-D TMA boxes) against the strided
    4-D map on the plain NHWC tensor: the same k-blocks in the same order, so the same bits; and against F.conv2d."""
    g = torch.Generator(device="cuda").manual_seed(c * 3 + o)
    n, pad = 37, (k - 1) // 2
    x = torch.randn((n, h, h, c), device="cuda", generator=g).bfloat16()
    w = (torch.randn((o, c, k, k), device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn((o,), device="cuda", generator=g)
    cp = (c + 63) // 64 * 64
    taps = torch.zeros((o, k * k, cp), device="cuda", dtype=torch.bfloat16)
    taps[:, :, :c] = w.permute(0, 2, 3, 1).reshape(o, k * k, c)
    xs = x.view(n, h // 2, 2, h // 2, 2, c).permute(0, 2, 4, 1, 3, 5).contiguous()     # [n, py, px, y', x', c]
    got = ops.conv2d_nhwc_split(xs, taps.view(o, -1), b, k, pad, relu=True, out_dtype=torch.float32)
    want = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=2, padding=pad))
    assert got.shape == (n, h // 2, h // 2, o)
    assert float((got.permute(0, 3, 1, 2).double() - want).abs().max()) <= 2e-5 * float(want.abs().max())
    if 128 % ((h // 2) ** 2) == 0:
        plain = ops.conv2d_nhwc(x, taps.view(o, -1), b, k, 2, pad, relu=True, out_dtype=torch.float32)
        assert torch.equal(got, plain)


def test_pair_conv1_parity_planes_are_a_permutation_of_the_nhwc_rows(ops):
    g = torch.Generator(device="cuda").manual_seed(4)
    n, oh, ow, c = 9, 16, 16, 96
    maps = torch.randn((n, oh * ow, 2 * c), device="cuda", generator=g)
    ixs = torch.randint(0, n, (50,), device="cuda", generator=g)
    ixo = torch.randint(0, n, (50,), device="cuda", generator=g)
    bias = torch.randn((c,), device="cuda", generator=g)
    plain = ops.pair_conv1_bf16(maps, ixs, ixo, bias).view(50, oh, ow, c)
    split = ops.pair_conv1_bf16(maps, ixs, ixo, bias, split_hw=(oh, ow))
    assert split.shape == (50, 2, 2, oh // 2, ow // 2, c)
    assert torch.equal(split, plain.view(50, oh // 2, 2, ow // 2, 2, c).permute(0, 2, 4, 1, 3, 5).contiguous())


@pytest.mark.parametrize("shape", [(1, 1024, 38, 63), (2, 64, 63, 38)])   # config-3 frame; portrait map (pitch 41)
def test_roi_pool_rows_bf16_planes_equal_fp32_planes(ops, shape):
    """The bf16-plane kernel (two channels per word, max.bf16x2) against the fp32-plane kernel rounded at the end and the
    exact RoIPool op: rounding to bf16 is monotone, so all three agree bit for bit -- union boxes of every size, tiny
    boxes, boxes over the border and stray frame indices included."""
    import os
    from i2vsgg_b200._lib import ARGMAX_PLANE
    B, C, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(C + H)
    feat = torch.randn(shape, device="cuda", generator=g) * 3
    rng = np.random.default_rng(H)
    N = 700
    iw, ih = W * 16.0, H * 16.0
    x1 = rng.uniform(-40, iw - 20, N); y1 = rng.uniform(-40, ih - 20, N)
    w = rng.choice([4.0, 30.0, 90.0, 250.0, 600.0, 1100.0], N); h = rng.choice([4.0, 30.0, 90.0, 250.0, 600.0], N)
    r = np.stack([rng.integers(0, B, N), x1, y1, np.minimum(x1 + w, iw + 30), np.minimum(y1 + h, ih + 30)], 1).astype(np.float32)
    r[5, 0], r[6, 0] = -1, B + 3
    rois = torch.from_numpy(r).cuda()
    got = ops.roi_pool_rows(feat, rois, 7, 7, 1 / 16)
    os.environ["I2V_POOL_F32_PLANES"] = "1"
    try:
        ref = ops.roi_pool_rows(feat, rois, 7, 7, 1 / 16)
    finally:
        del os.environ["I2V_POOL_F32_PLANES"]
    assert torch.equal(got, ref)
    want, _ = ops.roi_pool_forward(feat, rois, 7, 7, 1 / 16, ARGMAX_PLANE)
    assert torch.equal(got, want.reshape(N, -1).bfloat16())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_roi_pool_rows_with_rois_far_larger_than_the_map(ops, dtype):
    """RoIs reaching far past the map are legal (the start / end rows are clamped per bin, roi_pooling_kernel.cu:60-66), and
    their bins can be taller than the 12 rows the row sweep is specialised for (rs = -38, re = 76 on a 38-row map: bin 3
    covers rows 11..28): those take the generic loop and must equal the exact RoIPool op."""
    from i2vsgg_b200._lib import ARGMAX_PLANE
    B, C, H, W = 1, 64, 38, 63
    g = torch.Generator(device="cuda").manual_seed(77)
    feat = torch.randn((B, C, H, W), device="cuda", generator=g) * 2
    r = np.array([[0, -600, -608, 1600, 1216],       # rs = -38, re = 76 in rows
                  [0, -3000, -2000, 4000, 2600],
                  [0, 100, -900, 400, 1500],          # tall and narrow
                  [0, -2000, 100, 3000, 300],         # wide and low
                  [0, 0, 0, 999, 599]], np.float32)
    rois = torch.from_numpy(r).cuda()
    want, _ = ops.roi_pool_forward(feat, rois, 7, 7, 1 / 16, ARGMAX_PLANE)
    got = ops.roi_pool_rows(feat, rois, 7, 7, 1 / 16, dtype=dtype)
    assert torch.equal(got, want.reshape(len(r), -1).to(dtype))
