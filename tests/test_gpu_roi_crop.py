"""GPU parity tests of roi_crop: the C ABI against the CPU oracle (itself bit-identical to the
reference's roi_crop.c, tests/test_oracle.py) and against the reference's own roi_crop_cuda_kernel.cu compiled unmodified
for sm_100a (oracle/_ref/libref_cuda.so, when it travelled with the snapshot)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(seed, B, C, H, W, per, oh, ow):
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((B, C, H, W)).astype(np.float32)
    N = B * per
    # affine-grid-like boxes plus jitter; some samples well outside [-1, 1] (corners partly or wholly off the map)
    cy, cx = rng.uniform(-0.9, 0.9, (2, N, 1, 1))
    sy, sx = rng.uniform(0.05, 0.8, (2, N, 1, 1))
    ys = cy + sy * np.linspace(-1, 1, oh)[None, :, None]
    xs = cx + sx * np.linspace(-1, 1, ow)[None, None, :]
    grids = np.stack([np.broadcast_to(ys, (N, oh, ow)), np.broadcast_to(xs, (N, oh, ow))], 3).astype(np.float32)
    grids[0, 0, :, :] = [-1.0, 1.0]                                  # exactly on the border
    grids[-1, -1, :, 0] = 1.7                                        # off the map
    grad = rng.standard_normal((N, C, oh, ow)).astype(np.float32)
    return feat, grids, grad


@pytest.mark.parametrize("shape", [(2, 5, 11, 13, 3, 7, 7), (1, 64, 38, 63, 20, 7, 7), (3, 16, 20, 31, 1, 4, 9)])
def test_roi_crop_against_the_oracle(shape):
    from i2vsgg_b200 import ops
    from oracle import oracle
    feat, grids, grad = _case(7, *shape)
    f, g, go = (torch.from_numpy(a).cuda() for a in (feat, grids, grad))
    out = ops.roi_crop_forward(f, g)
    assert np.array_equal(out.cpu().numpy(), oracle.roi_crop_forward(feat, grids))          # same fp32 operations, bit for bit
    gin = ops.roi_crop_backward(go, g, feat.shape)
    want = oracle.roi_crop_backward(grad, grids, feat.shape)
    np.testing.assert_allclose(gin.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * float(np.abs(want).max()))   # atomic order


def test_roi_crop_reference_kernels_pin_us():
    from i2vsgg_b200 import ops
    from oracle import ref
    if not ref.have_cuda_ref():
        pytest.skip("oracle/_ref/libref_cuda.so did not travel")
    feat, grids, grad = _case(9, 2, 32, 38, 63, 8, 7, 7)
    f, g, go = (torch.from_numpy(a).cuda() for a in (feat, grids, grad))
    want = ref.cuda_roi_crop_forward(f, g)
    got = ops.roi_crop_forward(f, g)
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 1e-5 * scale           # nvcc contracts the reference's products into FMAs
    wgi, wgg = ref.cuda_roi_crop_backward(f, g, go)
    gin = ops.roi_crop_backward(go, g, feat.shape)
    assert float((gin - wgi).abs().max()) <= 1e-5 * float(wgi.abs().max())
    assert float(wgg.abs().max()) == 0.0                              # the reference never writes the grid gradient


def test_roi_crop_module_and_legacy_launchers():
    import i2vsgg_b200
    from i2vsgg_b200 import _lib, ops
    i2vsgg_b200.install_as_model()
    from model.roi_crop.modules.roi_crop import _RoICrop
    feat, grids, grad = _case(11, 2, 8, 15, 21, 4, 5, 6)
    f = torch.from_numpy(feat).cuda().requires_grad_(True)
    g = torch.from_numpy(grids).cuda().requires_grad_(True)
    out = _RoICrop()(f, g)
    out.backward(torch.from_numpy(grad).cuda())
    assert torch.equal(out.detach(), ops.roi_crop_forward(f.detach(), g.detach()))
    assert torch.equal(f.grad, ops.roi_crop_backward(torch.from_numpy(grad).cuda(), g.detach(), feat.shape)) or \
        float((f.grad - ops.roi_crop_backward(torch.from_numpy(grad).cuda(), g.detach(), feat.shape)).abs().max()) < 1e-5
    assert float(g.grad.abs().max()) == 0.0
    # the reference's launcher with explicit strides: a channel-last view of the same features gives the same rows
    lib = _lib.load()
    B, C, H, W = feat.shape
    N, oh, ow, _ = grids.shape
    nhwc = f.detach().permute(0, 2, 3, 1).contiguous()                # element (b,c,y,x) at b*HWC + y*WC + x*C + c
    o2 = torch.zeros((N, C, oh, ow), device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.BilinearSamplerBHWD_updateOutput_cuda_kernel(C, ow, oh, N, C, H, W, B, P(nhwc), H * W * C, 1, W * C, C, P(g.detach()),
                                                          oh * ow * 2, 1, ow * 2, 2, P(o2), C * oh * ow, oh * ow, ow, 1, s)
    assert rc == 1 and torch.equal(o2, out.detach())
