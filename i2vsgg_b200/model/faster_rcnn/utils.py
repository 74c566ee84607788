"""`Conv2d` and `FC` of lib/model/faster_rcnn/utils.py:32-60: parameter containers with the reference's names
(`.conv`, `.bn`, `.fc`), so a reference checkpoint loads with `load_state_dict`.  Their `forward` runs on the
sm_100a kernels of this package: bf16 tensor-core operands (or tf32 for fp32 inputs), fp32 accumulation.

The derived copies of the weights (bf16 / padded / folded BatchNorm) are rebuilt whenever a parameter's version
counter, data pointer or device changes: an optimizer step, `load_state_dict` or `.to()` after the first forward can
not leave a stale copy behind."""
from __future__ import annotations

import torch
from torch import nn

from ... import ops


class FC(nn.Module):
    def __init__(self, in_features, out_features, relu=True):
        super().__init__()
        self.fc = nn.Linear(in_features, out_features)
        self.relu = nn.ReLU(inplace=True) if relu else None
        self._w = None
        self._key = None

    def _state(self):
        w, b = self.fc.weight, self.fc.bias
        return (w._version, w.data_ptr(), b._version, b.data_ptr(), str(w.device))

    def prepare(self):
        """(Re)builds the bf16 copy of the weight the tensor cores read; the row pitch is padded to 16 bytes."""
        self._key = self._state()
        self._w32 = None
        w = self.fc.weight.detach()
        k = w.size(1)
        kp = (k + 7) // 8 * 8
        buf = torch.zeros((w.size(0), kp), dtype=torch.bfloat16, device=w.device)
        ops.cast_bf16(w.float().contiguous(), buf[:, :k])
        self._w = buf[:, :k]
        self._b = self.fc.bias.detach().float().contiguous()
        return self

    def forward(self, x, out=None, out_dtype=None, keep_mask=None, keep_scale=1.0):
        """x [M, in_features] (rows may be strided) -> [M, out_features]; `out` may be a column slice.  bf16 rows go
        through the bf16 copy of the weight (tcgen05 kind::f16), fp32 rows through the fp32 weight itself (kind::tf32).
        `keep_mask` [M, out_features] uint8: dropout behind the activation, applied in the kernel's epilogue."""
        if self._w is None or self._key != self._state():
            self.prepare()
        if x.dtype == torch.float32:
            if self._w32 is None:
                w = self.fc.weight.detach().float().clone().contiguous()     # a copy: the parameter itself stays fp32
                if (w.size(1) * 4) % 16:                       # TMA: 16-byte row pitch
                    kp = (w.size(1) + 3) // 4 * 4
                    buf = torch.zeros((w.size(0), kp), dtype=torch.float32, device=w.device)
                    buf[:, : w.size(1)] = w
                    w = buf[:, : self.fc.in_features]
                self._w32 = ops.round_tf32(w, w)           # the tensor core truncates: round the operands first
            x = ops.round_tf32(x)                          # out of place: the caller's activations are left alone
            return ops.linear(x, self._w32, self._b, relu=self.relu is not None, out=out,
                              out_dtype=out_dtype or torch.float32, keep_mask=keep_mask, keep_scale=keep_scale)
        return ops.linear(x, self._w, self._b, relu=self.relu is not None, out=out, out_dtype=out_dtype or torch.bfloat16,
                          keep_mask=keep_mask, keep_scale=keep_scale)


class Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, relu=True, same_padding=False, bn=False):
        super().__init__()
        padding = int((kernel_size - 1) / 2) if same_padding else 0
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding=padding)
        self.bn = nn.BatchNorm2d(out_channels, eps=0.001, momentum=0, affine=True) if bn else None
        self.relu = nn.ReLU(inplace=True) if relu else None
        self._w = None
        self._key = None

    def _state(self):
        ps = [self.conv.weight, self.conv.bias]
        if self.bn is not None:
            ps += [self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        return tuple((p._version, p.data_ptr()) for p in ps) + (str(self.conv.weight.device), self.training)

    def _fresh(self):
        if self._w is None or self._key != self._state():
            self.prepare()

    def prepare(self):
        """Weight as [out, (ky, kx, c)] bf16 rows (the patch order of `ops.im2col_bf16`), eval-mode BatchNorm folded
        into weight and bias (utils.py:43-44 in eval mode is an affine map per output channel)."""
        self._key = self._state()
        self._w32 = None
        w = self.conv.weight.detach().float()
        b = self.conv.bias.detach().float()
        if self.bn is not None:
            s = self.bn.weight.detach().float() / torch.sqrt(self.bn.running_var.float() + self.bn.eps)
            w = w * s[:, None, None, None]
            b = (b - self.bn.running_mean.float()) * s + self.bn.bias.detach().float()
        o, c, kh, kw = w.shape
        k = kh * kw * c
        kp = (k + 7) // 8 * 8
        buf = torch.zeros((o, kp), dtype=torch.bfloat16, device=w.device)
        ops.cast_bf16(w.permute(0, 2, 3, 1).reshape(o, k).contiguous(), buf[:, :k])
        self._w, self._kp = buf, kp
        self._b = b.contiguous()
        self._w_pair = None
        # the implicit-GEMM layout: every tap's channels padded to a multiple of 64 (one or more whole k-blocks per tap)
        cp = (c + 63) // 64 * 64
        taps = torch.zeros((o, kh * kw, cp), dtype=torch.bfloat16, device=w.device)
        taps[:, :, :c] = buf[:, :k].view(o, kh * kw, c)
        self._w_taps = taps.view(o, kh * kw * cp)
        return self

    def takes_split(self, h, w, c):
        """Whether `forward(x, 'nhwc_split')` applies to an [h, w, c] map: the stride-2 implicit GEMM on parity planes."""
        kh, st, pd = self.conv.kernel_size[0], self.conv.stride[0], self.conv.padding[0]
        return (st == 2 and kh % 2 == 1 and pd == (kh - 1) // 2 and h % 2 == 0 and w % 2 == 0 and c % 8 == 0
                and 128 % ((h // 2) * (w // 2)) == 0 and self.conv.out_channels <= 128 and c == self.conv.in_channels)

    def forward_pairs(self, obj_x, ixs, ixo, split: bool = False):
        """This layer applied to P two-channel inputs whose channels are `obj_x[ixs[p]]` and `obj_x[ixo[p]]`
        (obj_x [N,H,W]): the convolution is linear in its input channels, so each OBJECT map is convolved once with
        either half of the kernel and a pair is the sum of two rows plus the bias.  -> NHWC bf16 [P,OH,OW,out], or with
        `split` the parity-split layout [P,2,2,OH/2,OW/2,out] that the next layer's `forward(x, 'nhwc_split')` reads."""
        self._fresh()
        o, kh = self.conv.out_channels, self.conv.kernel_size[0]
        if self.conv.in_channels != 2:
            raise ValueError("forward_pairs needs a two-channel convolution")
        if getattr(self, "_w_pair", None) is None or self._w_pair.device != obj_x.device:
            taps = kh * kh
            kp = (taps + 7) // 8 * 8
            w = self._w[:, : 2 * taps].float().view(o, taps, 2)            # (ky, kx, c) order of prepare()
            wp = torch.zeros((2 * o, kp), dtype=torch.bfloat16, device=obj_x.device)
            wp[:o, :taps] = w[:, :, 0].bfloat16()                          # subject half (exact: already bf16 values)
            wp[o:, :taps] = w[:, :, 1].bfloat16()                          # object half
            self._w_pair, self._kp_pair = wp, kp
        n, h, wd = obj_x.shape
        patches, (_, oh, ow) = ops.im2col_bf16(obj_x.reshape(n, 1, h, wd).contiguous(), kh, self.conv.stride[0],
                                               self.conv.padding[0], "nchw", ld=self._kp_pair)
        maps = ops.linear(patches, self._w_pair, None, relu=False, out_dtype=torch.float32)       # [N*OH*OW, 2*out]
        if split:
            return ops.pair_conv1_bf16(maps.view(n, oh * ow, 2 * o), ixs, ixo, self._b, relu=self.relu is not None,
                                       split_hw=(oh, ow))
        y = ops.pair_conv1_bf16(maps.view(n, oh * ow, 2 * o), ixs, ixo, self._b, relu=self.relu is not None)
        return y.view(ixs.numel(), oh, ow, o)

    def forward_tf32(self, x, layout: str):
        """The layer on fp32 patches and fp32 weights through tcgen05 kind::tf32 (`vrd(..., precision="tf32")`):
        x fp32, [N,C,H,W] or [N,H,W,C] -> NHWC fp32 [N,OH,OW,out]."""
        self._fresh()
        kh = self.conv.kernel_size[0]
        if getattr(self, "_w32", None) is None:
            w = self.conv.weight.detach().float()
            b = self.conv.bias.detach().float()
            if self.bn is not None:
                sc = self.bn.weight.detach().float() / torch.sqrt(self.bn.running_var.float() + self.bn.eps)
                w = w * sc[:, None, None, None]
                b = (b - self.bn.running_mean.float()) * sc + self.bn.bias.detach().float()
            o, c = w.size(0), w.size(1)
            k = kh * kh * c
            kp = (k + 3) // 4 * 4
            buf = torch.zeros((o, kp), dtype=torch.float32, device=w.device)
            buf[:, :k] = w.permute(0, 2, 3, 1).reshape(o, k)
            self._w32, self._kp32, self._b32 = ops.round_tf32(buf, buf), kp, b.contiguous()
        patches, (n, oh, ow) = ops.im2col_bf16(x.float().contiguous(), kh, self.conv.stride[0], self.conv.padding[0], layout,
                                               ld=self._kp32, out_dtype=torch.float32)
        ops.round_tf32(patches, patches)
        y = ops.linear(patches, self._w32, self._b32, relu=self.relu is not None, out_dtype=torch.float32)
        return y.view(n, oh, ow, self.conv.out_channels)

    def forward(self, x, layout: str):
        """x [N,C,H,W] (`layout='nchw'`) or [N,H,W,C] (`'nhwc'`) -> NHWC bf16 [N,OH,OW,out]: im2col rows + one FC launch."""
        self._fresh()
        kh = self.conv.kernel_size[0]
        st, pd = self.conv.stride[0], self.conv.padding[0]
        if layout == "nhwc_split":      # [N,2,2,H/2,W/2,C] parity planes (see `takes_split`)
            return ops.conv2d_nhwc_split(x, self._w_taps, self._b, kh, pd, relu=self.relu is not None)
        if layout == "nhwc" and x.dtype == torch.bfloat16 and x.is_contiguous() and st > 1:
            oh, ow = (x.size(1) + 2 * pd - kh) // st + 1, (x.size(2) + 2 * pd - kh) // st + 1
            if 128 % (oh * ow) == 0 and x.size(3) % 8 == 0 and self.conv.out_channels <= 128:
                # strided layer: implicit GEMM, the 4-D TMA map reads the patches straight out of the activation
                return ops.conv2d_nhwc(x, self._w_taps, self._b, kh, st, pd, relu=self.relu is not None)
        if (layout == "nhwc" and x.dtype == torch.bfloat16 and x.size(1) == kh and x.size(2) == kh
                and self.conv.padding[0] == 0 and self._kp == kh * kh * x.size(3) and x.is_contiguous()):
            # the kernel covers the whole map: the NHWC activation row already is the (ky, kx, c) patch
            patches, (n, oh, ow) = x.view(x.size(0), kh * kh * x.size(3)), (x.size(0), 1, 1)
        else:
            patches, (n, oh, ow) = ops.im2col_bf16(x, kh, self.conv.stride[0], self.conv.padding[0], layout, ld=self._kp)
        y = ops.linear(patches, self._w, self._b, relu=self.relu is not None, out_dtype=torch.bfloat16)
        return y.view(n, oh, ow, self.conv.out_channels)
