#!/usr/bin/env python
"""conv_lo's second layer (96 -> 128 channels, 5x5, stride 2, 16x16 -> 8x8) over the 16128 pairs of a frame group, as the
implicit GEMM of i2v_conv2d_nhwc_forward (strided box) and of i2v_conv2d_nhwc_split_forward (parity planes)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from i2vsgg_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16128
x = torch.randn((n, 16, 16, 96), device="cuda").bfloat16()
w = (torch.randn((128, 25 * 128), device="cuda") * 0.02).bfloat16()
b = torch.zeros(128, device="cuda")
xs = x.view(n, 8, 2, 8, 2, 96).permute(0, 2, 4, 1, 3, 5).contiguous()
for name, fn in (("strided 4-D map", lambda: ops.conv2d_nhwc(x, w, b, 5, 2, 2)),
                 ("parity planes, dense 5-D map", lambda: ops.conv2d_nhwc_split(xs, w, b, 5, 2))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 10
    print(json.dumps({"input": name, "pairs": n, "ms": ms, "tflops": 2 * n * 64 * 128 * 2400 / ms / 1e9}))
