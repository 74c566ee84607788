"""Writes tests/golden/rpn_golden.npz by EXECUTING the reference's RPN-head softmax (lib/model/rpn/rpn.py:66-68, with
`_RPN.reshape` taken from the unmodified source; oracle/ref.py) on CPU tensors.
    python tests/golden/make_rpn_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rpn_golden.npz")


def main():
    rng = np.random.default_rng(21)
    g = {}
    for name, shape, scale in (("small", (2, 18, 5, 7), 3.0), ("wide", (1, 18, 12, 20), 12.0), ("a3", (3, 6, 4, 4), 1.0)):
        x = (rng.standard_normal(shape) * scale).astype(np.float32)
        if name == "wide":
            x[0, 0, 0, :4] = [80.0, -80.0, 0.0, 1e-3]        # saturated pairs
            x[0, 9, 0, :4] = [-80.0, 80.0, 0.0, -1e-3]
        g[f"{name}_score"] = x
        g[f"{name}_prob"] = ref.py_rpn_cls_prob(x)
        print(name, shape, float(g[f"{name}_prob"].min()), float(g[f"{name}_prob"].max()))
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
