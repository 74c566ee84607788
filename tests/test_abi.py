"""CPU tests of the boundary: the shared library loads without a GPU and exports every entry point that
include/i2vsgg_b200.h declares; the ctypes table mirrors the header; the drop-in module tree imports under the
reference's dotted names.  No compute call is made here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "i2vsgg_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return [n for n in names if n.startswith("i2v_") or n.endswith("Laucher") or n == "nms_cuda_compute" or
            n.startswith("BilinearSamplerBHWD_")]


@pytest.fixture(scope="module")
def lib():
    from i2vsgg_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "i2vsgg_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    return _lib.load()


def test_header_declares_the_reference_launchers():
    names = _declared()
    for n in ("ROIAlignForwardLaucher", "ROIAlignBackwardLaucher", "ROIPoolForwardLaucher", "ROIPoolBackwardLaucher",
              "nms_cuda_compute", "BilinearSamplerBHWD_updateOutput_cuda_kernel",
              "BilinearSamplerBHWD_updateGradInput_cuda_kernel"):
        assert n in names
    assert len(names) == len(set(names)) and len(names) >= 24


def test_library_exports_every_declared_symbol(lib):
    for n in _declared():
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_table_mirrors_header(lib):
    from i2vsgg_b200 import _lib
    assert set(_lib.SIGNATURES) == set(_declared())
    assert lib.i2v_abi_version() == 1
    assert lib.i2v_last_error() == b""
    # size queries are pure host arithmetic
    assert lib.i2v_roi_align_workspace_bytes(32, 9600) > 9600 * 272
    assert lib.i2v_nms_workspace_bytes(1, 12000) >= 12000 * 20
    assert lib.i2v_proposal_workspace_bytes(1, 9, 38, 63, 6000) > 21546 * 4 * 6
    assert lib.i2v_triplet_topk_workspace_bytes(4032, 132) >= 4032 * 132 * 4


def test_header_compiles_as_c():
    code = '#include "i2vsgg_b200.h"\nint main(void){return i2v_abi_version == 0;}\n'
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=code.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_invalid_arguments_return_status_not_exit(lib):
    # validated before any CUDA call, so this runs without a device (the reference calls exit(-1) on errors)
    rc = lib.i2v_roi_align_forward(None, None, None, 1, 16, 38, 63, 10, 7, 7, ctypes.c_float(1 / 16), 9, 0, None, 0, None)
    assert rc == 1 and b"pool_mode" in lib.i2v_last_error()
    rc = lib.i2v_triplet_topk(None, None, None, None, None, None, 10, 10, 5000, None, None, None, 0, None)
    assert rc == 1


def test_model_tree_imports_under_reference_names():
    code = ("import i2vsgg_b200; i2vsgg_b200.install_as_model();"
            "from model.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax;"
            "from model.roi_align.functions.roi_align import RoIAlignFunction;"
            "from model.roi_pooling.modules.roi_pool import _RoIPooling;"
            "from model.roi_pooling.functions.roi_pool import RoIPoolFunction;"
            "from model.nms.nms_wrapper import nms; from model.nms.nms_gpu import nms_gpu;"
            "from model.rpn.proposal_layer import _ProposalLayer;"
            "from model.roi_layers import ROIAlign, ROIPool, nms, roi_align, roi_pool;"
            "from model.utils.config import cfg;"
            "l=_ProposalLayer(cfg.FEAT_STRIDE[0], cfg.ANCHOR_SCALES, cfg.ANCHOR_RATIOS);"
            "assert l._num_anchors == 9 and cfg['TEST'].RPN_POST_NMS_TOP_N == 300;"
            "m=RoIAlignAvg(7,7,1/16.); assert m.aligned_height == 7")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_product_does_not_touch_the_oracle():
    # only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may use oracle/
    for base, _, files in os.walk(os.path.join(ROOT, "i2vsgg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_cpu_tensors_are_rejected():
    import torch
    from i2vsgg_b200 import ops, _lib
    with pytest.raises(_lib.I2VError):
        ops.roi_align_forward(torch.zeros(1, 16, 8, 8), torch.zeros(1, 5), 7, 7, 1 / 16)
    with pytest.raises(_lib.I2VError):
        ops.nms_dets(torch.zeros(4, 5), 0.7)
