"""CPU-side tests: the C-ABI library loads and exports every symbol include/d2b200.h declares, and the host
wrappers keep the reference's surface (names, repr strings, exception types).  No compute calls: there is no GPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from detectron2_b200 import _C

    lib = _C.lib()
    header = open(os.path.join(ROOT, "include", "d2b200.h")).read()
    declared = set(re.findall(r"\b(d2b_[a-z0-9_]+)\s*\(", header)) - {"d2b_dcn_params"}
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_C.EXPORTED) == declared
    assert lib.d2b_abi_version() == _C.ABI_VERSION == 4
    assert lib.d2b_arch() == b"sm_100a"
    assert _C.get_cuda_version().startswith("CUDA 12")


def test_workspace_queries_run_on_host():
    from detectron2_b200 import _C

    small, big = _C.lib().d2b_nms_workspace_bytes(1000, 0, 0), _C.lib().d2b_nms_workspace_bytes(10000, 0, 0)
    assert 0 < small < big
    assert _C.lib().d2b_nms_workspace_bytes(10000, 1, 0) > big  # 5 floats per rotated box
    assert _C.lib().d2b_nms_workspace_bytes(10000, 0, 1000) < big / 4  # bounded categories: bitmask linear in M


def test_surface_matches_reference_names():
    import detectron2_b200.layers as L

    for name in ["ROIAlign", "roi_align", "ROIAlignRotated", "roi_align_rotated", "DeformConv", "ModulatedDeformConv",
                 "batched_nms", "nms", "batched_nms_rotated", "nms_rotated", "paste_masks_in_image",
                 "pairwise_iou_rotated"]:
        assert hasattr(L, name), name
    for op in ["nms_rotated", "box_iou_rotated", "roi_align_rotated_forward", "roi_align_rotated_backward"]:
        assert hasattr(torch.ops.detectron2, op)


def test_repr_strings():  # /root/reference/tests/layers/test_deformable.py:157-171
    import detectron2_b200.layers as L

    assert repr(L.DeformConv(3, 10, kernel_size=3, padding=1, deformable_groups=2)) == (
        "DeformConv(in_channels=3, out_channels=10, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1), "
        "dilation=(1, 1), groups=1, deformable_groups=2, bias=False)")
    assert repr(L.ModulatedDeformConv(3, 10, kernel_size=3, padding=1, deformable_groups=2)) == (
        "ModulatedDeformConv(in_channels=3, out_channels=10, kernel_size=(3, 3), stride=1, padding=1, dilation=1, "
        "groups=1, deformable_groups=2, bias=True)")
    assert "aligned=True" in repr(L.ROIAlign((7, 7), 0.25, 0))
    m = L.ModulatedDeformConv(4, 8, 3)
    assert m.weight.shape == (8, 4, 3, 3) and m.bias.shape == (8,) and (m.bias == 0).all()


def test_no_cpu_fallback():
    import detectron2_b200.layers as L

    with pytest.raises(NotImplementedError):
        L.nms(torch.rand(4, 4), torch.rand(4), 0.5)
    with pytest.raises(NotImplementedError):
        L.ROIAlign((7, 7), 1.0, 0)(torch.rand(1, 1, 8, 8), torch.tensor([[0.0, 1, 1, 4, 4]]))
    with pytest.raises(NotImplementedError):
        L.DeformConv(1, 1, 3, padding=1)(torch.rand(1, 1, 5, 5), torch.zeros(1, 18, 5, 5))
    with pytest.raises(ValueError):
        L.deform_conv(torch.rand(1, 5, 5), torch.zeros(1, 18, 5, 5), torch.rand(1, 1, 3, 3))
    with pytest.raises(AssertionError):
        L.ROIAlign((7, 7), 1.0, 0)(torch.rand(1, 1, 8, 8), torch.rand(3, 4))


def test_empty_inputs_host_side():
    import detectron2_b200.layers as L

    assert L.batched_nms_rotated(torch.zeros(0, 5), torch.zeros(0), torch.zeros(0, dtype=torch.int64), 0.5).shape == (0,)
    assert L.batched_nms(torch.zeros(0, 4), torch.zeros(0), torch.zeros(0, dtype=torch.int64), 0.5).shape == (0,)
    out = L.paste_masks_in_image(torch.zeros(0, 28, 28), torch.zeros(0, 4), (10, 12))
    assert out.shape == (0, 10, 12) and out.dtype == torch.uint8
    y = L.DeformConv(2, 4, 3, padding=1)(torch.zeros(0, 2, 8, 8), torch.zeros(0, 18, 8, 8))
    assert y.shape == (0, 4, 8, 8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "detectron2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "import torchvision" not in src and "from torchvision" not in src, f


def test_fake_kernels_trace_shapes():
    """Every custom op has a fake (meta) kernel, so the ops trace / compile without touching the GPU library."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    from detectron2_b200 import ops

    with FakeTensorMode():
        x = torch.empty(2, 16, 20, 30, device="cuda")
        rois = torch.empty(7, 5, device="cuda")
        assert ops.roi_align_op(x, rois, 0.25, 7, 7, 0, True).shape == (7, 16, 7, 7)
        assert ops.roi_align_rotated_op(x, torch.empty(7, 6, device="cuda"), 0.25, 5, 4, 2).shape == (7, 16, 5, 4)
        feats = [torch.empty(2, 16, 40 // 2 ** i, 60 // 2 ** i, device="cuda") for i in range(4)]
        y = ops.roi_pooler_op(feats, rois, [1 / 4, 1 / 8, 1 / 16, 1 / 32], 14, 14, 0, True, 2, 5, 4, 224.0)
        assert y.shape == (7, 16, 14, 14)
        w = torch.empty(24, 8, 3, 3, device="cuda")
        off = torch.empty(2, 18, 10, 15, device="cuda")
        out = ops.deform_conv_op(x, off, None, w, None, [2, 2], [1, 1], [1, 1], 2, 1, -1)
        assert out.shape == (2, 24, 10, 15)
        assert ops.box_iou_rotated_op(torch.empty(5, 5, device="cuda"), torch.empty(9, 5, device="cuda")).shape == (5, 9)
        pm = ops.paste_masks_op(torch.empty(3, 28, 28, device="cuda"), torch.empty(3, 4, device="cuda"), 40, 50, 0.5)
        assert pm.shape == (3, 40, 50) and pm.dtype == torch.bool  # threshold >= 0: bool output like the reference


def test_pooler_layout_policy_host_logic(monkeypatch):
    """Which RoIAlign kernel a call gets (detectron2_b200/ops.py): pure host logic, checked on CPU tensors."""
    from detectron2_b200 import ops

    nchw = [torch.zeros(1, 256, 200 // 2 ** i, 336 // 2 ** i) for i in range(4)]
    cl = [t.contiguous(memory_format=torch.channels_last) for t in nchw]
    assert all(ops._is_channels_last(t) for t in cl) and not any(ops._is_channels_last(t) for t in nchw)
    # C == 1 or H*W == 1: both layouts coincide, treated as plain NCHW
    assert not ops._is_channels_last(torch.zeros(2, 1, 5, 5).contiguous(memory_format=torch.channels_last))
    monkeypatch.setattr(ops, "POOLER_LAYOUT", "auto")
    assert ops._pick_layout(cl, 1) == "cl"                        # channels_last is consumed in place
    assert ops._pick_layout(nchw, 1000 * 256 * 49) == "xpose"     # box head: layout change + NHWC kernel pays
    assert ops._pick_layout(nchw, 100 * 256 * 196) == "nchw"      # mask head alone: NCHW kernel
    assert ops._pick_layout(nchw, 10 * 256 * 49) == "nchw"
    assert ops._pick_layout([nchw[0], cl[1]], 10) == "nchw"        # mixed pyramid: falls back to the NCHW kernel
    odd = [torch.zeros(1, 6, 8, 8).contiguous(memory_format=torch.channels_last)]
    assert ops._pick_layout(odd, 10 ** 9) == "nchw"                # C % 4 != 0: NHWC kernel not applicable
    monkeypatch.setattr(ops, "POOLER_LAYOUT", "nchw")
    assert ops._pick_layout(cl, 1) == "nchw"
    monkeypatch.setattr(ops, "POOLER_LAYOUT", "nhwc")
    assert ops._pick_layout(nchw, 1) == "xpose"
    with pytest.raises(NotImplementedError):                       # no CPU fallback for the layout change either
        ops.pyramid_to_channels_last(nchw)


def test_roi_pooler_no_images():  # /root/reference/tests/modeling/test_roi_pooler.py:107-115
    from detectron2_b200.poolers import ROIPooler

    feature = torch.rand(0, 32, 32, 32) - 0.5
    pooler = ROIPooler(output_size=14, scales=(1.0,), sampling_ratio=0.0, pooler_type="ROIAlignV2")
    assert pooler.forward([feature], []).shape == (0, 32, 14, 14)
    with pytest.raises(ValueError):
        ROIPooler(output_size=7, scales=(0.25,), sampling_ratio=0, pooler_type="ROIPool")
    with pytest.raises(AssertionError):  # scales that do not form a pyramid (poolers.py:186-190)
        ROIPooler(output_size=7, scales=(0.25, 0.0625), sampling_ratio=0, pooler_type="ROIAlignV2")


class _RotNms(torch.nn.Module):  # /root/reference/tests/layers/test_nms_rotated.py:153-168
    def forward(self, boxes, scores, threshold: float):
        import detectron2_b200.layers as L

        return L.nms_rotated(boxes, scores, threshold)


def test_wrappers_are_scriptable():
    """The wrappers the reference scripts in its own tests (tests/layers/test_nms.py:16-29, test_nms_rotated.py:153-168,
    test_mask_ops.py:156-165) compile with torch.jit.script: their bodies are dispatcher ops only.  (Scripted == eager
    is checked on the GPU in tests/test_gpu_parity.py.)"""
    import detectron2_b200.layers as L

    for fn in (L.batched_nms, L.nms, L.batched_nms_rotated, L.paste_masks_in_image):
        f = fn.__original_fn if hasattr(fn, "__original_fn") else fn  # script_if_tracing wrapper
        assert torch.jit.script(f) is not None
    assert "detectron2::nms_rotated" in str(torch.jit.script(_nms_rotated_fn).graph)


def _nms_rotated_fn(boxes: torch.Tensor, scores: torch.Tensor, threshold: float) -> torch.Tensor:
    return torch.ops.detectron2.nms_rotated(boxes, scores, threshold)


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Status codes of the round-2 entry points on invalid arguments (checked before anything is launched): negative =
    d2b error (include/d2b200.h), what the Python host turns into RuntimeError."""
    import ctypes as C

    from detectron2_b200 import _C

    lib = _C.lib()
    EINVAL = -1
    # Fast R-CNN candidates: too many images for one call, class-specific boxes that do not match the class count, no row table
    rs = (C.c_int * 3)(0, 4, 8)
    assert lib.d2b_frcnn_prepare(None, None, rs, _C.MAX_IMAGES + 1, 80, 80, None, 0.05, 16, *([None] * 8), None) == EINVAL
    assert lib.d2b_frcnn_prepare(None, None, rs, 2, 80, 3, None, 0.05, 16, *([None] * 8), None) == EINVAL
    assert lib.d2b_frcnn_prepare(None, None, None, 2, 80, 80, None, 0.05, 16, *([None] * 8), None) == EINVAL
    assert lib.d2b_frcnn_prepare(None, None, rs, 0, 80, 80, None, 0.05, 16, *([None] * 8), None) == 0  # no images: nothing to do
    # dense head: level count out of range, missing weights
    lv = _C.DenseLevels()
    lv.num_levels = 0
    w = (C.c_float * 4)(1, 1, 1, 1)
    assert lib.d2b_dense_prepare(C.byref(lv), 2, 80, w, 4.135, *([None] * 6), None) == EINVAL
    lv.num_levels = 1
    assert lib.d2b_dense_prepare(C.byref(lv), 2, 80, None, 4.135, *([None] * 6), None) == EINVAL
    assert lib.d2b_dense_prepare(C.byref(lv), 0, 80, w, 4.135, *([None] * 6), None) == 0
    # dtype codes of the half-precision variants
    P = _C.Pyramid()
    P.num_levels = 1
    P.H[0], P.W[0] = 8, 8
    dst = (C.c_void_p * 1)(None)
    assert lib.d2b_pyramid_nchw_to_nhwc_t(C.byref(P), 1, 4, dst, 7, None) == EINVAL
    assert lib.d2b_pyramid_nhwc_to_nchw_t(C.byref(P), 1, 4, dst, -1, None) == EINVAL
    assert lib.d2b_roi_pooler_forward_nhwc_t(C.byref(P), 1, 4, None, 3, 7, 7, 0, 1, None, 5, None) == EINVAL
    assert lib.d2b_roi_pooler_backward_nhwc_t(C.byref(P), 1, 4, None, 9, None, 3, 7, 7, 0, 1, None) == EINVAL
    assert _C.DTYPE_CODE == {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
    # bit-packed paste: boolean output only
    assert lib.d2b_paste_masks_packed(None, None, 3, 28, 10, 10, -1.0, None, None) == EINVAL
    assert lib.d2b_paste_masks_packed(None, None, 0, 28, 10, 10, 0.5, None, None) == 0


def test_post_processing_dispatch_and_pyramid_struct():
    """CPU tensors take the torch-op restatements (pinned to the real reference functions in test_host_logic_cpu.py); the ABI
    struct carries the level_rois field of ABI v4."""
    from detectron2_b200 import _C, dense_inference, fast_rcnn_inference

    assert [n for n, _ in _C.Pyramid._fields_][-1] == "level_rois"
    assert callable(fast_rcnn_inference.fast_rcnn_inference_fixed) and callable(dense_inference.dense_detector_inference_fixed)
    with pytest.raises(NotImplementedError):  # the fixed-capacity forms are CUDA only
        fast_rcnn_inference.fast_rcnn_inference_fixed([torch.zeros(2, 4)], [torch.zeros(2, 3)], [(10, 10)], 0.05, 0.5, 10)
    with pytest.raises(NotImplementedError):
        dense_inference.dense_detector_inference_fixed([torch.zeros(4, 4)], [torch.zeros(1, 4, 2)], [torch.zeros(1, 4, 4)], 1,
                                                       0.05, 10, 0.5, 10)


def test_unpack_mask_bits_host_side():
    """The receiving side of the bit-packed paste: plain torch ops, CPU tensors (bit b of word w = pixel 32 w + b)."""
    import detectron2_b200.layers as L

    g = torch.Generator().manual_seed(0)
    ref = torch.rand(3, 5, 70, generator=g) > 0.5
    words = torch.zeros(3, 5, 3, dtype=torch.int64)
    for x in range(70):
        words[..., x // 32] |= ref[..., x].to(torch.int64) << (x % 32)
    packed = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)  # two's-complement int32 words
    assert torch.equal(L.unpack_mask_bits(packed, 70), ref)
    assert L.paste_masks_in_image_packed(torch.zeros(0, 28, 28), torch.zeros(0, 4), (10, 40)).shape == (0, 10, 2)


def test_ctypes_signatures_match_the_header_arity():
    """Every prototype of include/d2b200.h against the ctypes binding: same number of parameters (a missing / extra argument in
    the binding would shift every following pointer)."""
    from detectron2_b200 import _C

    _C.lib()
    header = open(os.path.join(ROOT, "include", "d2b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(d2b_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(protos) >= 30
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(_C.EXPORTED[name][1]) == n, (name, n, len(_C.EXPORTED[name][1]))
