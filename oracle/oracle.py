"""ctypes front-end of the CPU oracle (oracle/d2_oracle.c) and loader of the compiled reference.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; detectron2_b200/ never does.

All functions take/return CPU torch tensors (float32 / int64) so tests read like the reference's.
"""
import ctypes as C
import os

import torch

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.ORACLE_SO if os.path.exists(_build.ORACLE_SO) else _build.build_oracle()
        try:
            _lib = C.CDLL(path)
        except OSError:
            _lib = C.CDLL(_build.build_oracle(force=True))
    return _lib


_ref_loaded = None


def load_reference():
    """torch.ops.load_library(oracle/_ref/d2_ref_cpu.so): the reference's own CPU kernels
    (torch.ops.detectron2.{nms_rotated, box_iou_rotated, roi_align_rotated_forward/backward}).
    Returns True when available."""
    global _ref_loaded
    if _ref_loaded is None:
        try:  # another library (the product's ops.py) already DEFined the schemas: loading would abort the process
            torch._C._dispatch_find_schema_or_throw("detectron2::nms_rotated", "")
            _ref_loaded = False
            return False
        except RuntimeError:
            pass
        so = _build.build_ref()
        if so is None or not os.path.exists(so):
            _ref_loaded = False
        else:
            try:
                torch.ops.load_library(so)
                _ref_loaded = True
            except Exception:  # schema already registered by another library in this process
                _ref_loaded = False
    return _ref_loaded


def _f(t):
    t = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
    return t, C.cast(t.data_ptr(), C.POINTER(C.c_float))


def _optf(t):
    if t is None:
        return None, None
    return _f(t)


def roi_align_forward(inp, rois, spatial_scale, ph, pw, sampling_ratio, aligned):
    inp, pi = _f(inp)
    rois, pr = _f(rois)
    n, c, h, w = inp.shape
    k = rois.shape[0]
    out = torch.zeros(k, c, ph, pw, dtype=torch.float32)
    lib().orc_roi_align_forward(pi, n, c, h, w, pr, k, C.c_float(spatial_scale), ph, pw, sampling_ratio,
                                int(bool(aligned)), C.cast(out.data_ptr(), C.POINTER(C.c_float)))
    return out


def roi_align_backward(grad, rois, spatial_scale, ph, pw, n, c, h, w, sampling_ratio, aligned):
    grad, pg = _f(grad)
    rois, pr = _f(rois)
    gin = torch.zeros(n, c, h, w, dtype=torch.float32)
    lib().orc_roi_align_backward(pg, pr, rois.shape[0], C.c_float(spatial_scale), ph, pw, n, c, h, w,
                                 sampling_ratio, int(bool(aligned)), C.cast(gin.data_ptr(), C.POINTER(C.c_float)))
    return gin


def roi_align_rotated_forward(inp, rois, spatial_scale, ph, pw, sampling_ratio):
    inp, pi = _f(inp)
    rois, pr = _f(rois)
    n, c, h, w = inp.shape
    k = rois.shape[0]
    out = torch.zeros(k, c, ph, pw, dtype=torch.float32)
    lib().orc_roi_align_rotated_forward(pi, n, c, h, w, pr, k, C.c_float(spatial_scale), ph, pw, sampling_ratio,
                                        C.cast(out.data_ptr(), C.POINTER(C.c_float)))
    return out


def roi_align_rotated_backward(grad, rois, spatial_scale, ph, pw, n, c, h, w, sampling_ratio):
    grad, pg = _f(grad)
    rois, pr = _f(rois)
    gin = torch.zeros(n, c, h, w, dtype=torch.float32)
    lib().orc_roi_align_rotated_backward(pg, pr, rois.shape[0], C.c_float(spatial_scale), ph, pw, n, c, h, w,
                                         sampling_ratio, C.cast(gin.data_ptr(), C.POINTER(C.c_float)))
    return gin


def nms(boxes, scores, iou_threshold):
    boxes, pb = _f(boxes)
    scores, ps = _f(scores)
    m = boxes.shape[0]
    keep = torch.zeros(max(m, 1), dtype=torch.int64)
    fn = lib().orc_nms
    fn.restype = C.c_int64
    nk = fn(pb, ps, C.c_int64(m), C.c_double(iou_threshold), C.cast(keep.data_ptr(), C.POINTER(C.c_int64)))
    return keep[:nk].clone()


def batched_nms(boxes, scores, idxs, iou_threshold):
    """torchvision.ops.boxes.batched_nms, coordinate-offset strategy (_batched_nms_coordinate_trick):
    offsets = idxs * (max_coordinate + 1) in the boxes' dtype (detectron2/layers/nms.py:11-22)."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    boxes = boxes.float()
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
    return nms(boxes + offsets[:, None], scores, iou_threshold)


def box_iou_rotated(b1, b2):
    b1, p1 = _f(b1)
    b2, p2 = _f(b2)
    n, m = b1.shape[0], b2.shape[0]
    out = torch.zeros(n, m, dtype=torch.float32)
    lib().orc_box_iou_rotated(p1, C.c_int64(n), p2, C.c_int64(m), C.cast(out.data_ptr(), C.POINTER(C.c_float)))
    return out


def nms_rotated(dets, scores, iou_threshold):
    dets, pd = _f(dets)
    scores, ps = _f(scores)
    m = dets.shape[0]
    keep = torch.zeros(max(m, 1), dtype=torch.int64)
    fn = lib().orc_nms_rotated
    fn.restype = C.c_int64
    nk = fn(pd, ps, C.c_int64(m), C.c_double(iou_threshold), C.cast(keep.data_ptr(), C.POINTER(C.c_int64)))
    return keep[:nk].clone()


def batched_nms_rotated(boxes, scores, idxs, iou_threshold):
    """detectron2/layers/nms.py:96-147 (min/max coordinate offset trick, fp32)."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    boxes = boxes.float()
    max_c = (torch.max(boxes[:, 0], boxes[:, 1]) + torch.max(boxes[:, 2], boxes[:, 3]) / 2).max()
    min_c = (torch.min(boxes[:, 0], boxes[:, 1]) - torch.max(boxes[:, 2], boxes[:, 3]) / 2).min()
    offsets = idxs.to(boxes) * (max_c - min_c + 1)
    b = boxes.clone()
    b[:, :2] += offsets[:, None]
    return nms_rotated(b, scores, iou_threshold)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def deform_conv_forward(x, offset, mask, weight, bias, stride, padding, dilation, groups, deformable_groups):
    x, px = _f(x)
    offset, po = _f(offset)
    mask, pm = _optf(mask)
    weight, pw_ = _f(weight)
    bias, pb = _optf(bias)
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    n, cin, h, w = x.shape
    cout, _, kh, kw = weight.shape
    ho = (h + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    wo = (w + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    out = torch.zeros(n, cout, ho, wo, dtype=torch.float32)
    lib().orc_deform_conv_forward(px, po, pm, pw_, pb, n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw,
                                  groups, deformable_groups, C.cast(out.data_ptr(), C.POINTER(C.c_float)))
    return out


def deform_conv_backward(x, offset, mask, weight, grad_out, stride, padding, dilation, groups,
                         deformable_groups, with_bias):
    x, px = _f(x)
    offset, po = _f(offset)
    mask, pm = _optf(mask)
    weight, pw_ = _f(weight)
    grad_out, pg = _f(grad_out)
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    n, cin, h, w = x.shape
    cout, _, kh, kw = weight.shape
    gx = torch.zeros_like(x)
    goff = torch.zeros_like(offset)
    gmask = torch.zeros_like(mask) if mask is not None else None
    gw = torch.zeros_like(weight)
    gb = torch.zeros(cout, dtype=torch.float32)
    P = lambda t: None if t is None else C.cast(t.data_ptr(), C.POINTER(C.c_float))  # noqa: E731
    lib().orc_deform_conv_backward(px, po, pm, pw_, pg, n, cin, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw,
                                   groups, deformable_groups, int(bool(with_bias)),
                                   P(gx), P(goff), P(gmask), P(gw), P(gb))
    return gx, goff, gmask, gw, (gb if with_bias else None)


def paste_masks(masks, boxes, image_shape, threshold=0.5, return_soft=False):
    masks, pm = _f(masks)
    boxes, pb = _f(boxes)
    n = masks.shape[0]
    m = masks.shape[-1]
    h, w = image_shape
    out = torch.zeros(n, h, w, dtype=torch.uint8)
    soft = torch.zeros(n, h, w, dtype=torch.float32) if return_soft else None
    lib().orc_paste_masks(pm, pb, n, m, h, w, C.c_float(threshold), C.cast(out.data_ptr(), C.POINTER(C.c_uint8)),
                          None if soft is None else C.cast(soft.data_ptr(), C.POINTER(C.c_float)))
    res = out.bool() if threshold >= 0 else out
    return (res, soft) if return_soft else res
