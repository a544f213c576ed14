"""Host logic of the batched post-processing functions, checked on CPU against the fixtures generated from the
REAL reference functions (tests/golden/make_golden.py).  The product calls `ops.nms_fixed` (CUDA only, no CPU fallback); here
that one call is replaced by a stand-in built on the CPU oracle, so that everything around it -- candidate selection,
box decoding, category / offset construction, per-image extraction -- is verified without a GPU.  The `-m gpu` tests run
the same fixtures through the real kernels."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

T = torch.from_numpy


def oracle_nms_fixed(boxes, scores, idxs, iou_threshold, rotated, apply_offsets=True, max_segment=0):
    """Same contract as detectron2_b200.ops.nms_fixed (padded keep buffer + count), computed by the CPU oracle."""
    assert not rotated
    boxes, scores = boxes.float().contiguous(), scores.float().contiguous()
    m = boxes.shape[0]
    if idxs is None:
        kept = orc.nms(boxes, scores, iou_threshold)
    elif apply_offsets:
        kept = orc.batched_nms(boxes, scores, idxs, iou_threshold)
    else:  # D2B_NMS_NO_OFFSET: idxs are pure segment ids, coordinates used as given
        parts = []
        for c in torch.unique(idxs):
            if c < 0:  # ignored slots
                continue
            ii = torch.nonzero(idxs == c, as_tuple=True)[0]
            assert max_segment <= 0 or len(ii) <= max_segment, "caller's max_segment bound violated"
            parts.append(ii[orc.nms(boxes[ii], scores[ii], iou_threshold)])
        kept = torch.cat(parts).sort().values if parts else torch.zeros(0, dtype=torch.int64)
        kept = kept[torch.sort(scores[kept], descending=True, stable=True).indices]  # score order, lower index first on ties
    keep = torch.zeros(m, dtype=torch.int64)
    keep[: kept.numel()] = kept
    return keep, torch.tensor([kept.numel()], dtype=torch.int64)


@pytest.fixture()
def cpu_nms(monkeypatch):
    from detectron2_b200 import ops

    monkeypatch.setattr(ops, "nms_fixed", oracle_nms_fixed)


def test_find_top_rpn_proposals_host_logic(golden, cpu_nms):
    from detectron2_b200.proposal_utils import find_top_rpn_proposals
    from test_oracle_pins import _rpn_fixture

    d, props, logits, sizes, thr, pre, post, mbs = _rpn_fixture(golden)
    res = find_top_rpn_proposals(props, logits, sizes, thr, pre, post, mbs, False)
    for i, r in enumerate(res):
        assert torch.equal(r.proposal_boxes.tensor, T(d[f"boxes_img{i}"])), i
        assert torch.equal(r.objectness_logits, T(d[f"scores_img{i}"])), i
    with pytest.raises(FloatingPointError):
        find_top_rpn_proposals(props, logits, sizes, thr, pre, post, mbs, True)


def test_fast_rcnn_inference_host_logic(golden, cpu_nms):
    from detectron2_b200 import fast_rcnn_inference as fri

    d = golden("fast_rcnn_inference")
    thr, nms_thr, topk = float(d["cfg"][0]), float(d["cfg"][1]), int(d["cfg"][2])
    shapes = [tuple(int(v) for v in r) for r in d["shapes"]]
    res, rows = fri.fast_rcnn_inference([T(d["boxes0"]), T(d["boxes0"])], [T(d["scores0"]), T(d["scores0"])],
                                        [shapes[0]] * 2, thr, nms_thr, topk)
    for j in range(2):
        assert torch.equal(res[j].pred_boxes, T(d["out_boxes0"]))
        assert torch.equal(res[j].scores, T(d["out_scores0"]))
        assert torch.equal(res[j].pred_classes, T(d["out_classes0"]))
        assert torch.equal(rows[j], T(d["out_rows0"]))
    res, rows = fri.fast_rcnn_inference([T(d["boxes1"])], [T(d["scores1"])], [shapes[1]], thr, nms_thr, topk)
    assert torch.equal(res[0].pred_boxes, T(d["out_boxes1"])) and torch.equal(rows[0], T(d["out_rows1"]))


def _retinanet_fixture(golden):
    d = golden("retinanet_inference")
    thr, topk, nms_thr, max_det = float(d["cfg"][0]), int(d["cfg"][1]), float(d["cfg"][2]), int(d["cfg"][3])
    sizes = [tuple(int(v) for v in r) for r in d["image_sizes"]]
    nl = len([k for k in d.files if k.startswith("anchors")])
    anchors = [T(d[f"anchors{l}"]) for l in range(nl)]
    logits = [T(d[f"logits{l}"]) for l in range(nl)]
    deltas = [T(d[f"deltas{l}"]) for l in range(nl)]
    return d, anchors, logits, deltas, sizes, thr, topk, nms_thr, max_det


def check_retinanet(d, res, image_ids):
    for j, i in enumerate(image_ids):
        assert torch.equal(res[j].pred_boxes.cpu(), T(d[f"out_boxes{i}"])), i
        assert torch.equal(res[j].scores.cpu(), T(d[f"out_scores{i}"])), i
        assert torch.equal(res[j].pred_classes.cpu(), T(d[f"out_classes{i}"])), i


def test_apply_deltas_matches_reference_decode(golden):
    # single level, one candidate per anchor: decoded boxes of the fixture's kept detections appear in the output; the
    # clamp case (delta 9.0 > log(1000/16)) and the formula are checked directly
    from detectron2_b200.dense_inference import apply_deltas

    boxes = torch.tensor([[10.0, 20.0, 30.0, 60.0]])
    out = apply_deltas(torch.tensor([[0.5, -0.25, 9.0, 0.0]]), boxes)
    w = 20.0 * 1000.0 / 16
    assert torch.allclose(out, torch.tensor([[30.0 - w / 2, 30.0 - 20.0, 30.0 + w / 2, 30.0 + 20.0]]), rtol=1e-5)
    out2 = apply_deltas(torch.tensor([[1.0, 1.0, 0.0, 0.0]]), boxes, weights=(10.0, 10.0, 5.0, 5.0))
    assert torch.allclose(out2, torch.tensor([[12.0, 24.0, 32.0, 64.0]]))
    assert apply_deltas(torch.zeros(0, 4), torch.zeros(0, 4)).shape == (0, 4)


def test_retinanet_inference_host_logic(golden, cpu_nms):
    """Bit-exact against the REAL DenseDetector decode + batched_nms of the reference (fixture from make_golden.py)."""
    from detectron2_b200.dense_inference import dense_detector_inference, retinanet_inference

    d, anchors, logits, deltas, sizes, thr, topk, nms_thr, max_det = _retinanet_fixture(golden)
    res = retinanet_inference(anchors, logits, deltas, sizes, thr, topk, nms_thr, max_det)
    assert [len(r) for r in res] == [len(d["out_scores0"]), len(d["out_scores1"])]
    check_retinanet(d, res, [0, 1])
    # images in a different batch composition give the same per-image results (no cross-image leakage)
    res = retinanet_inference(anchors, [x[[1, 0, 1]] for x in logits], [x[[1, 0, 1]] for x in deltas],
                              [sizes[1], sizes[0], sizes[1]], thr, topk, nms_thr, max_det)
    check_retinanet(d, res, [1, 0, 1])
    # no candidate at all -> empty results; unlimited detections -> every survivor, still score-ordered
    res = dense_detector_inference(anchors, [torch.zeros_like(x) for x in logits], deltas, sizes, 0.5, topk, nms_thr, max_det)
    assert all(len(r) == 0 for r in res)
    res = retinanet_inference(anchors, logits, deltas, sizes, thr, topk, nms_thr, -1)
    for i, r in enumerate(res):
        assert len(r) >= max_det and torch.equal(r.scores[:max_det], T(d[f"out_scores{i}"]))
        assert (r.scores[:-1] >= r.scores[1:]).all()


# ------------------------------------------------------------------------------- 8f-4: mask targets, detector post-processing
class _OracleROIAlign:
    """Stand-in for layers.ROIAlign on CPU tensors (the product op has no CPU path)."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        self.a = (output_size, spatial_scale, sampling_ratio, aligned)

    def forward(self, x, rois):
        (ph, pw), scale, sr, aligned = self.a
        return orc.roi_align_forward(x, rois, scale, ph, pw, sr, aligned)


def _postprocess_fixture(golden):
    d = golden("postprocessing")
    h, w, oh, ow = [int(v) for v in d["hw"]]
    return d, h, w, oh, ow


def check_postprocess(d, res, oh, ow):
    assert res.image_size == (oh, ow)
    assert torch.equal(res.pred_boxes.cpu(), T(d["out_boxes"]))
    assert torch.equal(res.scores.cpu(), T(d["out_scores"])) and torch.equal(res.pred_classes.cpu(), T(d["out_classes"]))
    ref = T(d["out_masks"])
    assert res.pred_masks.shape == ref.shape and res.pred_masks.dtype == torch.bool
    # the reference pastes through grid_sample; pixels whose soft value sits within rounding of the threshold may differ
    assert (res.pred_masks.cpu() != ref).sum().item() <= 3


def check_crops(d, crops):
    ref = T(d["crops"])
    assert crops.shape == ref.shape and crops.dtype == torch.bool
    # RoIAlign of a 0/1 mask produces exact halves (2 of 4 samples inside); `>= 0.5` there depends on the summation order
    assert (crops.cpu() != ref).float().mean().item() <= 0.003


def test_detector_postprocess_and_crop_host_logic(golden, monkeypatch):
    from detectron2_b200 import postprocessing as pp
    from detectron2_b200.fast_rcnn_inference import Detections

    monkeypatch.setattr(pp, "paste_masks_in_image", lambda m, b, hw, threshold=0.5: orc.paste_masks(m, b, hw, threshold))
    monkeypatch.setattr(pp, "ROIAlign", _OracleROIAlign)
    d, h, w, oh, ow = _postprocess_fixture(golden)
    det = Detections((h, w), T(d["boxes"]), T(d["scores"]), T(d["classes"]))
    res = pp.detector_postprocess(det, oh, ow, 0.5, pred_masks=T(d["masks"]))
    check_postprocess(d, res, oh, ow)
    assert torch.equal(det.pred_boxes, T(d["boxes"]))  # the input record is not modified
    nomask = pp.detector_postprocess(det, oh, ow)
    assert nomask.pred_masks is None and torch.equal(nomask.pred_boxes, T(d["out_boxes"]))
    check_crops(d, pp.crop_and_resize(T(d["bit_masks"]), T(d["crop_boxes"]), int(d["mask_size"])))


def test_batched_nms_of_several_images_host_logic(cpu_nms):
    """`layers.batched_nms_images_fixed`: the category / per-image-offset construction around the one NMS call, against
    torchvision's own batched_nms run per image (the reference's loop, proposal_utils.py:96-133) on CPU."""
    import torchvision

    import detectron2_b200.layers as L

    g = torch.Generator().manual_seed(2)
    n, m = 3, 700
    boxes, scores = [], []
    for i in range(n):
        ctr = torch.rand(m, 2, generator=g) * torch.tensor([900.0 * (1 + i), 600.0])  # a different max coordinate per image
        wh = 8 + torch.rand(m, 2, generator=g) * 150
        boxes.append(torch.cat([ctr - wh / 2, ctr + wh / 2], 1))
        scores.append((torch.rand(m, generator=g) * 128).round() / 128)
    idxs = torch.randint(0, 5, (m,), generator=g)
    keep, num = L.batched_nms_images_fixed(boxes, scores, idxs, 0.7, 5, max_segment=m)
    kept = keep[: int(num)]
    for i in range(n):
        mine = kept[(kept >= i * m) & (kept < (i + 1) * m)] - i * m
        ref = torchvision.ops.boxes.batched_nms(boxes[i], scores[i], idxs, 0.7)  # CPU: coordinate trick below 4 000 elements
        assert sorted(mine.tolist()) == sorted(ref.tolist()), i
        assert torch.equal(mine, orc.batched_nms(boxes[i], scores[i], idxs, 0.7)), i  # order: score desc, lower index first on ties
