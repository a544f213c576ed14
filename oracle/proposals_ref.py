"""CPU restatement of find_top_rpn_proposals (TEST INFRASTRUCTURE), following
detectron2/modeling/proposal_generator/proposal_utils.py:22-135 step by step (per-image loop, boolean filtering,
batched_nms through the oracle).  Pinned against the real reference in tests/test_oracle_pins.py via the fixture
tests/golden/rpn_proposals.npz."""
import torch

from . import oracle as orc


def find_top_rpn_proposals(proposals, pred_objectness_logits, image_sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_size, training):
    num_images = len(image_sizes)
    batch_idx = torch.arange(num_images)
    topk_scores, topk_proposals, level_ids = [], [], []
    for level_id, (p_i, l_i) in enumerate(zip(proposals, pred_objectness_logits)):  # :70-94
        k = min(l_i.shape[1], pre_nms_topk)
        s_i, idx = l_i.topk(k, dim=1)
        topk_proposals.append(p_i[batch_idx[:, None], idx])
        topk_scores.append(s_i)
        level_ids.append(torch.full((k,), level_id, dtype=torch.int64))
    topk_scores = torch.cat(topk_scores, dim=1)
    topk_proposals = torch.cat(topk_proposals, dim=1)
    level_ids = torch.cat(level_ids, dim=0)
    results = []
    for n, (h, w) in enumerate(image_sizes):  # :100-134
        boxes, scores, lvl = topk_proposals[n].clone(), topk_scores[n], level_ids
        valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores)
        if not valid.all():
            if training:
                raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
            boxes, scores, lvl = boxes[valid], scores[valid], lvl[valid]
        boxes[:, 0].clamp_(min=0, max=w)  # Boxes.clip
        boxes[:, 1].clamp_(min=0, max=h)
        boxes[:, 2].clamp_(min=0, max=w)
        boxes[:, 3].clamp_(min=0, max=h)
        keep = ((boxes[:, 2] - boxes[:, 0]) > min_box_size) & ((boxes[:, 3] - boxes[:, 1]) > min_box_size)  # nonempty
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]
        keep = orc.batched_nms(boxes, scores, lvl, nms_thresh)[:post_nms_topk]
        results.append((boxes[keep], scores[keep]))
    return results
