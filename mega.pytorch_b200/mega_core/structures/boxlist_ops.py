"""boxlist_nms / remove_small_boxes / cat_boxlist (contract of structures/boxlist_ops.py:9-48,
:103-133 of the reference), on top of the CUDA `_C.nms`."""
import torch

from .bounding_box import BoxList
from ..layers import nms as _box_nms


def boxlist_nms(boxlist, nms_thresh, max_proposals=-1, score_field="scores"):
    if nms_thresh <= 0:
        return boxlist
    mode = boxlist.mode
    boxlist = boxlist.convert("xyxy")
    keep = _box_nms(boxlist.bbox, boxlist.get_field(score_field), nms_thresh)
    if max_proposals > 0:
        keep = keep[:max_proposals]
    return boxlist[keep].convert(mode)


def remove_small_boxes(boxlist, min_size):
    wh = boxlist.convert("xywh").bbox
    keep = ((wh[:, 2] >= min_size) & (wh[:, 3] >= min_size)).nonzero().squeeze(1)
    return boxlist[keep]


def cat_boxlist(bboxes):
    assert isinstance(bboxes, (list, tuple)) and all(isinstance(b, BoxList) for b in bboxes)
    size, mode, fields = bboxes[0].size, bboxes[0].mode, set(bboxes[0].fields())
    assert all(b.size == size and b.mode == mode and set(b.fields()) == fields for b in bboxes)
    out = BoxList(torch.cat([b.bbox for b in bboxes], 0), size, mode)
    for f in fields:
        out.add_field(f, torch.cat([b.get_field(f) for b in bboxes], 0))
    return out
