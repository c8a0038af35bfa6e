"""BoxList: the reference's container type for boxes + per-box fields
(contract of mega_core/structures/bounding_box.py:9-266 of the reference: `.bbox` [N,4] fp32,
`.size` = (image_width, image_height), `.mode` in {"xyxy","xywh"}, named extra fields that are
indexed together with the boxes). Only what the inference path touches is provided."""
import torch


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should be [N,4], got %s" % (tuple(bbox.shape),))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox, self.size, self.mode = bbox, image_size, mode
        self.extra_fields = {}

    # ---- fields
    def add_field(self, name, data):
        self.extra_fields[name] = data

    def get_field(self, name):
        return self.extra_fields[name]

    def has_field(self, name):
        return name in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _like(self, bbox, mode=None):
        out = BoxList(bbox, self.size, mode or self.mode)
        return out

    def copy_with_fields(self, names, skip_missing=False):
        out = self._like(self.bbox)
        for n in ([names] if not isinstance(names, (list, tuple)) else names):
            if self.has_field(n):
                out.add_field(n, self.get_field(n))
            elif not skip_missing:
                raise KeyError("Field '%s' not found in %s" % (n, self))
        return out

    # ---- geometry
    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        x1, y1, a, b = self.bbox.unbind(1)
        if mode == "xyxy":      # from xywh; the "+1" pixel convention of the reference
            box = torch.stack((x1, y1, x1 + (a - 1).clamp(min=0), y1 + (b - 1).clamp(min=0)), 1)
        else:
            box = torch.stack((x1, y1, a - x1 + 1, b - y1 + 1), 1)
        out = self._like(box, mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v)
        return out

    def clip_to_image(self, remove_empty=True):
        w, h = self.size
        self.bbox[:, 0].clamp_(min=0, max=w - 1)
        self.bbox[:, 1].clamp_(min=0, max=h - 1)
        self.bbox[:, 2].clamp_(min=0, max=w - 1)
        self.bbox[:, 3].clamp_(min=0, max=h - 1)
        if remove_empty:
            b = self.bbox
            return self[(b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])]
        return self

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return b[:, 2] * b[:, 3]

    def resize(self, size):
        rw, rh = float(size[0]) / self.size[0], float(size[1]) / self.size[1]
        bl = self.convert("xyxy")
        scale = torch.tensor([rw, rh, rw, rh], device=bl.bbox.device)
        out = BoxList(bl.bbox * scale, size, "xyxy")
        for k, v in self.extra_fields.items():
            out.add_field(k, v)
        return out.convert(self.mode)

    # ---- container protocol
    def to(self, device):
        out = self._like(self.bbox.to(device))
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = self._like(self.bbox[item])
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def __repr__(self):
        return "BoxList(num_boxes=%d, image_width=%d, image_height=%d, mode=%s)" % (
            len(self), self.size[0], self.size[1], self.mode)
