"""ImageList / to_image_list (contract of the reference's structures/image_list.py:7-72):
a batch tensor [N,C,H,W] plus the un-padded (height, width) of every image."""
import torch


class ImageList(object):
    def __init__(self, tensors, image_sizes):
        self.tensors = tensors
        self.image_sizes = image_sizes

    def to(self, *args, **kwargs):
        return ImageList(self.tensors.to(*args, **kwargs), self.image_sizes)


def to_image_list(tensors, size_divisible=0):
    """Tensor [C,H,W] / [N,C,H,W], list of [C,H,W] tensors, or ImageList -> ImageList (zero padded)."""
    if isinstance(tensors, ImageList):
        return tensors
    if isinstance(tensors, torch.Tensor):
        if size_divisible > 0:
            tensors = [tensors] if tensors.dim() == 3 else list(tensors)
        else:
            if tensors.dim() == 3:
                tensors = tensors[None]
            assert tensors.dim() == 4
            return ImageList(tensors, [tuple(t.shape[-2:]) for t in tensors])
    if isinstance(tensors, (tuple, list)):
        c = tensors[0].shape[0]
        hm = max(t.shape[1] for t in tensors)
        wm = max(t.shape[2] for t in tensors)
        if size_divisible > 0:
            hm = (hm + size_divisible - 1) // size_divisible * size_divisible
            wm = (wm + size_divisible - 1) // size_divisible * size_divisible
        batch = tensors[0].new_zeros((len(tensors), c, hm, wm))
        for img, pad in zip(tensors, batch):
            pad[:, :img.shape[1], :img.shape[2]].copy_(img)
        return ImageList(batch, [tuple(t.shape[-2:]) for t in tensors])
    raise TypeError("Unsupported type for to_image_list: %s" % type(tensors))
