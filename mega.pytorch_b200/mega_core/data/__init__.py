"""Input side of the hot path: the reference's test-time transforms on the B200, and the
test-time ImageNet-VID datasets / loaders around them."""
from .build import make_data_loader  # noqa: F401
from .transforms import build_transforms  # noqa: F401
