"""Input side of the hot path: the reference's test-time transforms on the B200."""
from .transforms import build_transforms  # noqa: F401
