from ..nets import build_backbone  # noqa: F401
