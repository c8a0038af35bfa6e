from ..nets import CombinedROIHeads, build_roi_heads  # noqa: F401
