"""Stub of NVIDIA apex for importing the reference (fp32 inference needs none of it)."""
from . import amp  # noqa: F401
