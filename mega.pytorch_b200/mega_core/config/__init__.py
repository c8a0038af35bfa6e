from .defaults import _C as cfg, CfgNode  # noqa: F401
