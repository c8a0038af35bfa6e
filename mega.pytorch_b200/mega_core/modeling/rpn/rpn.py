from ..nets import RPNHead, RPNModule, build_rpn  # noqa: F401
