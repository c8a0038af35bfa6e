from ...nets import FPNPredictor, MEGAFeatureExtractor, ResNetConv52MLPFeatureExtractor, ROIBoxHead  # noqa: F401
