from .detectors import build_detection_model, build_detection_model_from_state_dict  # noqa: F401
