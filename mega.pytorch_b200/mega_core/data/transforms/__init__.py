from .build import build_transforms  # noqa: F401
from .transforms import DeviceTestTransform, get_size, resample_tables  # noqa: F401
