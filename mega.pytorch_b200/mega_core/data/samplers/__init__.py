from .distributed import VIDTestDistributedSampler  # noqa: F401
