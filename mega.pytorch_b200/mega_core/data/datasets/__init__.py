from .vid import VIDDataset, VIDDFFDataset, VIDFGFADataset, VIDMEGADataset, VIDRDNDataset  # noqa: F401

__all__ = ["VIDDataset", "VIDRDNDataset", "VIDMEGADataset", "VIDFGFADataset", "VIDDFFDataset"]
