"""Dataset locations with the reference's names (config/paths_catalog.py:7-232, VID / DET entries)."""
import os


class DatasetCatalog(object):
    DATA_DIR = "datasets"
    DATASETS = {name: {"img_dir": "ILSVRC2015/Data/" + kind, "anno_path": "ILSVRC2015/Annotations/" + kind,
                       "img_index": "ILSVRC2015/ImageSets/%s.txt" % name}
                for name, kind in (("DET_train_30classes", "DET"), ("VID_train_15frames", "VID"),
                                   ("VID_train_every10frames", "VID"), ("VID_val_frames", "VID"), ("VID_val_videos", "VID"))}
    FACTORIES = {"base": "VIDDataset", "rdn": "VIDRDNDataset", "mega": "VIDMEGADataset", "fgfa": "VIDFGFADataset",
                 "dff": "VIDDFFDataset"}

    @classmethod
    def get(cls, name, method="base"):
        """(a classmethod, so a subclass that only overrides DATA_DIR relocates the tree)"""
        if name not in cls.DATASETS:
            raise RuntimeError("Dataset not available: {}".format(name))
        root, attrs = cls.DATA_DIR, cls.DATASETS[name]
        return dict(factory=cls.FACTORIES[method],
                    args=dict(image_set=name, data_dir=root, img_dir=os.path.join(root, attrs["img_dir"]),
                              anno_path=os.path.join(root, attrs["anno_path"]), img_index=os.path.join(root, attrs["img_index"])))
