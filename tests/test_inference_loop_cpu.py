"""mega_core.engine.inference.inference (engine/inference.py:72-134 of the reference): the loop around the hot path --
batches to the device, model(images), per-rank results to rank 0, predictions.pth, VID evaluation -- with a stand-in
model that answers with the fixture's detections; the AP it reports must be the reference evaluator's."""
import logging
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Cfg:
    class MODEL:
        class VID:
            METHOD = "mega"


def test_inference_loop_reports_the_reference_ap(tmp_path):
    from mega_core.engine.inference import inference, inference_no_model
    from mega_core.structures.bounding_box import BoxList
    case = torch.load(os.path.join(ROOT, "tests", "golden", "vid_eval.pt"), weights_only=False)[0]
    images = case["images"]

    class Dataset(object):
        def __len__(self):
            return len(images)

        def get_img_info(self, i):
            return {"width": images[i]["size"][0], "height": images[i]["size"][1]}

        def get_groundtruth(self, i):
            gt = BoxList(images[i]["gt"], images[i]["size"], mode="xyxy")
            gt.add_field("labels", images[i]["gt_labels"])
            return gt

        def map_class_id_to_class_name(self, i):
            return "class%d" % i

    class Loader(object):
        dataset = Dataset()

        def __iter__(self):
            for i in range(len(images)):            # the dict VIDMEGADataset builds, one image per batch
                yield {"cur": torch.zeros(3, 4, 4), "ref_l": [torch.zeros(3, 4, 4)], "frame_category": 1, "id": i}, None, [i]

    class Model(torch.nn.Module):
        def forward(self, batch):
            im = images[batch["id"]]
            out = BoxList(im["boxes"], im["size"], mode="xyxy")
            out.add_field("scores", im["scores"])
            out.add_field("labels", im["labels"])
            return [out]

    logging.getLogger("mega_core.inference").setLevel(logging.ERROR)
    res = inference(_Cfg, Model(), Loader(), "VID_val_synthetic", device="cpu", output_folder=str(tmp_path))
    want = case["reference"]["all"]["ap"]
    assert np.array_equal(res[0]["ap"], want, equal_nan=True) and abs(res[0]["map"] - np.nanmean(want)) < 1e-15
    assert os.path.exists(os.path.join(tmp_path, "predictions.pth")) and os.path.exists(os.path.join(tmp_path, "result.txt"))
    again = inference_no_model(Loader(), output_folder=str(tmp_path))
    assert np.array_equal(again[0]["ap"], want, equal_nan=True)
