// Host-side native helper of the VID evaluator: the per-(image, class) greedy matching of
// score-sorted detections against ground truth, which the reference runs as nested Python loops over
// 176 k images x classes x detections x boxes (data/datasets/evaluation/vid/vid_eval.py:201-262). No device code: it
// lives in libmega_b200.so so that the evaluator needs no second extension. float32 IoU arithmetic in the order of
// boxlist_iou (structures/boxlist_ops.py:75-88) on the "+1" boxes of vid_eval.py:213-217, so decisions are identical.
#include <vector>
#include "common.cuh"
#include "mega_b200.h"

extern "C" int mega_vid_match_host(const float* pred_boxes, int n_pred, const float* gt_boxes,
                                   const unsigned char* gt_ignore, int n_gt, float iou_thresh, double empty_weight,
                                   signed char* match_out, double* pred_ignore_out) {
  MEGA_ARG_CHECK(n_pred >= 0 && n_gt >= 0, "vid_match: negative sizes");
  if (n_pred == 0) return MEGA_OK;
  if (n_gt == 0) {                                   // vid_eval.py:207-210
    for (int j = 0; j < n_pred; ++j) match_out[j] = 0, pred_ignore_out[j] = empty_weight;
    return MEGA_OK;
  }
  std::vector<float> garea(n_gt), gx1(n_gt), gy1(n_gt), gx2(n_gt), gy2(n_gt);
  int n_ignored = 0;
  for (int k = 0; k < n_gt; ++k) {
    gx1[k] = gt_boxes[4 * k], gy1[k] = gt_boxes[4 * k + 1];
    gx2[k] = gt_boxes[4 * k + 2] + 1.0f, gy2[k] = gt_boxes[4 * k + 3] + 1.0f;       // integer-typed boxes: [:, 2:] += 1
    garea[k] = (gx2[k] - gx1[k] + 1.0f) * (gy2[k] - gy1[k] + 1.0f);
    n_ignored += gt_ignore[k] ? 1 : 0;
  }
  std::vector<unsigned char> selec(n_gt, 0);
  for (int j = 0; j < n_pred; ++j) {
    const float px1 = pred_boxes[4 * j], py1 = pred_boxes[4 * j + 1];
    const float px2 = pred_boxes[4 * j + 2] + 1.0f, py2 = pred_boxes[4 * j + 3] + 1.0f;
    const float parea = (px2 - px1 + 1.0f) * (py2 - py1 + 1.0f);
    double iou_match = iou_thresh, iou_match_ig = -1.0, iou_match_nig = -1.0;
    int arg_match = -1;
    for (int k = 0; k < n_gt; ++k) {
      const float ltx = px1 > gx1[k] ? px1 : gx1[k], lty = py1 > gy1[k] ? py1 : gy1[k];
      const float rbx = px2 < gx2[k] ? px2 : gx2[k], rby = py2 < gy2[k] ? py2 : gy2[k];
      float w = rbx - ltx + 1.0f, h = rby - lty + 1.0f;
      w = w < 0.f ? 0.f : w, h = h < 0.f ? 0.f : h;
      const float inter = w * h;
      const float iou = inter / (parea + garea[k] - inter);
      const double v = iou;
      if (gt_ignore[k] && v > iou_match_ig) iou_match_ig = v;
      if (!gt_ignore[k] && v > iou_match_nig) iou_match_nig = v;
      if (selec[k] || v < iou_match) continue;
      if (v == iou_match) {
        if (arg_match < 0 || gt_ignore[arg_match]) arg_match = k;
      } else {
        arg_match = k;
      }
      iou_match = v;
    }
    if (arg_match >= 0) {
      match_out[j] = 1;
      pred_ignore_out[j] = gt_ignore[arg_match] ? 1.0 : 0.0;
      selec[arg_match] = 1;
    } else {
      match_out[j] = 0;
      if (iou_match_nig > iou_match_ig) pred_ignore_out[j] = 0.0;
      else if (iou_match_ig > iou_match_nig) pred_ignore_out[j] = 1.0;
      else pred_ignore_out[j] = static_cast<double>(n_ignored) / static_cast<double>(n_gt);
    }
  }
  return MEGA_OK;
}
