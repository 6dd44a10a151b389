"""Oracle: evaluation adaptors in NumPy (test infrastructure, see oracle/__init__.py).

Restates the keypoint-geometry scoring of the reference's evaluation scripts:
  compute_ground_truth_matches   test/test_descriptor_quality.py:144-183
  evaluate_matches               test/test_descriptor_quality.py:185-231
  compute_repeatability          test/test_repeatability.py:79-128
Pinned by tests/golden/evaluation.npz (written by oracle/pin_against_reference.py from the
reference's own methods).
"""

import numpy as np


def warp_points(kpts, H):
    """Homogeneous warp in float64 (the ones column is float64, :164 / :100)."""
    homo = np.concatenate([kpts, np.ones((len(kpts), 1))], axis=1)
    w = (H @ homo.T).T
    return w[:, :2] / w[:, 2:3]


def nearest(kpts1_warped, kpts2):
    """min / argmin over the (N, M) Euclidean distance matrix (:169-176 / :108-113)."""
    d = np.linalg.norm(kpts1_warped[:, None, :] - kpts2[None, :, :], axis=2)
    return d.min(axis=1), d.argmin(axis=1)


def compute_ground_truth_matches(kpts1, kpts2, H, threshold=3.0):
    md, am = nearest(warp_points(kpts1, H), kpts2)
    idx1 = np.where(md < threshold)[0]                                     # :178-181
    return np.stack([idx1, am[idx1]], axis=1)


def evaluate_matches(pred_matches, gt_matches, num_kpts1, num_kpts2):
    pred = set(map(tuple, pred_matches))
    gt = set(map(tuple, gt_matches))
    tp, fp, fn = len(pred & gt), len(pred - gt), len(gt - pred)            # :202-209
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0.0
    return {"tp": tp, "fp": fp, "fn": fn, "precision": precision, "recall": recall, "f1": f1,
            "inlier_ratio": tp / len(pred_matches) if len(pred_matches) > 0 else 0.0,
            "num_pred_matches": len(pred_matches), "num_gt_matches": len(gt_matches)}


def compute_repeatability(kpts1, kpts2, H=None, threshold=3.0):
    w = warp_points(kpts1, H) if H is not None else kpts1                  # :98-105
    md, _ = nearest(w, kpts2)
    repeatable = (md < threshold).sum()                                    # :116
    return {"repeatability": repeatable / len(kpts1), "repeatable_count": repeatable,
            "total_keypoints": len(kpts1), "mean_nn_distance": md.mean(),
            "median_nn_distance": np.median(md)}
