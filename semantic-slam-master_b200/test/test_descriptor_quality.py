"""Drop-in for the hot-path methods of the reference's ``test/test_descriptor_quality.py``:
``DescriptorQualityTester.extract_features`` (:69-95 there), ``find_mutual_nearest_neighbors``
(:97-142), ``compute_ground_truth_matches`` (:144-183), ``evaluate_matches`` (:185-231) — same names,
arguments and return values; arithmetic in the sm_100a kernels.  Dataset iteration, pose handling
and plotting of that script are out of scope."""

import numpy as np

from sslam_b200 import evaluation, matchers
from sslam_b200.harness import ModelHarness


class DescriptorQualityTester(ModelHarness):
    def find_mutual_nearest_neighbors(self, desc1: np.ndarray, desc2: np.ndarray, ratio_threshold: float = 0.9):
        return matchers.find_mutual_nearest_neighbors(desc1, desc2, ratio_threshold)

    def compute_ground_truth_matches(self, kpts1: np.ndarray, kpts2: np.ndarray, H: np.ndarray,
                                     threshold: float = 3.0):
        return evaluation.compute_ground_truth_matches(kpts1, kpts2, H, threshold)

    def evaluate_matches(self, pred_matches: np.ndarray, gt_matches: np.ndarray, num_kpts1: int, num_kpts2: int):
        return evaluation.evaluate_matches(pred_matches, gt_matches, num_kpts1, num_kpts2)
