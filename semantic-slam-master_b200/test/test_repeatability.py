"""Drop-in for the hot-path methods of the reference's ``test/test_repeatability.py``:
``RepeatabilityTester.detect_keypoints`` (:60-77 there) and ``compute_repeatability`` (:79-128)."""

import numpy as np
import torch

from sslam_b200 import evaluation
from sslam_b200.harness import ModelHarness


class RepeatabilityTester(ModelHarness):
    @torch.no_grad()
    def detect_keypoints(self, image: torch.Tensor):
        dino_features = self._features(image)
        saliency_map = self.selector(dino_features)
        keypoints_patch, scores = self.selector.select_keypoints(
            saliency_map, num_keypoints=self.config["model"]["num_keypoints"])
        keypoints_pixel = self.backbone.patch_to_pixel(keypoints_patch)
        return keypoints_pixel[0].cpu().numpy(), scores[0].cpu().numpy()

    def compute_repeatability(self, kpts1: np.ndarray, kpts2: np.ndarray, H: np.ndarray = None,
                              threshold: float = 3.0) -> dict:
        return evaluation.compute_repeatability(kpts1, kpts2, H, threshold)
