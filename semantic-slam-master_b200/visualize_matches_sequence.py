"""Drop-in for the matching surface of the reference's ``visualize_matches_sequence.py``.

``SequenceMatcher.match_with_quality`` (a ``@staticmethod``; visualize_matches_sequence.py:106-197
there) keeps its signature, defaults and return types: matches (K', 2) int64 in ascending i and
quality (K',) fp32, empty arrays when nothing passes.  ``process_spacing`` is provided as a
device-resident driver (every frame extracted once).  Plotting is out of scope.
"""

from typing import Optional

import numpy as np
import torch

from sslam_b200 import matchers
from sslam_b200.pipeline import FrontEnd
from visualize_matches import MatchVisualizer


class SequenceMatcher(MatchVisualizer):
    @torch.no_grad()
    def extract_from_patch_map(self, dino_features):
        out = self.features_from_patch_map(dino_features)
        sal = self.selector(dino_features)
        out["saliency"] = sal[0, :, :, 0].cpu().numpy()
        return out

    @staticmethod
    def match_with_quality(desc1: np.ndarray, desc2: np.ndarray, scores1: np.ndarray,
                           scores2: np.ndarray, saliency_weight: float = 0.3,
                           min_saliency: float = 0.2, min_descriptor_sim: float = 0.7,
                           intensity1: Optional[np.ndarray] = None,
                           intensity2: Optional[np.ndarray] = None, min_intensity: float = 0.1):
        return matchers.match_with_quality(desc1, desc2, scores1, scores2, saliency_weight,
                                           min_saliency, min_descriptor_sim, intensity1, intensity2,
                                           min_intensity)

    @torch.no_grad()
    def process_spacing(self, saliency, features, spacing=1, num_keypoints=None, **thresholds):
        """Match frame i with frame i+spacing for i = 0, spacing, 2*spacing, ... (the loop of
        visualize_matches_sequence.py:297 there) from device tensors; returns padded device lists."""
        K = num_keypoints or self.config["model"]["num_keypoints"]
        fe = FrontEnd(self.refiner, num_keypoints=K, grid="pixel")
        feats = fe.extract(saliency, features)
        T = saliency.shape[0]
        idx = torch.arange(0, T - spacing, spacing, device=saliency.device, dtype=torch.int32)
        pair_index = torch.stack([idx, idx + spacing], dim=1)
        return fe.match_pairs(feats, pair_index, matchers.M2, **thresholds)
