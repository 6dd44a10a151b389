"""Drop-in for the hot-path part of the reference's ``test/test_tracking.py``:
``TrackingTester.extract_features`` (:63-85 there) and the frame-to-frame loop of
``track_frame_sequence`` (:139-199) — here fed with tensors that are already past the backbone
(there is no TUM data or ViT in this image) and run from an HBM frame store with the counters on
the device (sslam_b200/framestore.py)."""

import torch

from sslam_b200 import framestore, matchers
from sslam_b200.harness import ModelHarness
from sslam_b200.pipeline import FrontEnd


class TrackingTester(ModelHarness):
    def count_matches(self, desc_prev, desc_curr, match_threshold: float = 0.8) -> int:
        """``(desc_prev @ desc_curr.T).max(axis=1) > match_threshold`` summed (:159-161)."""
        return matchers.tracking_count(desc_prev, desc_curr, match_threshold)

    @torch.no_grad()
    def track_frame_sequence(self, saliency, features, max_frames: int = 100, min_matches: int = 50,
                             match_threshold: float = 0.8, frame_spacing: int = 1, sequence: str = "",
                             grid: str = "pixel"):
        """The statistics dict of the reference's track_frame_sequence, computed from saliency maps
        (T,H,W,1) and feature maps (T,h,w,C) on the device."""
        fe = FrontEnd(self.refiner, num_keypoints=self.config["model"]["num_keypoints"], grid=grid)
        res = framestore.track_sequence(fe, saliency, features, frame_spacing=frame_spacing, max_frames=max_frames,
                                        min_matches=min_matches, match_threshold=match_threshold)
        res["sequence"] = sequence
        return res
