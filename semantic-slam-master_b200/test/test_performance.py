"""Drop-in for the hot-path methods of the reference's ``test/test_performance.py``:
``PerformanceTester.forward_pass`` (:146-157 there) and ``measure_component_times`` (:88-144),
timed with CUDA events on the launching stream instead of ``perf_counter`` + synchronize."""

import numpy as np
import torch

from sslam_b200.harness import ModelHarness


class PerformanceTester(ModelHarness):
    @torch.no_grad()
    def measure_component_times(self, image: torch.Tensor, num_runs: int = 100):
        """Per-component milliseconds (mean / std / min / max / median), same keys as the reference:
        backbone, selector, selector_nms, refiner, total."""
        keys = ["backbone", "selector", "selector_nms", "refiner", "total"]
        times = {k: [] for k in keys}
        for _ in range(10):
            self.forward_pass(image)
        K = self.config["model"]["num_keypoints"]
        marks = []
        for _ in range(num_runs):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            dino_features = self._features(image)
            ev[1].record()
            saliency_map = self.selector(dino_features)
            ev[2].record()
            keypoints_patch, scores = self.selector.select_keypoints(saliency_map, num_keypoints=K)
            ev[3].record()
            feat_at_kpts = self.backbone.extract_at_keypoints(dino_features, keypoints_patch)
            ev[4].record()
            self.refiner(feat_at_kpts)
            ev[5].record()
            marks.append(ev)
        torch.cuda.synchronize()
        for ev in marks:
            times["backbone"].append(ev[0].elapsed_time(ev[1]))
            times["selector"].append(ev[1].elapsed_time(ev[2]))
            times["selector_nms"].append(ev[2].elapsed_time(ev[3]))
            times["refiner"].append(ev[4].elapsed_time(ev[5]))
            times["total"].append(ev[0].elapsed_time(ev[5]))
        return {k: {"mean": np.mean(v), "std": np.std(v), "min": np.min(v), "max": np.max(v),
                    "median": np.median(v)} for k, v in times.items()}
