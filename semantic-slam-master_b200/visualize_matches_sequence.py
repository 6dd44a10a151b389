"""Drop-in for the matching surface of the reference's ``visualize_matches_sequence.py``.

``SequenceMatcher.extract`` (visualize_matches_sequence.py:69-104 there),
``SequenceMatcher.match_with_quality`` (a ``@staticmethod``, :106-197) and the module-level
``process_spacing`` (:272-357) keep their names, arguments, defaults and return types: matches
(K', 2) int64 in ascending i and quality (K',) fp32, empty arrays when nothing passes.  Drawing
(``visualize_matches`` there, matplotlib) is presentation only and out of scope; where the
reference saves a PNG per pair, ``process_spacing`` here saves the pair's match list in the on-disk
format of ``sslam_b200.evaluation.write_match_lists``.

Additive, device-resident drivers: ``SequenceMatcher.process_spacings_device`` extracts every frame
once into an HBM frame store and matches all spacings from it.
"""

from pathlib import Path
from typing import Optional

import numpy as np
import torch

from sslam_b200 import evaluation, matchers
from sslam_b200.framestore import FrameStore
from sslam_b200.pipeline import FrontEnd
from visualize_matches import MatchVisualizer


class SequenceMatcher(MatchVisualizer):
    @staticmethod
    def _intensity(gray, keypoints_pixel):
        """Per-keypoint grey value at the rounded pixel (visualize_matches_sequence.py:91-94)."""
        xs = np.clip(keypoints_pixel[:, 0].round().astype(int), 0, gray.shape[1] - 1)
        ys = np.clip(keypoints_pixel[:, 1].round().astype(int), 0, gray.shape[0] - 1)
        return gray[ys, xs]

    @torch.no_grad()
    def extract_from_patch_map(self, dino_features, gray=None, image=None):
        """``extract`` after the backbone: (1,h,w,C) patch features [+ the resized grey image (H,W)
        float32 in 0..1] -> the dict ``extract`` returns (saliency (h,w), keypoints_pixel (K,2),
        scores (K,), intensity (K,), descriptors (K,D) [, image])."""
        f = self._extract_device(dino_features)
        kpts = f["keypoints_pixel"][0].cpu().numpy()
        out = {"image": image, "saliency": f["saliency"][0, :, :, 0].cpu().numpy(), "keypoints_pixel": kpts,
               "scores": f["scores"][0].cpu().numpy(),
               "intensity": self._intensity(gray, kpts) if gray is not None else None,
               "descriptors": f["descriptors"][0].cpu().numpy()}
        return out

    @torch.no_grad()
    def extract(self, image_path: str) -> dict:
        """visualize_matches_sequence.py:69-104 there (needs the ViT: ``timm`` + DINOv3 weights)."""
        from PIL import Image
        from torchvision import transforms
        size = self.config["model"]["input_size"]
        tf = transforms.Compose([transforms.Resize((size, size)), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        image = Image.open(image_path).convert("RGB")
        patch_features = self.backbone(tf(image).unsqueeze(0).to(self.device))
        gray = np.array(image.resize((size, size)).convert("L"), dtype=np.float32) / 255.0
        return self.extract_from_patch_map(patch_features, gray=gray, image=image)

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
    def process_spacings_device(self, saliency, features, spacings=(1, 5, 10, 15, 20), max_pairs=None,
                                num_keypoints=None, chunk=64, **thresholds):
        """Device-resident form of the ``for spacing in args.spacings: process_spacing(...)`` loop
        (:398-411 there) on pixel-resolution saliency maps (T,H,W,1) and feature maps (T,h,w,C) that
        are already past the backbone: frame i against frame i + spacing for i = 0, spacing,
        2*spacing, ... (:297), every frame extracted ONCE into an HBM frame store, all spacings
        matched from the resident descriptors with M2.  Returns {spacing: (pair_index (P,2) frame
        ids, pairs, quality, counts)} with padded device lists."""
        K = num_keypoints or self.config["model"]["num_keypoints"]
        T = saliency.shape[0]
        fe = FrontEnd(self.refiner, num_keypoints=K, grid="pixel")
        store = FrameStore(fe, capacity=T)
        for s in range(0, T, chunk):
            store.push(saliency[s:s + chunk], features[s:s + chunk])
        out = {}
        for sp in spacings:
            fp = [(i, i + sp) for i in range(0, T - sp, sp)]
            if max_pairs is not None:
                fp = fp[:max_pairs]
            if not fp:
                continue
            pairs, quality, counts = store.match_pairs(fp, matchers.M2, **thresholds)
            out[sp] = (torch.tensor(fp, dtype=torch.int32), pairs, quality, counts)
        return out


def visualize_matches(*args, output_path=None, **kwargs):
    """Drawing is out of scope in this build: nothing is rendered."""
    return None


def process_spacing(matcher: SequenceMatcher, images: list, spacing: int, output_dir: Path, max_pairs: int,
                    max_matches: int, gap: int, saliency_weight: float, min_saliency: float,
                    min_descriptor_sim: float, min_intensity: float):
    """Process all pairs for a given spacing (visualize_matches_sequence.py:272-357 there): pair
    (images[i], images[i + spacing]) for i = 0, spacing, ... up to ``max_pairs`` pairs, extract both,
    ``match_with_quality`` with the intensity filter, collect the quality scores, print the summary,
    return the list of all quality scores.

    ``images`` holds image paths (``matcher.extract``; needs the ViT) or, for work that starts after
    the backbone, ready feature dicts (as ``extract`` / ``extract_from_patch_map`` return them) or
    (patch_features (1,h,w,C) tensor, gray (H,W) array or None) tuples.  Each distinct frame is
    extracted once (the reference extracts both frames of every pair again, :306-307)."""
    print(f"\n{'=' * 70}\nProcessing spacing={spacing} frames\n{'=' * 70}")
    spacing_dir = Path(output_dir) / f"spacing_{spacing}"
    spacing_dir.mkdir(parents=True, exist_ok=True)
    cache = {}

    def feats(i):
        if i not in cache:
            item = images[i]
            if isinstance(item, dict):
                cache[i] = item
            elif isinstance(item, tuple):
                cache[i] = matcher.extract_from_patch_map(item[0], gray=item[1])
            else:
                cache[i] = matcher.extract(str(item))
        return cache[i]

    def name(i):
        item = images[i]
        return Path(str(item)).stem if not isinstance(item, (dict, tuple)) else f"frame{i:06d}"

    pair_count = 0
    all_quality_scores = []
    for i in range(0, len(images) - spacing, spacing):
        if pair_count >= max_pairs:
            break
        f1, f2 = feats(i), feats(i + spacing)
        print(f"\nPair {pair_count + 1}/{max_pairs}: {name(i)} → {name(i + spacing)}")
        matches, match_quality = matcher.match_with_quality(
            f1["descriptors"], f2["descriptors"], f1["scores"], f2["scores"], saliency_weight=saliency_weight,
            min_saliency=min_saliency, min_descriptor_sim=min_descriptor_sim, intensity1=f1["intensity"],
            intensity2=f2["intensity"], min_intensity=min_intensity)
        if len(matches) > 0:
            all_quality_scores.extend(match_quality.tolist())
        # where the reference draws matches_{a}_to_{b}.png, keep the list itself (best max_matches first)
        order = np.argsort(-match_quality, kind="stable")[:max_matches]
        n = len(order)
        evaluation.write_match_lists(
            str(spacing_dir / f"matches_{name(i)}_to_{name(i + spacing)}.npz"),
            torch.as_tensor(matches[order].astype(np.int32)).reshape(1, n, 2),
            torch.as_tensor(match_quality[order]).reshape(1, n), torch.tensor([n], dtype=torch.int32),
            pair_index=np.array([[i, i + spacing]]), meta={"spacing": spacing, "total_matches": int(len(matches))})
        pair_count += 1
    if all_quality_scores:
        print(f"\n{'=' * 70}\nSummary for spacing={spacing}\n{'=' * 70}")
        print(f"Total pairs processed: {pair_count}")
        print(f"Total matches: {len(all_quality_scores)}")
        print(f"Average quality: {np.mean(all_quality_scores):.3f}")
        print(f"Quality range: [{np.min(all_quality_scores):.3f}, {np.max(all_quality_scores):.3f}]")
        print(f"High quality matches (>0.8): {sum(q > 0.8 for q in all_quality_scores)}")
        print(f"Output: {spacing_dir}")
    return all_quality_scores
