"""Batched, device-resident extract + match (the hot path as one object).

Mirrors the orchestration of ``MatchVisualizer.extract_features`` /
``SequenceMatcher.extract`` (visualize_matches.py:70-100, visualize_matches_sequence.py:69-104
there) without the per-frame ``.cpu().numpy()`` and without re-extracting shared frames
(visualize_matches_sequence.py:306-307): every frame is decoded once, its descriptors stay in HBM
and consecutive (or listed) pairs are matched from the resident bank.

Two coordinate conventions:
  * ``grid="pixel"`` — pipeline P: saliency at image resolution (H, W), features at (H/16, W/16);
    keypoints are pixel coordinates and ``pixel_to_patch`` is fused into the sampler.
  * ``grid="patch"`` — the reference-native case: saliency and features on the same patch grid;
    keypoints are patch coordinates, ``patch_to_pixel`` gives the pixel output.
"""

import torch

from . import matchers, ops
from .ops import SIM_BF16, SIM_F32


class FrontEnd:
    def __init__(self, refiner, num_keypoints=2048, nms_radius=2, min_score_percentile=0.5,
                 grid="pixel", sim_mode=SIM_F32, patch_size=16):
        self.refiner = refiner.eval()
        self.K, self.r, self.pct = int(num_keypoints), int(nms_radius), float(min_score_percentile)
        self.grid, self.sim_mode, self.patch = grid, sim_mode, patch_size

    @torch.no_grad()
    def extract(self, saliency, features, out=None):
        """saliency (B,H,W,1)|(B,H,W), features (B,h,w,C) -> dict of device tensors:
        keypoints_pixel (B,K,2), scores (B,K), descriptors (B,K,D) [, descriptors_bf16], info."""
        kp, sc, info = ops.decode_topk(saliency, self.K, self.r, self.pct)
        sampled = ops.gather_bilinear(features, kp, pixel_coords=(self.grid == "pixel"))
        raw = self.refiner.forward_unnormalized(sampled)
        B = kp.shape[0]
        res = dict(scores=sc, info=info)
        if self.sim_mode == SIM_BF16:
            d32, d16 = ops.l2norm_rows(raw, want_bf16=True)
            res["descriptors_bf16"] = d16.reshape(B, self.K, -1)
        else:
            d32 = ops.l2norm_rows(raw)
        res["descriptors"] = d32.reshape(B, self.K, -1)
        res["keypoints_pixel"] = kp if self.grid == "pixel" else kp * self.patch + self.patch / 2
        return res

    def bank(self, feats):
        return feats["descriptors_bf16"] if self.sim_mode == SIM_BF16 else feats["descriptors"]

    @torch.no_grad()
    def match_consecutive(self, feats, variant=matchers.M1, **kw):
        """Match frame t with frame t+1 for every t of an extracted batch (P = B-1 pairs)."""
        bank = self.bank(feats)
        F = bank.shape[0]
        sc = feats["scores"]
        return matchers.match(bank, bank[1:], variant, num_pairs=F - 1, mode=self.sim_mode,
                              scores1=sc, scores2=sc[1:], **kw)[:3]

    @torch.no_grad()
    def match_pairs(self, feats, pair_index, variant=matchers.M1, **kw):
        """Match the listed (a, b) frame pairs of an extracted batch (loop-closure style)."""
        bank = self.bank(feats)
        sc = feats["scores"]
        return matchers.match(bank, bank, variant, pair_index=pair_index, mode=self.sim_mode,
                              scores1=sc, scores2=sc, **kw)[:3]

    @torch.no_grad()
    def run_sequence(self, saliency, features, variant=matchers.M1, chunk=64, **kw):
        """Extract every frame once (in chunks) and match consecutive pairs.  Returns padded pair
        lists for the T-1 pairs, all on device."""
        T = saliency.shape[0]
        descs, descs16, scores, kps = [], [], [], []
        for s in range(0, T, chunk):
            f = self.extract(saliency[s:s + chunk], features[s:s + chunk])
            descs.append(f["descriptors"]); scores.append(f["scores"]); kps.append(f["keypoints_pixel"])
            if self.sim_mode == SIM_BF16:
                descs16.append(f["descriptors_bf16"])
        feats = dict(descriptors=torch.cat(descs), scores=torch.cat(scores),
                     keypoints_pixel=torch.cat(kps))
        if descs16:
            feats["descriptors_bf16"] = torch.cat(descs16)
        pairs, pscores, counts = self.match_consecutive(feats, variant, **kw)
        return feats, pairs, pscores, counts
