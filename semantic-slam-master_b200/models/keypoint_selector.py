"""KeypointSelector with a CUDA decode.

Same surface as the reference class (models/keypoint_selector.py:15-226 there): constructor
arguments, ``state_dict`` keys (``conv.0.*``, ``conv.2.*``), ``forward``, ``select_keypoints`` and
``_apply_nms`` keep their names, keywords, defaults, shapes and dtypes, so a reference checkpoint
loads unchanged and callers need no edits.  The saliency head (3x3 conv -> ReLU -> 1x1 conv ->
sigmoid) is one tcgen05 implicit-GEMM kernel under ``torch.no_grad()`` (``sslam_selector_head_f32``;
PyTorch layers when autograd is recording); ``select_keypoints`` — percentile threshold, NMS, candidate compaction,
top-k and the fallback branches — is one batched, sync-free call into ``sslam_decode_topk_f32``.
"""

from typing import Tuple

import torch
import torch.nn as nn

from sslam_b200 import ops


class KeypointSelector(nn.Module):
    def __init__(self, input_dim: int = 384, hidden_dim: int = 128):
        super().__init__()
        head = [nn.Conv2d(input_dim, hidden_dim, kernel_size=3, padding=1),
                nn.ReLU(inplace=True),
                nn.Conv2d(hidden_dim, 1, kernel_size=1)]
        self.conv = nn.Sequential(*head)
        for layer in (head[0], head[2]):                 # reference init: xavier(gain .5), zero bias
            nn.init.xavier_uniform_(layer.weight, gain=0.5)
            nn.init.constant_(layer.bias, 0.0)

    head = "tcgen05"         # "tcgen05" (default) or "torch" (cuDNN convolutions, for A/B comparisons)

    def _plan(self):
        ps = [self.conv[0].weight, self.conv[0].bias, self.conv[2].weight, self.conv[2].bias]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_plan_key", None) != key:
            self._plan_obj = ops.SelectorPlan(*ps)
            self._plan_key = key
        return self._plan_obj

    def _kernel_ok(self, x):
        c0, c2 = self.conv[0], self.conv[2]
        return (self.head == "tcgen05" and x.is_cuda and x.dtype == torch.float32
                and c0.in_channels % 64 == 0 and c0.out_channels % 8 == 0 and c0.out_channels <= 512
                and c2.out_channels == 1)

    def forward(self, dino_features: torch.Tensor, return_logits: bool = False) -> torch.Tensor:
        """(B, H, W, C) patch features -> (B, H, W, 1) sigmoid saliency.  Without autograd on a CUDA
        device the whole head (3x3 conv, ReLU, 1x1 conv, sigmoid) is one tcgen05 implicit-GEMM kernel
        (sslam_selector_head_f32); when gradients are needed the PyTorch layers run."""
        grad = torch.is_grad_enabled() and (dino_features.requires_grad or
                                            any(p.requires_grad for p in self.parameters()))
        if not grad and self._kernel_ok(dino_features):
            return ops.selector_head(self._plan(), dino_features, apply_sigmoid=not return_logits).unsqueeze(-1)
        logits = self.conv(dino_features.permute(0, 3, 1, 2))
        return (logits if return_logits else torch.sigmoid(logits)).permute(0, 2, 3, 1)

    def select_keypoints(self, saliency_map: torch.Tensor, num_keypoints: int = 500,
                         nms_radius: int = 2, min_score_percentile: float = 0.50
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
        """(B, H, W, 1) saliency -> keypoints (B, K, 2) fp32 (x, y) in the map's own grid, and
        scores (B, K) fp32.  Ordering and fallbacks follow the reference exactly; equal scores are
        ordered by ascending y*W+x (torch.topk leaves that unspecified)."""
        kp, sc, info = ops.decode_topk(saliency_map, num_keypoints, nms_radius=nms_radius,
                                       min_score_percentile=min_score_percentile)
        B, H, W = saliency_map.shape[0], saliency_map.shape[1], saliency_map.shape[2]
        if num_keypoints > H * W:
            # only then can a fallback ask topk for more elements than exist; the reference raises
            # (host sync on this rare path only)
            if bool((info[:, 0] < 0).any()):
                raise RuntimeError("selected index k out of range")
        return kp, sc

    def select_keypoints_with_info(self, saliency_map, num_keypoints=500, nms_radius=2,
                                   min_score_percentile=0.50):
        """As select_keypoints, plus the (B, 4) int32 info tensor (branch, candidates, ties,
        local maxima) — additive API, stays on device."""
        return ops.decode_topk(saliency_map, num_keypoints, nms_radius=nms_radius,
                               min_score_percentile=min_score_percentile)

    def select_keypoints_from_cells(self, cell_logits: torch.Tensor, num_keypoints: int = 500,
                                    nms_radius: int = 2, min_score_percentile: float = 0.50,
                                    cell: int = 8, border: int = 4):
        """OPTIONAL decode mode, off by default (the reference code has no such mode; see
        csrc/heatmap.cu): (B, cell*cell+1, Hc, Wc) cell logits -> channel softmax, dustbin dropped,
        depth-to-space, border mask -> then the same NMS / percentile threshold / top-k as
        ``select_keypoints`` on the (B, Hc*cell, Wc*cell) heatmap.  Returns (keypoints, scores, heatmap)."""
        heat = ops.heatmap_from_cells(cell_logits, cell=cell, border=border)
        kp, sc, _ = ops.decode_topk(heat, num_keypoints, nms_radius=nms_radius,
                                    min_score_percentile=min_score_percentile)
        return kp, sc, heat

    def _apply_nms(self, saliency: torch.Tensor, radius: int) -> torch.Tensor:
        """(B, H, W) -> (B, H, W): keep values equal to their (2r+1)^2 neighbourhood maximum."""
        if radius == 0:
            return saliency
        return ops.nms(saliency, radius)
