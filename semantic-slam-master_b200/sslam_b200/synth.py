"""Seeded synthetic inputs shaped like the reference's data.

There is no TUM RGB-D data, no DINOv3 checkpoint and no network, so every benchmark and parity
test feeds the hot path the tensors it would see *after* the backbone and the selector head:
a saliency map (B, H, W, 1) fp32 and an NHWC feature map (B, H/16, W/16, 384) fp32.  A sequence
is a crop window sliding over a fixed "world canvas" (TUM fr1/desk-like small motion: one patch
every 8 frames), so consecutive frames overlap and mutual matches exist.

All randomness comes from ``torch.Generator`` objects with the seeds listed in SURVEY.md §8(d);
identical tensors are handed to the oracle and to the kernels.
"""

import torch
import torch.nn.functional as F

PATCH = 16
MARGIN = 600
FEAT_DIM = 384


class WorldCanvas:
    """Smoothed-noise logits canvas + iid feature canvas for one synthetic sequence."""

    def __init__(self, seq_id=0, height=480, width=640, feat_dim=FEAT_DIM, device="cpu"):
        assert height % PATCH == 0 and width % PATCH == 0
        self.seq_id, self.H, self.W, self.C = int(seq_id), int(height), int(width), int(feat_dim)
        self.device = torch.device(device)
        g = torch.Generator(device="cpu").manual_seed(1000 + self.seq_id)
        ch, cw = self.H + 2 * MARGIN, self.W + 2 * MARGIN
        raw = torch.randn(ch, cw, generator=g)
        smooth = F.avg_pool2d(raw[None, None], kernel_size=5, stride=1, padding=2,
                              count_include_pad=False)[0, 0]
        self.logits = (smooth * 4.0).to(self.device)
        self.feats = torch.randn(ch // PATCH, cw // PATCH, self.C, generator=g).to(self.device)

    def frame(self, t):
        """Return (logits (H, W), features (H/16, W/16, C)) of frame ``t`` on ``self.device``."""
        ox = (PATCH * (t // 8)) % (2 * MARGIN)
        oy = 0
        gen_dev = "cuda" if self.device.type == "cuda" else "cpu"
        g = torch.Generator(device=gen_dev).manual_seed(2000 + int(t) + 100003 * self.seq_id)
        lg = self.logits[oy:oy + self.H, ox:ox + self.W]
        lg = lg + 0.05 * torch.randn(self.H, self.W, generator=g, device=self.device)
        h, w = self.H // PATCH, self.W // PATCH
        ft = self.feats[oy // PATCH:oy // PATCH + h, ox // PATCH:ox // PATCH + w]
        ft = ft + 0.02 * torch.randn(h, w, self.C, generator=g, device=self.device)
        return lg.contiguous(), ft.contiguous()


def make_sequence(num_frames, seq_id=0, height=480, width=640, feat_dim=FEAT_DIM,
                  device="cpu", start=0, stride=1, out_saliency=None, out_features=None):
    """Frames ``start, start+stride, ...`` of sequence ``seq_id``.

    Returns saliency (T, H, W, 1) fp32 — ``torch.sigmoid`` of the logits, computed once on
    ``device`` and shared by every implementation — and features (T, H/16, W/16, C) fp32 NHWC.
    Pre-allocated (e.g. pinned) outputs may be supplied.
    """
    canvas = WorldCanvas(seq_id, height, width, feat_dim, device)
    h, w = height // PATCH, width // PATCH
    sal = out_saliency if out_saliency is not None else torch.empty(
        num_frames, height, width, 1, device=canvas.device)
    feat = out_features if out_features is not None else torch.empty(
        num_frames, h, w, feat_dim, device=canvas.device)
    for i in range(num_frames):
        lg, ft = canvas.frame(start + i * stride)
        sal[i, :, :, 0].copy_(torch.sigmoid(lg))
        feat[i].copy_(ft)
    return sal, feat


def make_pairs(num_pairs, height=480, width=640, feat_dim=FEAT_DIM, device="cpu"):
    """c3-style independent pairs: frames (2p, 2p+1) of sequences p = 0..num_pairs-1.
    Returns saliency (P, 2, H, W, 1) and features (P, 2, h, w, C)."""
    sals, feats = [], []
    for p in range(num_pairs):
        s, f = make_sequence(2, seq_id=p, height=height, width=width, feat_dim=feat_dim,
                             device=device, start=2 * p)
        sals.append(s)
        feats.append(f)
    return torch.stack(sals, 0), torch.stack(feats, 0)


def native_grid_case(batch=4, grid=28, feat_dim=FEAT_DIM, seed=7):
    """c0: the reference-native 28x28 patch grid — iid features; the saliency comes from the
    selector head with ``torch.manual_seed(0)`` weights (made by the caller)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(batch, grid, grid, feat_dim, generator=g)


def noisy_permutation_descriptors(n, d=256, noise=0.05, seed=0, m=None):
    """Unit-norm descriptor sets (n, d) and (m, d) where set 2 is a noisy permutation of set 1
    (plus fresh rows if m > n); used by matcher tests."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    m = n if m is None else m
    a = F.normalize(torch.randn(n, d, generator=g), dim=-1)
    perm = torch.randperm(n, generator=g)
    b = a[perm][:min(n, m)] + noise * torch.randn(min(n, m), d, generator=g)
    if m > n:
        b = torch.cat([b, torch.randn(m - n, d, generator=g)], 0)
    b = F.normalize(b, dim=-1)
    return a.contiguous(), b.contiguous(), perm
