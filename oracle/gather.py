"""Oracle: bilinear descriptor sampling, patch<->pixel coordinates, L2 normalisation.

Test infrastructure (see oracle/__init__.py).  Restates
``DinoBackbone.extract_at_keypoints`` / ``patch_to_pixel`` / ``pixel_to_patch``
(models/dino_backbone.py:114-152, 154-165, 167-178) and the ``F.normalize`` tail of
``DescriptorRefiner.forward`` (models/descriptor_refiner.py:86).
"""

import numpy as np

F32 = np.float32
PATCH_SIZE = 16          # models/dino_backbone.py:35


def fma_f32(a, b, c):
    """Element-wise single-rounding ``a*b+c`` for fp32 arrays.

    The product of two fp32 numbers is exact in fp64; the fp64 sum is then rounded *to odd*
    (TwoSum error term decides), which makes the final fp64->fp32 rounding the correctly
    rounded FMA result.
    """
    a64 = np.asarray(a, dtype=F32).astype(np.float64)
    b64 = np.asarray(b, dtype=F32).astype(np.float64)
    c64 = np.asarray(c, dtype=F32).astype(np.float64)
    p = a64 * b64                      # exact
    s = p + c64
    bb = s - p                         # TwoSum
    err = (p - (s - bb)) + (c64 - bb)
    s, err = np.broadcast_arrays(s, err)
    s = s.copy()
    inexact = (err != 0) & np.isfinite(s)
    even = (s.view(np.int64) & 1) == 0
    fix = inexact & even
    if fix.any():
        toward = np.where(err > 0, np.inf, -np.inf)
        s[fix] = np.nextafter(s[fix], toward[fix])
    return s.astype(F32)


def patch_to_pixel(patch_coords):
    """``pix = p*16 + 8`` in fp32 (models/dino_backbone.py:164)."""
    p = np.asarray(patch_coords, dtype=F32)
    return (p * F32(PATCH_SIZE) + F32(PATCH_SIZE / 2)).astype(F32)


def pixel_to_patch(pixel_coords):
    """``p = (pix - 8)/16`` in fp32 (models/dino_backbone.py:177)."""
    p = np.asarray(pixel_coords, dtype=F32)
    return ((p - F32(PATCH_SIZE / 2)) / F32(PATCH_SIZE)).astype(F32)


def extract_at_keypoints(patch_features, keypoints):
    """``DinoBackbone.extract_at_keypoints`` (models/dino_backbone.py:114-152).

    patch_features (B, h, w, C) fp32 NHWC; keypoints (B, N, 2) fp32 (x, y) in patch units.
    Coordinates are normalised ``2x/(w-1)-1`` (:134-136) and un-normalised by
    ``grid_sample(align_corners=True)`` as ``((g+1)/2)*(w-1)`` (ATen GridSampler.h:27-31);
    taps outside the map contribute zero (padding_mode='zeros').  torch 2.11.0's CPU kernel
    accumulates ``nw*w_nw`` then FMAs the ne, sw, se taps in that order — reproduced here so
    the oracle is bit-identical to the reference on the pinned fixtures.
    """
    feat = np.asarray(patch_features, dtype=F32)
    kp = np.asarray(keypoints, dtype=F32)
    B, H, W, C = feat.shape
    N = kp.shape[1]
    x = kp[..., 0]
    y = kp[..., 1]
    nx = ((F32(2) * x) / F32(W - 1) - F32(1)).astype(F32)        # :135
    ny = ((F32(2) * y) / F32(H - 1) - F32(1)).astype(F32)        # :136
    ix = (((nx + F32(1)) / F32(2)) * F32(W - 1)).astype(F32)     # GridSampler.h:27-31
    iy = (((ny + F32(1)) / F32(2)) * F32(H - 1)).astype(F32)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + F32(1)
    y1 = y0 + F32(1)
    wx0 = (x1 - ix).astype(F32)
    wx1 = (ix - x0).astype(F32)
    wy0 = (y1 - iy).astype(F32)
    wy1 = (iy - y0).astype(F32)
    taps = ((x0, y0, (wx0 * wy0).astype(F32)), (x1, y0, (wx1 * wy0).astype(F32)),
            (x0, y1, (wx0 * wy1).astype(F32)), (x1, y1, (wx1 * wy1).astype(F32)))
    out = np.zeros((B, N, C), dtype=F32)
    for b in range(B):
        acc = None
        for (tx, ty, w) in taps:
            txi = tx[b].astype(np.int64)
            tyi = ty[b].astype(np.int64)
            inside = (txi >= 0) & (txi < W) & (tyi >= 0) & (tyi < H)
            v = feat[b][np.clip(tyi, 0, H - 1), np.clip(txi, 0, W - 1)]
            v = np.where(inside[:, None], v, F32(0)).astype(F32)
            wb = w[b][:, None]
            acc = (v * wb).astype(F32) if acc is None else fma_f32(v, wb, acc)
        out[b] = acc
    return out


def l2_normalize(x, eps=1e-12):
    """``F.normalize(x, p=2, dim=-1)`` = x / max(||x||_2, eps) (models/descriptor_refiner.py:86).

    The reduction order of the sum of squares is device specific, so consumers compare with a
    1e-6 relative tolerance.
    """
    x = np.asarray(x, dtype=F32)
    nrm = np.sqrt(np.sum(x.astype(np.float64) ** 2, axis=-1, keepdims=True)).astype(F32)
    return (x / np.maximum(nrm, F32(eps))).astype(F32)
