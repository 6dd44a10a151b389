"""Oracle: keypoint heatmap decode (test infrastructure, see oracle/__init__.py).

Restates ``KeypointSelector.select_keypoints`` / ``_apply_nms`` and the sigmoid tail of
``KeypointSelector.forward`` (models/keypoint_selector.py:45-67, 69-207, 209-226) in NumPy.
"""

from fractions import Fraction

import numpy as np

F32 = np.float32

BRANCH_MAIN = 0        # >= K candidates above threshold                 (keypoint_selector.py:120-128)
BRANCH_LOWER = 1       # 0 < n < K, a lower percentile supplied the rest  (:139-156)
BRANCH_RAW_PAD = 2     # 0 < n < K, padded from the raw map               (:157-173)
BRANCH_RAW_ALL = 3     # n == 0, top-K of the raw map                     (:174-184)

LOWER_PERCENTILES = (0.40, 0.30, 0.20, 0.10)   # keypoint_selector.py:139
MAIN_FLOOR = 0.1                               # :109
LOWER_FLOOR = 0.05                             # :141


def sigmoid_f32(logits):
    """fp32 logistic, the tail of ``KeypointSelector.forward`` (keypoint_selector.py:61-62)."""
    x = np.asarray(logits, dtype=F32)
    return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def _round_fraction_to_f32(x):
    """Exactly rounded (nearest-even) fp32 of a ``Fraction``."""
    c = F32(float(x))
    best = c
    for other in (np.nextafter(c, F32(-np.inf)), np.nextafter(c, F32(np.inf))):
        if not np.isfinite(other):
            continue
        d_best = abs(x - Fraction(float(best)))
        d_other = abs(x - Fraction(float(other)))
        if d_other < d_best:
            best = other
        elif d_other == d_best:
            # tie: even mantissa wins
            if (np.asarray(other, dtype=F32).view(np.uint32) & 1) == 0:
                best = other
    return F32(best)


def _fma_f32(a, b, c):
    """Single-rounding a*b+c in fp32 (what ATen's vectorised lerp emits on CPU)."""
    return _round_fraction_to_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def quantile_f32(flat, q):
    """``torch.quantile(flat, q)`` for a 1-D fp32 tensor (keypoint_selector.py:105-106).

    Observed ATen behaviour (torch 2.11.0 CPU; pinned by tests/golden/quantile.npz):
    full ascending sort; ``rank = fp32(q) * fp32(n-1)`` rounded to fp32; the two bracketing
    order statistics are blended with ``lerp`` whose CPU kernel is one FMA:
    ``w < 0.5 ? fma(w, b-a, a) : fma(w-1, b-a, b)``.
    """
    s = np.sort(np.asarray(flat, dtype=F32).ravel())
    n = s.size
    rank = F32(F32(q) * F32(n - 1))
    lo = int(np.floor(rank))
    hi = int(np.ceil(rank))
    w = F32(rank - F32(lo))
    a, b = s[lo], s[hi]
    d = F32(b - a)
    if abs(w) < F32(0.5):
        return _fma_f32(w, d, a)
    return _fma_f32(F32(w - F32(1)), d, b)


def _threshold(flat, q, floor):
    """``max(quantile.item(), floor)`` then compared against fp32 data (keypoint_selector.py:109,115).

    The Python-float threshold is cast to fp32 when compared with the fp32 map; for
    floor in {0.1, 0.05} that is exactly ``max(thr32, fp32(floor))``.
    """
    thr = max(float(quantile_f32(flat, q)), floor)
    return F32(thr)


def apply_nms(sal, radius):
    """``_apply_nms`` on one (H, W) map (keypoint_selector.py:209-226).

    (2r+1)^2 stride-1 max-pool with implicit -inf padding; every member of a plateau
    equals the pooled value and survives; everything else becomes 0.0.
    """
    sal = np.asarray(sal, dtype=F32)
    if radius == 0:
        return sal
    H, W = sal.shape
    r = int(radius)
    padded = np.full((H + 2 * r, W + 2 * r), -np.inf, dtype=F32)
    padded[r:r + H, r:r + W] = sal
    # separable running max
    hmax = padded[:, 0:W].copy()
    for dx in range(1, 2 * r + 1):
        np.maximum(hmax, padded[:, dx:dx + W], out=hmax)
    pooled = hmax[0:H].copy()
    for dy in range(1, 2 * r + 1):
        np.maximum(pooled, hmax[dy:dy + H], out=pooled)
    mask = (sal == pooled)
    return (sal * mask.astype(F32)).astype(F32)


def _topk(values, lin_index, k):
    """``torch.topk(values, k)`` with the tie order fixed to (score desc, linear index asc).

    Returns positions into ``values``.  Raises like ATen when k exceeds the length
    (keypoint_selector.py:123,148,166,178).
    """
    if k > values.shape[0]:
        raise RuntimeError("selected index k out of range")
    order = np.lexsort((lin_index, -values.astype(np.float64)))
    return order[:k]


def _decode_one(sal_b, K, nms_radius, pct):
    H, W = sal_b.shape
    flat = sal_b.ravel()
    lin_all = np.arange(H * W, dtype=np.int64)
    thr = _threshold(flat, pct, MAIN_FLOOR)                       # :105-109
    nms = apply_nms(sal_b, nms_radius)                            # :112
    nms_flat = nms.ravel()
    valid = nms_flat > thr                                        # :115
    cand_lin = lin_all[valid]                                     # row-major, :116
    cand_scr = nms_flat[valid]                                    # :117
    n = cand_lin.shape[0]
    ties = 0
    if n >= K:                                                    # :120-128
        sel = _topk(cand_scr, cand_lin, K)
        lin, scr = cand_lin[sel], cand_scr[sel]
        if K > 0:
            kth = scr[-1]
            ties = int((cand_scr == kth).sum() - (scr == kth).sum())
        branch = BRANCH_MAIN
    elif n > 0:                                                   # :130-173
        remaining = K - n                                         # :136
        lin, scr, branch = None, None, None
        for p in LOWER_PERCENTILES:                               # :139-156
            lthr = _threshold(flat, p, LOWER_FLOOR)
            extra = (nms_flat > lthr) & (~valid)
            e_lin, e_scr = lin_all[extra], nms_flat[extra]
            if e_lin.shape[0] >= remaining:
                sel = _topk(e_scr, e_lin, remaining)
                lin = np.concatenate([cand_lin, e_lin[sel]])
                scr = np.concatenate([cand_scr, e_scr[sel]])
                kth = e_scr[sel][-1]
                ties = int((e_scr == kth).sum() - (e_scr[sel] == kth).sum())
                branch = BRANCH_LOWER
                break
        if branch is None:                                        # :157-173 (for/else)
            sel = _topk(flat, lin_all, remaining)
            lin = np.concatenate([cand_lin, lin_all[sel]])
            scr = np.concatenate([cand_scr, flat[sel]])
            kth = flat[sel][-1]
            ties = int((flat == kth).sum() - (flat[sel] == kth).sum())
            branch = BRANCH_RAW_PAD
    else:                                                         # :174-184
        sel = _topk(flat, lin_all, K)
        lin, scr = lin_all[sel], flat[sel]
        if K > 0:
            kth = scr[-1]
            ties = int((flat == kth).sum() - (scr == kth).sum())
        branch = BRANCH_RAW_ALL
    # :186-199 — unreachable in practice (every branch above yields exactly K rows or raises),
    # restated for completeness.
    padded = 0
    if lin.shape[0] > K:
        lin, scr = lin[:K], scr[:K]
    elif lin.shape[0] < K:
        padded = K - lin.shape[0]
        best = int(np.argmax(scr))
        lin = np.concatenate([lin, np.full(padded, lin[best], dtype=lin.dtype)])
        scr = np.concatenate([scr, np.full(padded, scr[best], dtype=scr.dtype)])
    kp = np.stack([(lin % W).astype(F32), (lin // W).astype(F32)], axis=1)   # (x, y), :127
    return kp, scr.astype(F32), (branch, n, ties, padded)


def select_keypoints(saliency_map, num_keypoints=500, nms_radius=2, min_score_percentile=0.50):
    """``KeypointSelector.select_keypoints`` (keypoint_selector.py:69-207).

    Args mirror the reference; ``saliency_map`` is (B, H, W, 1) or (B, H, W) fp32.
    Returns ``keypoints`` (B, K, 2) fp32 in (x, y) order, ``scores`` (B, K) fp32 and an
    ``info`` (B, 4) int32 array: branch taken, candidates above the main threshold, ties at
    the k-th boundary left unselected (the only freedom ``torch.topk`` has), rows padded.
    """
    sal = np.asarray(saliency_map, dtype=F32)
    if sal.ndim == 4:
        sal = sal[..., 0]                                         # :95
    B = sal.shape[0]
    kps, scs, infos = [], [], []
    for b in range(B):                                            # :100
        kp, sc, info = _decode_one(sal[b], int(num_keypoints), int(nms_radius),
                                   float(min_score_percentile))
        kps.append(kp)
        scs.append(sc)
        infos.append(info)
    return (np.stack(kps, 0), np.stack(scs, 0), np.asarray(infos, dtype=np.int32))
