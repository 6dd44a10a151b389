"""Platform-independent input recipes for the golden fixtures (test infrastructure).

Full-size inputs (a 480x640 saliency map is 1.2 MB) are not committed; they are regenerated
from integer-only arithmetic so that every machine produces the same bytes, and the fixture
stores the SHA-256 of the input next to the reference's outputs.  Only NumPy integer streams
(PCG64), exact integer box sums and one correctly-rounded division are used — no
transcendental functions, no BLAS.
"""

import hashlib

import numpy as np

F32 = np.float32


def sha256(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def box_saliency(H, W, seed, bits=20, quant=None):
    """Smooth saliency map in (0, 1): 5x5 box sums of iid integers, scaled.

    ``quant`` (e.g. 256) rounds to that many levels to force plateaus and k-th boundary ties.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.integers(0, 1 << bits, size=(H + 4, W + 4), dtype=np.int64)
    c = np.zeros((H + 5, W + 5), dtype=np.int64)
    c[1:, 1:] = u.cumsum(0).cumsum(1)
    box = c[5:, 5:] - c[:-5, 5:] - c[5:, :-5] + c[:-5, :-5]          # (H, W) exact
    if quant is not None:
        box = (box * quant) // (25 << bits)
        return ((box.astype(np.float64) + 0.5) / quant).astype(F32)
    return ((box.astype(np.float64) + 0.5) / float(25 << bits)).astype(F32)


def spread_saliency(H, W, seed, lo=0.02, hi=0.98):
    """box_saliency stretched to roughly [lo, hi] with exact fp64 ops (used for maps whose
    median must sit well above the 0.1 floor and whose maxima approach 1)."""
    s = box_saliency(H, W, seed).astype(np.float64)
    z = (s - 0.5) / 0.0577 / 3.5                                     # ~N(0,1)/3.5
    z = np.clip(z, -1.0, 1.0)
    return (0.5 * (lo + hi) + 0.5 * (hi - lo) * z).astype(F32)


def int_features(B, h, w, C, seed):
    """NHWC feature maps with entries k/64, k in [-256, 256]: exactly representable."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.integers(-256, 257, size=(B, h, w, C)).astype(F32) / F32(64)).astype(F32)


def pixel_keypoints(B, N, H, W, seed):
    """Integer pixel keypoints (x, y) as fp32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.integers(0, W, size=(B, N))
    y = rng.integers(0, H, size=(B, N))
    return np.stack([x, y], axis=-1).astype(F32)


def _normalise_int_rows(v):
    nrm = np.sqrt((v * v).sum(1, keepdims=True))          # exact integer sum, rounded sqrt
    nrm[nrm == 0] = 1.0
    return (v / nrm).astype(F32)


def descriptor_pair(n, m, d, seed, noise=1, dup_every=0, near_dup_every=0):
    """Two unit-norm descriptor sets built from small integers (exact sums of squares,
    correctly rounded sqrt/div, so the bytes are identical on every machine).

    Set 2 holds a permuted, integer-perturbed copy of min(n, m) rows of set 1 (plus fresh rows
    when m > n), so most rows have a mutual nearest neighbour.  ``dup_every`` > 0 copies every
    such row of set 1 onto its successor to force exact argmax ties; ``near_dup_every`` > 0 makes
    every such row a copy of its predecessor with one coordinate nudged, so that second-best
    similarities sit just below the best and the ratio tests reject.
    Returns (desc1 (n, d), desc2 (m, d), perm).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    v = rng.integers(-8, 9, size=(n, d)).astype(np.float64)
    v[:, 0] += (np.abs(v).sum(1) == 0)                                # no zero rows
    if dup_every:
        for r in range(dup_every, n, dup_every):
            v[r] = v[r - 1]
    if near_dup_every:
        for r in range(near_dup_every, n, near_dup_every):
            v[r] = v[r - 1]
            v[r, r % d] += 3.0
    perm = rng.permutation(n)
    k = min(n, m)
    w = v[perm[:k]] * 4.0 + rng.integers(-noise, noise + 1, size=(k, d))
    if m > n:
        w = np.concatenate([w, rng.integers(-8, 9, size=(m - n, d)).astype(np.float64)], 0)
    return _normalise_int_rows(v), _normalise_int_rows(w), perm
