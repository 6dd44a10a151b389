"""Oracle: mutual-nearest-neighbour descriptor matching, variants M1-M5.

Test infrastructure (see oracle/__init__.py).  One primitive — per row (argmax, best,
second-best) and per column (argmax, best) of S = D1 @ D2^T — serves the five acceptance rules
the reference has:

  M1  MatchVisualizer.find_matches                       visualize_matches.py:102-124
  M2  SequenceMatcher.match_with_quality                 visualize_matches_sequence.py:106-197
  M3  DescriptorQualityTester.find_mutual_nearest_neighbors   test/test_descriptor_quality.py:97-142
  M4  SemanticSLAMTrainer._find_matches                  train.py:410-449
  M5  tracking count in TrackingTester                   test/test_tracking.py:159-161

Scalars: the reference compares fp32 arrays with Python floats; under NumPy 2 (NEP 50) and
torch type promotion the Python float is cast to fp32 first, which is what is restated here.
"""

import numpy as np

F32 = np.float32


def similarity_top2(desc1, desc2):
    """Row top-2 and column top-1 of S = desc1 @ desc2.T (fp32).

    Returns nn12 (N,) int64, best12 (N,), second12 (N,), nn21 (M,) int64, best21 (M,), S.
    ``second12`` is the second entry of the row sorted descending (duplicates of the maximum
    count, as in ``np.sort(...)[:, ::-1][:, 1]``, test_descriptor_quality.py:129-130, and as in
    the masked-row maximum of visualize_matches.py:117-119); -inf when M == 1.
    argmax returns the lowest maximal index (visualize_matches.py:108-109).
    """
    d1 = np.asarray(desc1, dtype=F32)
    d2 = np.asarray(desc2, dtype=F32)
    S = np.dot(d1, d2.T).astype(F32)                                 # visualize_matches.py:105
    N, M = S.shape
    nn12 = S.argmax(axis=1)
    nn21 = S.argmax(axis=0)
    best12 = S[np.arange(N), nn12]
    best21 = S[nn21, np.arange(M)]
    if M >= 2:
        second12 = np.partition(S, M - 2, axis=1)[:, M - 2]
    else:
        second12 = np.full(N, -np.inf, dtype=F32)
    return nn12, best12, second12, nn21, best21, S


def match_m1(desc1, desc2, ratio_thresh=0.8, top=None, promotion="nep50"):
    """``MatchVisualizer.find_matches`` (visualize_matches.py:102-124): list of (i, j, sim),
    ascending i; mutual and ``sim > second_best * ratio_thresh`` where the best column is
    replaced by -1 before taking the second maximum (:117-119).

    ``promotion``: the comparison at :121 multiplies a NumPy fp32 *scalar* by a Python float.  NumPy >= 2
    (NEP 50; this container, and what the golden fixtures were written under) casts the float to fp32
    ("nep50"); NumPy < 2, which the reference's requirements.txt:1 pins, promotes the product to
    float64 ("legacy").  Rows within one fp32 ulp of the threshold can differ between the two."""
    nn12, best12, second12, nn21, _, _ = top if top is not None else similarity_top2(desc1, desc2)
    out = []
    rt = F32(ratio_thresh)
    for i in range(nn12.shape[0]):
        j = int(nn12[i])
        if nn21[j] == i:                                             # :114
            second = max(F32(second12[i]), F32(-1))                  # :118-119
            if promotion == "legacy":
                ok = float(best12[i]) > float(second) * float(ratio_thresh)
            else:
                ok = best12[i] > F32(second * rt)
            if ok:                                                   # :121
                out.append((i, j, F32(best12[i])))
    return out


def match_m2(desc1, desc2, scores1, scores2, saliency_weight=0.3, min_saliency=0.2,
             min_descriptor_sim=0.7, intensity1=None, intensity2=None, min_intensity=0.1,
             top=None):
    """``SequenceMatcher.match_with_quality`` (visualize_matches_sequence.py:106-197).

    Returns matches (K', 2) int64 ascending i and quality (K',) fp32; empty results are
    ``zeros((0, 2), int64)``, ``zeros((0,), float32)`` (:154-155, 178-180).
    """
    nn12, best12, _, nn21, _, _ = top if top is not None else similarity_top2(desc1, desc2)
    s1 = np.asarray(scores1, dtype=F32)
    s2 = np.asarray(scores2, dtype=F32)
    N = nn12.shape[0]
    idx1 = np.nonzero(nn21[nn12] == np.arange(N))[0]                 # :147-150
    idx2 = nn12[idx1]
    empty = (np.zeros((0, 2), dtype=np.int64), np.zeros((0,), dtype=F32))
    if idx1.size == 0:
        return empty
    sim = best12[idx1]                                               # :158
    avg_sal = ((s1[idx1] + s2[idx2]) / F32(2)).astype(F32)           # :161-163
    valid = (avg_sal >= F32(min_saliency)) & (sim >= F32(min_descriptor_sim))   # :166-168
    if intensity1 is not None and intensity2 is not None:            # :171-176
        i1 = np.asarray(intensity1, dtype=F32)[idx1]
        i2 = np.asarray(intensity2, dtype=F32)[idx2]
        valid &= (((i1 + i2) / F32(2)).astype(F32) >= F32(min_intensity))
    if valid.sum() == 0:
        return empty
    idx1, idx2, sim, avg_sal = idx1[valid], idx2[valid], sim[valid], avg_sal[valid]
    quality = ((F32(1 - saliency_weight) * sim).astype(F32)
               + (F32(saliency_weight) * avg_sal).astype(F32)).astype(F32)      # :189-192
    return np.stack([idx1, idx2], axis=1).astype(np.int64), quality


def match_m3(desc1, desc2, ratio_threshold=0.9, top=None):
    """``find_mutual_nearest_neighbors`` (test/test_descriptor_quality.py:97-142): mutual and
    ``second/(best + 1e-8) < ratio_threshold``; returns matches (K', 2) and ``1 - best``."""
    nn12, best12, second12, nn21, _, _ = top if top is not None else similarity_top2(desc1, desc2)
    N = nn12.shape[0]
    mutual = nn21[nn12] == np.arange(N)                              # :126
    ratio = (second12 / (best12 + F32(1e-8)).astype(F32)).astype(F32)   # :129-130
    valid = mutual & (ratio < F32(ratio_threshold))                  # :131-134
    idx1 = np.where(valid)[0]
    idx2 = nn12[idx1]
    return (np.stack([idx1, idx2], axis=1).astype(np.int64),
            (F32(1.0) - best12[idx1]).astype(F32))                   # :139-140


def match_m4(desc1, desc2):
    """``SemanticSLAMTrainer._find_matches`` (train.py:410-449): batched mutual NN,
    (B, maxM, 2) int64 padded with (0, 0) rows; all-empty -> zeros (B, 1, 2) (:440)."""
    d1 = np.asarray(desc1, dtype=F32)
    d2 = np.asarray(desc2, dtype=F32)
    per = []
    for b in range(d1.shape[0]):                                     # :419
        nn12, _, _, nn21, _, _ = similarity_top2(d1[b], d2[b])
        idx1 = np.nonzero(nn21[nn12] == np.arange(nn12.shape[0]))[0]
        per.append(np.stack([idx1, nn12[idx1]], axis=1).astype(np.int64))
    mx = max(m.shape[0] for m in per)                                # :438
    if mx == 0:
        return np.zeros((d1.shape[0], 1, 2), dtype=np.int64)
    out = np.zeros((d1.shape[0], mx, 2), dtype=np.int64)             # :442-447
    for b, m in enumerate(per):
        out[b, :m.shape[0]] = m
    return out


def match_m5(desc_prev, desc_curr, match_threshold=0.8, top=None):
    """Tracking count (test/test_tracking.py:159-161): number of rows whose maximum
    similarity exceeds the threshold; no mutual check."""
    _, best12, _, _, _, _ = top if top is not None else similarity_top2(desc_prev, desc_curr)
    return int((best12 > F32(match_threshold)).sum())
