"""Oracle: the whole extract+match path on the CPU (test infrastructure / CPU baseline).

Composition "pipeline P" of SURVEY.md §8(d): select_keypoints on the pixel-resolution saliency map
-> pixel_to_patch -> extract_at_keypoints -> DescriptorRefiner (+ F.normalize) per frame, then a
matcher per consecutive pair — the order of ``MatchVisualizer.extract_features`` /
``SequenceMatcher.extract`` and ``process_spacing`` (visualize_matches.py:70-100,
visualize_matches_sequence.py:69-104, 297-320), each frame extracted once.

``run_sequence`` can fan frames / pairs out over a process pool so the CPU baseline uses every
host core (the reference itself is single-process Python with intra-op BLAS threads).
"""

import os
import time
import zlib
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from . import decode, gather, match, refiner

_W = {}


def extract_frame(sal, feat, weights, K, nms_radius=2, pct=0.5):
    """sal (H,W) fp32, feat (h,w,C) fp32 -> keypoints_pixel (K,2), scores (K,), descriptors (K,D)."""
    kp, sc, _ = decode.select_keypoints(sal[None], K, nms_radius, pct)
    g = gather.extract_at_keypoints(feat[None], gather.pixel_to_patch(kp))
    d = refiner.refiner_forward(weights, g)
    return kp[0], sc[0], d[0]


def match_pair(variant, d1, d2, s1, s2, **kw):
    if variant == 1:
        return match.match_m1(d1, d2, kw.get("ratio_thresh", 0.8))
    if variant == 2:
        return match.match_m2(d1, d2, s1, s2, **kw)
    if variant == 3:
        return match.match_m3(d1, d2, kw.get("ratio_threshold", 0.9))
    raise ValueError(variant)


def _init_worker(params, blas_threads):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(blas_threads)
    except Exception:
        pass
    _W["weights"] = refiner.RefinerWeights(params)


def _extract_job(args):
    sal, feat, K = args
    return extract_frame(sal, feat, _W["weights"], K)


def pair_record(ij, kp1, kp2):
    """(count, crc32) of a match list in a form that does not depend on the ORDER of equal-score
    keypoints (``torch.topk`` leaves it unspecified): every match (i, j) becomes
    the int32 row (x1, y1, x2, y2) of its two keypoints and the rows are sorted.  Lets bench.py
    compare whole lists between implementations, not just counts."""
    ij = np.asarray(ij, dtype=np.int64).reshape(-1, 2)
    rows = np.concatenate([np.asarray(kp1)[ij[:, 0]], np.asarray(kp2)[ij[:, 1]]], axis=1).astype(np.int32)
    rows = rows[np.lexsort(rows.T[::-1])] if rows.shape[0] else rows
    return int(rows.shape[0]), zlib.crc32(np.ascontiguousarray(rows).tobytes())


def _match_job(args):
    variant, d1, d2, s1, s2, kp1, kp2 = args
    r = match_pair(variant, d1, d2, s1, s2)
    ij = [(i, j) for i, j, _ in r] if isinstance(r, list) else r[0]
    return pair_record(ij, kp1, kp2)


def run_sequence(sal, feat, weights, K, variant=1, workers=1):
    """sal (T,H,W) fp32, feat (T,h,w,C) fp32 -> (per-pair (match count, crc32 of the pair list), seconds).
    workers > 1 distributes frames, then pairs, over that many processes (1 BLAS thread each)."""
    T = sal.shape[0]
    t0 = time.perf_counter()
    if workers <= 1:
        ex = [extract_frame(sal[t], feat[t], weights, K) for t in range(T)]
        counts = []
        for t in range(T - 1):
            counts.append(_match_job((variant, ex[t][2], ex[t + 1][2], ex[t][1], ex[t + 1][1], ex[t][0], ex[t + 1][0])))
    else:
        with ProcessPoolExecutor(workers, initializer=_init_worker,
                                 initargs=(weights.p, 1)) as pool:
            t0 = time.perf_counter()                      # pool start-up is not the algorithm
            ex = list(pool.map(_extract_job, [(sal[t], feat[t], K) for t in range(T)]))
            counts = list(pool.map(_match_job, [(variant, ex[t][2], ex[t + 1][2], ex[t][1], ex[t + 1][1],
                                                 ex[t][0], ex[t + 1][0]) for t in range(T - 1)]))
    return counts, time.perf_counter() - t0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
