"""Evaluation adaptors and the match-list wire / on-disk format.

Host API — NumPy in, NumPy out, with the names, arguments and return values of the reference's
evaluation scripts (the N x M distance matrices run as kernels behind the C ABI):

  compute_ground_truth_matches   test/test_descriptor_quality.py:144-183 there
  evaluate_matches               test/test_descriptor_quality.py:185-231
  compute_repeatability          test/test_repeatability.py:79-128

Device API — ``*_device`` functions take banks of keypoint sets and padded pair lists and keep the
results in HBM (what the batched pipeline uses).

Wire format — ``pack_match_records`` / ``unpack_match_records`` define the fixed-size record the
final NCCL gather moves (one int32 row of ``3*N + 1`` words per pair: count, N (i, j) pairs with -1
padding, N fp32 scores bit-cast), and ``write_match_lists`` / ``read_match_lists`` the on-disk form
the evaluation scripts consume: per pair a ``(K', 2) int64`` array and a ``(K',) float32`` array in
one ``.npz``, plus a JSON summary shaped like test_descriptor_quality.py:472-489.
"""

import json

import numpy as np
import torch

from . import ops


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("sslam_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _kp(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(dev)[None].contiguous()


# ------------------------------------------------------------------------------------ host API
def compute_ground_truth_matches(kpts1, kpts2, H, threshold=3.0):
    """(N,2), (M,2) keypoints + (3,3) homography frame1->frame2 -> gt_matches (K,2) int64: every i
    whose warped point has a keypoint of frame 2 within ``threshold`` pixels, paired with the nearest."""
    dev = _device()
    Hd = torch.as_tensor(np.asarray(H, dtype=np.float64)).to(dev).reshape(1, 3, 3)
    md, am = ops.nn_points(_kp(kpts1, dev), _kp(kpts2, dev), H=Hd)
    pairs, counts = ops.gt_matches(md, am, threshold)
    n = int(counts[0])
    return pairs[0, :n].cpu().numpy().astype(np.int64)


def evaluate_matches(pred_matches, gt_matches, num_kpts1, num_kpts2):
    """Precision / recall / F1 / inlier ratio of predicted against ground-truth (K,2) index lists."""
    dev = _device()
    pred = np.asarray(pred_matches).reshape(-1, 2)
    gt = np.asarray(gt_matches).reshape(-1, 2)
    for name, a in (("pred_matches", pred), ("gt_matches", gt)):
        if np.unique(a[:, 0]).size != a.shape[0]:
            raise ValueError(f"{name}: a first index occurs twice (the kernel scores one-to-one lists)")

    def pad(a):
        t = torch.full((1, max(a.shape[0], 1), 2), -1, dtype=torch.int32)
        t[0, :a.shape[0]] = torch.as_tensor(a.astype(np.int32))
        return t.to(dev), torch.tensor([a.shape[0]], dtype=torch.int32, device=dev)

    pp, pc = pad(pred)
    gp, gc = pad(gt)
    tp, fp, fn = (int(v) for v in ops.eval_matches(pp, pc, gp, gc, num_kpts1)[0].cpu())
    return _scores(tp, fp, fn, pred.shape[0], gt.shape[0])


def _scores(tp, fp, fn, n_pred, n_gt):
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    f1 = 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0.0
    return {"tp": tp, "fp": fp, "fn": fn, "precision": precision, "recall": recall, "f1": f1,
            "inlier_ratio": tp / n_pred if n_pred > 0 else 0.0,
            "num_pred_matches": n_pred, "num_gt_matches": n_gt}


def compute_repeatability(kpts1, kpts2, H=None, threshold=3.0):
    """Fraction of frame-1 keypoints with a frame-2 keypoint within ``threshold`` pixels after the
    optional homography; also the count, the mean and the median nearest distance."""
    dev = _device()
    Hd = None if H is None else torch.as_tensor(np.asarray(H, dtype=np.float64)).to(dev).reshape(1, 3, 3)
    md, am = ops.nn_points(_kp(kpts1, dev), _kp(kpts2, dev), H=Hd)
    _, counts = ops.gt_matches(md, am, threshold, want_pairs=False)
    d = md[0].cpu().numpy()                       # float64 with H, float32 without — as the reference
    repeatable = np.int64(int(counts[0]))
    return {"repeatability": repeatable / len(d), "repeatable_count": repeatable,
            "total_keypoints": len(d), "mean_nn_distance": d.mean(),
            "median_nn_distance": np.median(d)}


# ------------------------------------------------------------------------------------ device API
def ground_truth_matches_device(kpts1, kpts2, H, threshold=3.0, pair_index=None, num_pairs=None):
    """Banks (F1,N,2) / (F2,M,2) on device, H (P,3,3) -> padded gt pairs (P,N,2) int32, counts (P,)."""
    md, am = ops.nn_points(kpts1, kpts2, H=H, pair_index=pair_index, num_pairs=num_pairs)
    return ops.gt_matches(md, am, threshold)


def evaluate_matches_device(pred_pairs, pred_counts, gt_pairs, gt_counts, num_kpts1):
    """Padded lists on device -> (P,3) int32 tp / fp / fn on device."""
    return ops.eval_matches(pred_pairs, pred_counts, gt_pairs, gt_counts, num_kpts1)


def repeatability_device(kpts1, kpts2, H=None, threshold=3.0, pair_index=None, num_pairs=None):
    """Banks on device -> (repeatable counts (P,) int32, min distances (P,N)) on device."""
    md, am = ops.nn_points(kpts1, kpts2, H=H, pair_index=pair_index, num_pairs=num_pairs)
    _, counts = ops.gt_matches(md, am, threshold, want_pairs=False)
    return counts, md


# ------------------------------------------------------------------------------------ wire format
def pack_match_records(pairs, pair_scores, counts):
    """(P,N,2) int32, (P,N) fp32, (P,) int32 -> one contiguous (P, 3N+1) int32 record block:
    [count | i0 j0 i1 j1 ... (-1 padded) | score bits].  One tensor = one collective."""
    P, N = pair_scores.shape
    rec = torch.empty(P, 3 * N + 1, dtype=torch.int32, device=pairs.device)
    rec[:, 0] = counts
    rec[:, 1:1 + 2 * N] = pairs.reshape(P, 2 * N)
    rec[:, 1 + 2 * N:] = pair_scores.contiguous().view(torch.int32)
    return rec


def unpack_match_records(rec):
    """Inverse of pack_match_records (views into ``rec`` where possible)."""
    P = rec.shape[0]
    N = (rec.shape[1] - 1) // 3
    counts = rec[:, 0]
    pairs = rec[:, 1:1 + 2 * N].reshape(P, N, 2)
    scores = rec[:, 1 + 2 * N:].contiguous().view(torch.float32)
    return pairs, scores, counts


def to_match_lists(pairs, pair_scores, counts):
    """Padded device/host lists -> list of ((K',2) int64 ndarray, (K',) float32 ndarray), the form the
    reference's scripts work with (visualize_matches_sequence.py:154-155 for the empty case)."""
    pairs, pair_scores, counts = pairs.cpu().numpy(), pair_scores.cpu().numpy(), counts.cpu().numpy()
    return [(pairs[p, :int(counts[p])].astype(np.int64), pair_scores[p, :int(counts[p])].astype(np.float32))
            for p in range(pairs.shape[0])]


def write_match_lists(path, pairs, pair_scores, counts, pair_index=None, meta=None):
    """Write the match lists of P pairs to ``path`` (.npz): ``matches_{p}`` (K',2) int64,
    ``scores_{p}`` (K',) float32, ``pair_index`` (P,2) int32 frame ids, ``meta`` JSON string."""
    lists = to_match_lists(pairs, pair_scores, counts)
    arrays = {}
    for p, (m, s) in enumerate(lists):
        arrays[f"matches_{p}"] = m
        arrays[f"scores_{p}"] = s
    P = len(lists)
    if pair_index is None:
        pair_index = np.stack([np.arange(P), np.arange(P) + 1], 1)
    arrays["pair_index"] = np.asarray(pair_index.cpu() if hasattr(pair_index, "cpu") else pair_index, dtype=np.int32)
    arrays["meta"] = np.array(json.dumps(meta or {}))
    np.savez_compressed(path, **arrays)
    return P


def read_match_lists(path):
    z = np.load(path, allow_pickle=False)
    P = z["pair_index"].shape[0]
    return ([(z[f"matches_{p}"], z[f"scores_{p}"]) for p in range(P)], z["pair_index"],
            json.loads(str(z["meta"])))


def write_results_json(path, sequences):
    """JSON summary in the layout of test_descriptor_quality.py:472-489: ``sequences`` is a list of
    dicts with name / mean_precision / mean_recall / mean_inlier_ratio."""
    out = {"overall_inlier_ratio": float(np.mean([s["mean_inlier_ratio"] for s in sequences])) if sequences else 0.0,
           "overall_precision": float(np.mean([s["mean_precision"] for s in sequences])) if sequences else 0.0,
           "overall_recall": float(np.mean([s["mean_recall"] for s in sequences])) if sequences else 0.0,
           "sequences": [{"name": s["sequence"], "precision": s["mean_precision"], "recall": s["mean_recall"],
                          "inlier_ratio": s["mean_inlier_ratio"]} for s in sequences]}
    with open(path, "w") as f:
        json.dump(out, f, indent=2)
    return out
