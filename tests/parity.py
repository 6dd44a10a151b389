"""Shared parity comparators (tie-aware), used by CPU and GPU tests.

Tie rules (north_star): ``torch.topk`` leaves the order of equal scores
unspecified, so keypoint lists are compared as score-sequences (bit-exact) plus coordinate
multisets per run of equal scores, and members of the k-th boundary run may differ; match lists
must be identical except for decisions whose deciding margin is < 1e-6, which are counted.
"""

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NEAR_TIE = 1e-6
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, entry):
    """Persist a parity record (near-tie exception counts, index agreement, margins): north_star asks
    for the exceptions to be "excepted and counted".  Entries accumulate in one JSON file —
    ``$SSLAM_PARITY_RECORD`` or ``gpurun_out/parity_record.json`` (merged back from the GPU box);
    the copy judged is committed under ``profiles/``."""
    path = os.environ.get("SSLAM_PARITY_RECORD") or os.path.join(ROOT, "gpurun_out", "parity_record.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = entry
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass
    print(f"[parity] {name}: {json.dumps(entry)}")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"])) if "meta" in z.files else {}
    return z, meta


def compare_keypoints(sal, kp_ref, sc_ref, kp_got, sc_got, sorted_from=0):
    """Assert two (K,2)/(K,) decode results agree up to top-k tie freedom.

    ``sorted_from``: index where the score-sorted tail starts (0 for the main branch / branch C;
    n for branch B, whose first n rows are in row-major order and must match exactly).
    Returns the number of positions that differ only by a permitted tie.
    """
    kp_ref, kp_got = np.asarray(kp_ref), np.asarray(kp_got)
    sc_ref, sc_got = np.asarray(sc_ref), np.asarray(sc_got)
    assert kp_ref.shape == kp_got.shape and sc_ref.shape == sc_got.shape
    assert kp_got.dtype == np.float32 and sc_got.dtype == np.float32
    assert np.array_equal(sc_ref.view(np.uint32), sc_got.view(np.uint32)), "scores differ"
    if sorted_from:
        assert np.array_equal(kp_ref[:sorted_from], kp_got[:sorted_from]), "unsorted head differs"
    xs, ys = kp_got[:, 0].astype(np.int64), kp_got[:, 1].astype(np.int64)
    assert np.array_equal(sal[ys, xs].view(np.uint32), sc_got.view(np.uint32)), \
        "a keypoint's score is not the map value at its coordinate"
    diff = np.nonzero((kp_ref != kp_got).any(axis=1))[0]
    if diff.size == 0:
        return 0
    tail_ref, tail_got = kp_ref[sorted_from:], kp_got[sorted_from:]
    tail_sc = sc_ref[sorted_from:]
    boundary = tail_sc[-1]
    for s in np.unique(tail_sc[diff - sorted_from]):
        run = tail_sc == s
        assert run.sum() >= 2 or s == boundary, f"non-tie mismatch at score {s}"
        if s != boundary:
            a = sorted(map(tuple, tail_ref[run].tolist()))
            b = sorted(map(tuple, tail_got[run].tolist()))
            assert a == b, f"tie run at score {s} holds different points"
    return int(diff.size)


def _margins(S):
    S = np.asarray(S, dtype=np.float64)
    part = -np.partition(-S, 1, axis=1) if S.shape[1] > 1 else None
    row_margin = (part[:, 0] - part[:, 1]) if part is not None else np.full(S.shape[0], np.inf)
    partc = -np.partition(-S.T, 1, axis=1) if S.shape[0] > 1 else None
    col_margin = (partc[:, 0] - partc[:, 1]) if partc is not None else np.full(S.shape[1], np.inf)
    return row_margin, col_margin


def compare_matches(S, ref_pairs, got_pairs, threshold_margin=None):
    """Assert two (K',2) match lists agree except for near-tie decisions (< 1e-6).

    ``threshold_margin(i, j)`` optionally returns |score - threshold| for a pair so that
    accept/reject flips at a threshold are excused the same way.  Returns the exception count.
    """
    ref = {tuple(map(int, r)) for r in np.asarray(ref_pairs).reshape(-1, 2)}
    got = {tuple(map(int, r)) for r in np.asarray(got_pairs).reshape(-1, 2)}
    if ref == got:
        return 0
    row_margin, col_margin = _margins(S)
    excused = 0
    for (i, j) in ref ^ got:
        near = row_margin[i] < NEAR_TIE or col_margin[j] < NEAR_TIE
        if not near:
            # the row's other candidate column may be the near-tied one
            jj = int(np.argmax(S[i]))
            near = col_margin[jj] < NEAR_TIE
        if not near and threshold_margin is not None:
            near = threshold_margin(i, j) < NEAR_TIE
        assert near, f"pair ({i},{j}) differs and is not a near tie " \
                     f"(row margin {row_margin[i]:.3g}, col margin {col_margin[j]:.3g})"
        excused += 1
    return excused
