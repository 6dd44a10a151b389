"""Evaluation adaptors: oracle vs the fixtures written by the reference's own
``compute_ground_truth_matches`` / ``evaluate_matches`` / ``compute_repeatability`` (CPU), and the CUDA
kernels vs the same fixtures (``-m gpu``).  Index lists bit-exact; float statistics to 1e-12 relative
(double path) / 1e-6 (the float32 no-homography path, whose mean NumPy sums pairwise)."""

import json

import numpy as np
import pytest
import torch

import oracle
from oracle import evaluation as oev
from parity import load_golden

EV, _ = load_golden("evaluation")
EV_META = json.loads(str(EV["meta"]))


@pytest.mark.parametrize("name", list(EV_META))
def test_oracle_evaluation_golden(name):
    k1, k2, H = EV[name + ".k1"], EV[name + ".k2"], EV[name + ".H"]
    for thr in (3.0, 1.0):
        assert np.array_equal(oev.compute_ground_truth_matches(k1, k2, H, thr), EV[f"{name}.gt{thr:g}"])
    ev = oev.evaluate_matches(EV[name + ".pred"], EV[name + ".gt3"], len(k1), len(k2))
    assert [ev["tp"], ev["fp"], ev["fn"]] == EV[name + ".eval"].tolist()
    assert np.allclose([ev["precision"], ev["recall"], ev["f1"], ev["inlier_ratio"]], EV[name + ".evalf"], rtol=1e-15)
    for tag, HH in (("repH", H), ("rep0", None)):
        r = oev.compute_repeatability(k1, k2, HH, 3.0)
        got = [r["repeatability"], float(r["repeatable_count"]), float(r["mean_nn_distance"]),
               float(r["median_nn_distance"])]
        assert np.array_equal(got, EV[f"{name}.{tag}"])


def test_match_record_roundtrip_cpu():
    """Wire format: pack -> unpack is the identity (host tensors; no kernel involved)."""
    from sslam_b200 import evaluation as ev
    g = torch.Generator().manual_seed(1)
    P, N = 5, 17
    counts = torch.randint(0, N + 1, (P,), generator=g, dtype=torch.int32)
    pairs = torch.randint(0, 100, (P, N, 2), generator=g, dtype=torch.int32)
    scores = torch.rand(P, N, generator=g)
    for p in range(P):
        pairs[p, int(counts[p]):] = -1
        scores[p, int(counts[p]):] = 0
    rec = ev.pack_match_records(pairs, scores, counts)
    assert rec.shape == (P, 3 * N + 1) and rec.dtype == torch.int32
    p2, s2, c2 = ev.unpack_match_records(rec)
    assert torch.equal(p2, pairs) and torch.equal(s2, scores) and torch.equal(c2, counts)


def test_match_list_writer_roundtrip(tmp_path):
    from sslam_b200 import evaluation as ev
    P, N = 3, 8
    pairs = torch.full((P, N, 2), -1, dtype=torch.int32)
    scores = torch.zeros(P, N)
    counts = torch.tensor([2, 0, 3], dtype=torch.int32)
    pairs[0, :2] = torch.tensor([[0, 5], [3, 1]])
    pairs[2, :3] = torch.tensor([[1, 1], [2, 7], [6, 0]])
    scores[0, :2] = torch.tensor([0.9, 0.8])
    scores[2, :3] = torch.tensor([0.7, 0.95, 0.85])
    path = tmp_path / "lists.npz"
    ev.write_match_lists(str(path), pairs, scores, counts, meta={"matcher": "M2"})
    lists, pidx, meta = ev.read_match_lists(str(path))
    assert meta == {"matcher": "M2"} and pidx.tolist() == [[0, 1], [1, 2], [2, 3]]
    assert lists[0][0].dtype == np.int64 and lists[0][1].dtype == np.float32
    assert lists[0][0].tolist() == [[0, 5], [3, 1]] and lists[1][0].shape == (0, 2) and lists[1][1].shape == (0,)
    assert np.allclose(lists[2][1], [0.7, 0.95, 0.85])
    out = ev.write_results_json(str(tmp_path / "r.json"), [
        {"sequence": "fr1_desk", "mean_precision": 0.5, "mean_recall": 0.25, "mean_inlier_ratio": 0.44}])
    assert out["sequences"][0] == {"name": "fr1_desk", "precision": 0.5, "recall": 0.25, "inlier_ratio": 0.44}
    assert out["overall_inlier_ratio"] == 0.44


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(EV_META))
def test_gpu_evaluation_golden(name):
    from sslam_b200 import evaluation as ev
    k1, k2, H = EV[name + ".k1"], EV[name + ".k2"], EV[name + ".H"]
    for thr in (3.0, 1.0):
        gt = ev.compute_ground_truth_matches(k1, k2, H, thr)
        assert gt.dtype == np.int64 and np.array_equal(gt, EV[f"{name}.gt{thr:g}"])
    res = ev.evaluate_matches(EV[name + ".pred"], EV[name + ".gt3"], len(k1), len(k2))
    assert [res["tp"], res["fp"], res["fn"]] == EV[name + ".eval"].tolist()
    assert np.allclose([res["precision"], res["recall"], res["f1"], res["inlier_ratio"]], EV[name + ".evalf"],
                       rtol=1e-15)
    for tag, HH, tol in (("repH", H, 1e-12), ("rep0", None, 1e-6)):
        r = ev.compute_repeatability(k1, k2, HH, 3.0)
        ref = EV[f"{name}.{tag}"]
        assert r["repeatability"] == ref[0] and float(r["repeatable_count"]) == ref[1]
        assert abs(float(r["mean_nn_distance"]) - ref[2]) <= tol * max(ref[2], 1.0)
        assert abs(float(r["median_nn_distance"]) - ref[3]) <= tol * max(ref[3], 1.0)


@pytest.mark.gpu
def test_gpu_evaluation_batched_device_api():
    """Banks + pair_index on device: every pair equals the single-pair host API; distances equal the
    oracle's bit for bit (double path)."""
    from sslam_b200 import evaluation as ev, ops
    dev = torch.device("cuda", 0)
    names = ["small_shift", "ident"]
    N = 64
    k1 = np.stack([EV[n + ".k1"][:N] for n in names])
    k2 = np.stack([EV[n + ".k2"][:N] for n in names])
    Hs = np.stack([EV[n + ".H"] for n in names])
    idx = torch.tensor([[0, 0], [1, 1], [0, 1]], dtype=torch.int32, device=dev)
    Hp = torch.as_tensor(np.stack([Hs[0], Hs[1], Hs[0]])).to(dev)
    t1, t2 = torch.as_tensor(k1).to(dev), torch.as_tensor(k2).to(dev)
    md, am = ops.nn_points(t1, t2, H=Hp, pair_index=idx)
    pairs, counts = ev.ground_truth_matches_device(t1, t2, Hp, 3.0, pair_index=idx)
    for p, (a, b) in enumerate(idx.cpu().tolist()):
        omd, oam = oev.nearest(oev.warp_points(k1[a], Hp[p].cpu().numpy()), k2[b])
        assert np.array_equal(am[p].cpu().numpy(), oam)
        assert np.allclose(md[p].cpu().numpy(), omd, rtol=1e-14, atol=1e-12)
        ref = oev.compute_ground_truth_matches(k1[a], k2[b], Hp[p].cpu().numpy(), 3.0)
        assert np.array_equal(pairs[p, :int(counts[p])].cpu().numpy(), ref)
        assert bool((pairs[p, int(counts[p]):] == -1).all())
    tpfpfn = ev.evaluate_matches_device(pairs, counts, pairs, counts, N).cpu().numpy()
    assert np.array_equal(tpfpfn[:, 0], counts.cpu().numpy()) and (tpfpfn[:, 1:] == 0).all()
    cnt, dist = ev.repeatability_device(t1, t2, None, 3.0, pair_index=idx)
    for p, (a, b) in enumerate(idx.cpu().tolist()):
        r = oev.compute_repeatability(k1[a], k2[b], None, 3.0)
        assert int(cnt[p]) == int(r["repeatable_count"])
        assert dist.dtype == torch.float32
