"""GPU tests of the reference-named surface, the frame store / sequence drivers
(§8(f) N3) and the optional cell-softmax decode mode — against the oracle on identical inputs.

The selector's saliency is computed once on the device and handed to the oracle (SURVEY.md §8(d):
"computed once ... and shared by both implementations"); from there keypoints must be bit-exact,
descriptors within 1e-5 abs and match lists identical up to counted near ties."""

import sys
import os

import numpy as np
import pytest
import torch

import oracle
from parity import compare_matches, record

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-slam-master_b200", "test"))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def native_modules(dev, D=128, K=500):
    from models.dino_backbone import DinoBackbone
    from models.keypoint_selector import KeypointSelector
    from models.descriptor_refiner import DescriptorRefiner
    torch.manual_seed(0)
    sel = KeypointSelector(384, 256)
    ref = DescriptorRefiner(384, 384, D, 4)
    cfg = {"model": {"input_size": 448, "num_keypoints": K, "selector_hidden": 256, "descriptor_dim": D,
                     "refiner_hidden": 384, "backbone": "none"}}
    return dict(backbone=DinoBackbone(load_vit=False), selector=sel, refiner=ref, config=cfg)


def oracle_frame(sal_hw, feat_hwc, weights, K):
    """Reference-native grid: keypoints in patch units on the map's own grid."""
    kp, sc, _ = oracle.select_keypoints(sal_hw[None], K)
    d = oracle.refiner_forward(weights, oracle.extract_at_keypoints(feat_hwc[None], kp))
    return oracle.patch_to_pixel(kp)[0], sc[0], d[0]


def test_match_visualizer_native_grid(dev):
    """MatchVisualizer.features_from_patch_map + find_matches (default = tcgen05 f16x3 matcher) on the
    reference-native 28x28 grid (c0): always branch B with duplicate keypoints."""
    from sslam_b200 import ops, synth
    from visualize_matches import MatchVisualizer
    mods = native_modules(dev)
    mv = MatchVisualizer(**mods)
    F = synth.native_grid_case(batch=3).to(dev)
    w = oracle.RefinerWeights.from_state_dict(mods["refiner"].state_dict())
    feats, refs = [], []
    for b in range(3):
        f = mv.features_from_patch_map(F[b:b + 1])
        assert set(f) == {"keypoints_pixel", "scores", "descriptors"}
        assert f["keypoints_pixel"].shape == (500, 2) and f["descriptors"].shape == (500, 128)
        assert all(isinstance(v, np.ndarray) and v.dtype == np.float32 for v in f.values())
        with torch.no_grad():
            sal = mv.selector(F[b:b + 1])[0, :, :, 0].cpu().numpy()
        okp, osc, od = oracle_frame(sal, F[b].cpu().numpy(), w, 500)
        assert np.array_equal(f["keypoints_pixel"], okp) and np.array_equal(f["scores"], osc)
        assert np.abs(f["descriptors"] - od).max() < 1e-5
        assert len({tuple(r) for r in okp.tolist()}) < 500            # duplicates: the fallback branch fired
        feats.append(f); refs.append(od)
    ops.profile_enable(True)
    got = mv.find_matches(feats[0]["descriptors"], feats[1]["descriptors"], ratio_thresh=0.8)
    torch.cuda.synchronize()
    kinds = ops.profile_read()
    ops.profile_enable(False)
    assert "match_tc" in kinds and "match_f32" not in kinds, kinds       # the tcgen05 kernel is the default
    assert all(isinstance(t, tuple) and len(t) == 3 for t in got)
    d0, d1 = feats[0]["descriptors"], feats[1]["descriptors"]
    S = d0.astype(np.float64) @ d1.astype(np.float64).T
    ref = oracle.match_m1(d0, d1, 0.8)
    exc = compare_matches(S, np.array([(i, j) for i, j, _ in ref]).reshape(-1, 2),
                          np.array([(i, j) for i, j, _ in got]).reshape(-1, 2))
    record("surface.match_visualizer_c0", {"matches": len(got), "near_tie_exceptions": int(exc)})


def test_sequence_matcher_extract_and_process_spacing(dev, tmp_path):
    """SequenceMatcher.extract_from_patch_map (the part of ``extract`` after the backbone),
    match_with_quality and the module-level process_spacing with the reference's signature."""
    import visualize_matches_sequence as vms
    from sslam_b200 import evaluation, synth
    mods = native_modules(dev)
    sm = vms.SequenceMatcher(**mods)
    F = synth.native_grid_case(batch=5, seed=9).to(dev)
    rng = np.random.Generator(np.random.PCG64(4))
    grays = [(rng.integers(0, 256, size=(448, 448)) / 255.0).astype(np.float32) for _ in range(5)]
    w = oracle.RefinerWeights.from_state_dict(mods["refiner"].state_dict())
    f0 = sm.extract_from_patch_map(F[0:1], gray=grays[0])
    assert set(f0) == {"image", "saliency", "keypoints_pixel", "scores", "intensity", "descriptors"}
    assert f0["saliency"].shape == (28, 28) and f0["intensity"].shape == (500,)
    okp, osc, od = oracle_frame(f0["saliency"], F[0].cpu().numpy(), w, 500)
    assert np.array_equal(f0["keypoints_pixel"], okp) and np.array_equal(f0["scores"], osc)
    xs = np.clip(okp[:, 0].round().astype(int), 0, 447); ys = np.clip(okp[:, 1].round().astype(int), 0, 447)
    assert np.array_equal(f0["intensity"], grays[0][ys, xs])

    images = [(F[b:b + 1], grays[b]) for b in range(5)]
    kw = dict(saliency_weight=0.3, min_saliency=0.3, min_descriptor_sim=0.5, min_intensity=0.15)
    for spacing in (1, 2):
        scores = vms.process_spacing(sm, images, spacing, tmp_path, max_pairs=3, max_matches=50, gap=20, **kw)
        assert isinstance(scores, list)
        ref_scores, exc = [], 0
        npairs = 0
        for i in range(0, 5 - spacing, spacing):
            if npairs >= 3:
                break
            a = sm.extract_from_patch_map(F[i:i + 1], gray=grays[i])
            b = sm.extract_from_patch_map(F[i + spacing:i + spacing + 1], gray=grays[i + spacing])
            rm, rq = oracle.match_m2(a["descriptors"], b["descriptors"], a["scores"], b["scores"], 0.3, 0.3, 0.5,
                                     a["intensity"], b["intensity"], 0.15)
            gm, gq = sm.match_with_quality(a["descriptors"], b["descriptors"], a["scores"], b["scores"],
                                           intensity1=a["intensity"], intensity2=b["intensity"], **kw)
            assert gm.dtype == np.int64 and gq.dtype == np.float32
            S = a["descriptors"].astype(np.float64) @ b["descriptors"].astype(np.float64).T
            exc += compare_matches(S, rm, gm, threshold_margin=lambda i_, j_: abs(S[i_, j_] - 0.5))
            ref_scores.extend(gq.tolist())
            npairs += 1
            lists, pidx, meta = evaluation.read_match_lists(
                str(tmp_path / f"spacing_{spacing}" / f"matches_frame{i:06d}_to_frame{i + spacing:06d}.npz"))
            assert pidx.tolist() == [[i, i + spacing]] and meta["total_matches"] == len(gm)
            assert lists[0][0].shape[0] == min(50, len(gm))
        assert scores == ref_scores
        record(f"surface.process_spacing.{spacing}", {"pairs": npairs, "near_tie_exceptions": int(exc)})


def test_process_spacings_device_and_frame_store(dev):
    """Device-resident multi-spacing driver and the ring-buffer frame store against the oracle."""
    import visualize_matches_sequence as vms
    from sslam_b200 import matchers, synth
    from sslam_b200.framestore import FrameStore
    from sslam_b200.pipeline import FrontEnd
    mods = native_modules(dev, D=128, K=96)
    sm = vms.SequenceMatcher(**mods)
    T, K = 12, 96
    sal, feat = synth.make_sequence(T, seq_id=7, height=96, width=128)
    w = oracle.RefinerWeights.from_state_dict(mods["refiner"].state_dict())
    okp, osc, _ = oracle.select_keypoints(sal.numpy(), K)
    od = oracle.refiner_forward(w, oracle.extract_at_keypoints(feat.numpy(), oracle.pixel_to_patch(okp)))
    res = sm.process_spacings_device(sal.to(dev), feat.to(dev), spacings=(1, 5, 10, 20), chunk=5,
                                     min_saliency=0.2, min_descriptor_sim=0.7)
    assert set(res) == {1, 5, 10}                                  # spacing 20 has no pair in 12 frames
    exc = 0
    for sp, (pidx, pairs, q, counts) in res.items():
        assert pidx.tolist() == [[i, i + sp] for i in range(0, T - sp, sp)]
        for p, (a, b) in enumerate(pidx.tolist()):
            S = od[a].astype(np.float64) @ od[b].astype(np.float64).T
            rm, rq = oracle.match_m2(od[a], od[b], osc[a], osc[b])
            gm = pairs[p, :int(counts[p])].cpu().numpy()
            exc += compare_matches(S, rm, gm, threshold_margin=lambda i, j: abs(S[i, j] - 0.7))
    # ring buffer: capacity 5, three pushes that wrap; resident frames equal a direct extraction
    fe = FrontEnd(mods["refiner"].to(dev), num_keypoints=K, grid="pixel")
    direct = fe.extract(sal.to(dev), feat.to(dev))
    store = FrameStore(fe, capacity=5)
    seen = []
    for s in range(0, 9, 3):
        fr = store.push(sal[s:s + 3].to(dev), feat[s:s + 3].to(dev))
        seen.append(list(fr))
        fp, pairs, q, counts = store.match_spacings(fr, spacings=(1, 2, 4), variant=matchers.M2)
        assert fp == [(t - s_, t) for s_ in (1, 2, 4) for t in fr if t - s_ >= store.oldest()]
        for p, (a, b) in enumerate(fp):
            S = od[a].astype(np.float64) @ od[b].astype(np.float64).T
            rm, _ = oracle.match_m2(od[a], od[b], osc[a], osc[b])
            exc += compare_matches(S, rm, pairs[p, :int(counts[p])].cpu().numpy(),
                                   threshold_margin=lambda i, j: abs(S[i, j] - 0.7))
    assert seen == [[0, 1, 2], [3, 4, 5], [6, 7, 8]] and store.oldest() == 4
    for t in range(4, 9):
        f = store.frame(t)
        assert torch.equal(f["descriptors"], direct["descriptors"][t])
        assert torch.equal(f["keypoints_pixel"], direct["keypoints_pixel"][t])
    with pytest.raises(KeyError):
        store.frame(3)
    with pytest.raises(KeyError):
        store.match_pairs([(2, 8)])
    record("surface.frame_store_multi_spacing", {"near_tie_exceptions": int(exc)})


def test_tracking_counters_on_device(dev):
    """TrackingTester.track_frame_sequence statistics (test/test_tracking.py:139-199 there) from the
    device-side counters equal the oracle's M5 counts frame by frame."""
    from test_tracking import TrackingTester
    from sslam_b200 import synth
    mods = native_modules(dev, D=128, K=96)
    tt = TrackingTester(**mods)
    T, K = 14, 96
    sal, feat = synth.make_sequence(T, seq_id=8, height=96, width=128)
    w = oracle.RefinerWeights.from_state_dict(mods["refiner"].state_dict())
    okp, osc, _ = oracle.select_keypoints(sal.numpy(), K)
    od = oracle.refiner_forward(w, oracle.extract_at_keypoints(feat.numpy(), oracle.pixel_to_patch(okp)))
    for spacing, max_frames, min_matches in ((1, 100, 40), (3, 100, 40), (2, 4, 60)):
        res = tt.track_frame_sequence(sal.to(dev), feat.to(dev), max_frames=max_frames, min_matches=min_matches,
                                      match_threshold=0.8, frame_spacing=spacing, sequence="synthetic")
        ids = list(range(0, min(max_frames * spacing, T), spacing))
        ref = [oracle.match_m5(od[a], od[b], 0.8) for a, b in zip(ids[:-1], ids[1:])]
        got = [int(c) for c in res["match_counts"]]
        # a row whose best similarity is within 2e-6 of the threshold may fall either side
        for (a, b), r, g in zip(zip(ids[:-1], ids[1:]), ref, got):
            if r != g:
                best = (od[a].astype(np.float64) @ od[b].astype(np.float64).T).max(1)
                assert abs(r - g) <= int((np.abs(best - 0.8) < 2e-6).sum())
        assert res["total_frames"] == len(ref) and res["frame_spacing"] == spacing
        assert res["tracked_frames"] == sum(c >= min_matches for c in got)
        assert res["lost_frames"] == len(ref) - res["tracked_frames"]
        assert abs(res["mean_match_ratio"] - np.mean(np.array(got) / K)) < 1e-12
    assert tt.count_matches(od[0], od[1], 0.8) == oracle.match_m5(od[0], od[1], 0.8)


def test_tester_entry_points(dev):
    """DescriptorQualityTester.extract_features- and PerformanceTester.forward_pass-shaped entry points."""
    from test_descriptor_quality import DescriptorQualityTester
    from test_performance import PerformanceTester
    from test_repeatability import RepeatabilityTester
    from sslam_b200 import synth
    mods = native_modules(dev)
    F = synth.native_grid_case(batch=2, seed=21).to(dev)
    dq = DescriptorQualityTester(**mods)
    kp, desc, sc = dq.extract_features(F[0:1])
    assert kp.shape == (500, 2) and desc.shape == (500, 128) and sc.shape == (500,)
    assert all(a.dtype == np.float32 for a in (kp, desc, sc))
    w = oracle.RefinerWeights.from_state_dict(mods["refiner"].state_dict())
    with torch.no_grad():
        sal = dq.selector(F[0:1])[0, :, :, 0].cpu().numpy()
    okp, osc, od = oracle_frame(sal, F[0].cpu().numpy(), w, 500)
    assert np.array_equal(kp, okp) and np.array_equal(sc, osc) and np.abs(desc - od).max() < 1e-5
    kp2, desc2, _ = dq.extract_features(F[1:2])
    m, dist = dq.find_mutual_nearest_neighbors(desc, desc2, 0.9)
    rm, rd = oracle.match_m3(desc, desc2, 0.9)
    S = desc.astype(np.float64) @ desc2.astype(np.float64).T
    second = lambda i: np.partition(S[i], -2)[-2]  # noqa: E731
    compare_matches(S, rm, m, threshold_margin=lambda i, j: abs(second(i) / (S[i, j] + 1e-8) - 0.9))
    pt = PerformanceTester(**mods)
    kpp, d, s = pt.forward_pass(F)
    assert kpp.shape == (2, 500, 2) and d.shape == (2, 500, 128) and s.shape == (2, 500) and d.is_cuda
    assert np.array_equal(dq.backbone.patch_to_pixel(kpp)[0].cpu().numpy(), kp)
    times = pt.measure_component_times(F[0:1], num_runs=3)
    assert set(times) == {"backbone", "selector", "selector_nms", "refiner", "total"}
    assert set(times["total"]) == {"mean", "std", "min", "max", "median"}
    rt = RepeatabilityTester(**mods)
    rk, rs = rt.detect_keypoints(F[0:1])
    assert np.array_equal(rk, kp) and np.array_equal(rs, sc)
    rep = rt.compute_repeatability(rk, kp2, None, 3.0)
    orep = oracle.evaluation.compute_repeatability(rk, kp2, None, 3.0)
    assert rep["repeatable_count"] == orep["repeatable_count"] and rep["repeatability"] == orep["repeatability"]


def test_sampler_autograd_path(dev):
    """extract_at_keypoints under autograd (the reference trainer samples features that require grad,
    train.py:324 there) uses the torch ops and has a gradient; the kernel path equals it to 1e-6."""
    from models.dino_backbone import DinoBackbone
    bb = DinoBackbone(load_vit=False).to(dev)
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(2, 6, 7, 16, generator=g).to(dev).requires_grad_(True)
    kp = (torch.rand(2, 9, 2, generator=g) * torch.tensor([6.0, 5.0])).to(dev)
    out = bb.extract_at_keypoints(feat, kp)
    out.sum().backward()
    assert feat.grad is not None and float(feat.grad.abs().sum()) > 0
    with torch.no_grad():
        k = bb.extract_at_keypoints(feat.detach(), kp)
    assert torch.allclose(out.detach(), k, rtol=1e-6, atol=1e-6)


def test_cell_softmax_decode_mode(dev):
    """OPTIONAL softmax / depth-to-space / border-mask decode (north_star vocabulary; no reference
    code exists for it): heatmap against a plain PyTorch restatement written here, then the ordinary
    decode on that heatmap against the oracle."""
    from models.keypoint_selector import KeypointSelector
    from sslam_b200 import ops
    sel = KeypointSelector(8, 8)
    g = torch.Generator().manual_seed(17)
    for (B, Hc, Wc, cell, border) in ((2, 30, 40, 8, 4), (1, 7, 5, 8, 0), (3, 12, 9, 4, 3), (1, 6, 6, 2, 1)):
        logits = (torch.randn(B, cell * cell + 1, Hc, Wc, generator=g) * 3).to(dev)
        heat = ops.heatmap_from_cells(logits, cell=cell, border=border)
        # restatement: softmax over channels, drop the dustbin, depth-to-space, zero the border
        p = torch.softmax(logits.double(), dim=1)[:, :-1]
        ref = p.reshape(B, cell, cell, Hc, Wc).permute(0, 3, 1, 4, 2).reshape(B, Hc * cell, Wc * cell)
        if border:
            mask = torch.zeros_like(ref)
            mask[:, border:Hc * cell - border, border:Wc * cell - border] = 1
            ref = ref * mask
        assert heat.shape == ref.shape and heat.dtype == torch.float32
        assert torch.allclose(heat.double(), ref, rtol=2e-6, atol=1e-9)
        if border:
            assert float(heat[:, :border].abs().max()) == 0 and float(heat[:, :, -border:].abs().max()) == 0
        K = 40
        kp, sc, h2 = sel.select_keypoints_from_cells(logits, num_keypoints=K, cell=cell, border=border)
        assert torch.equal(h2, heat)
        okp, osc, _ = oracle.select_keypoints(heat.cpu().numpy(), K)
        assert np.array_equal(kp.cpu().numpy(), okp) and np.array_equal(sc.cpu().numpy(), osc)


def test_selector_head_kernel(dev):
    """the selector head (3x3 conv -> ReLU -> 1x1 conv -> sigmoid) as one tcgen05
    implicit-GEMM kernel.  Logits within 1e-5 x max(1, max|logit|) of the fp32 PyTorch head (cuDNN with
    TF32 off, and the CPU head) and of an fp64 evaluation, on the reference-native 28x28 grid, the 30x40
    and 60x80 grids and ragged shapes.  The weights here are inflated (1x1 weights x4, non-zero biases) so
    that logits reach |4|; the tolerance is stated relative to that scale (K = 9*384 = 3456 products per
    hidden unit; the error of the fp32 PyTorch heads against fp64 is recorded beside ours).  The
    reference-initialised head is held to an absolute 1e-5 by test_selector_head_c0_fixture."""
    from models.keypoint_selector import KeypointSelector
    from sslam_b200 import ops
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        g = torch.Generator().manual_seed(31)
        for (B, H, W, hidden) in ((2, 28, 28, 256), (3, 30, 40, 128), (1, 60, 80, 256), (2, 5, 7, 136), (1, 1, 1, 8),
                                  (1, 33, 130, 64)):
            torch.manual_seed(hidden)
            sel = KeypointSelector(384, hidden).to(dev).eval()
            with torch.no_grad():
                sel.conv[0].bias.uniform_(-0.2, 0.2)             # the reference initialises biases to 0
                sel.conv[2].bias.fill_(0.3)
                sel.conv[2].weight.mul_(4.0)
            feat = torch.randn(B, H, W, 384, generator=g).to(dev)
            ops.profile_enable(True)
            with torch.no_grad():
                sal = sel(feat)
                logit = sel(feat, return_logits=True)
            torch.cuda.synchronize()
            kinds = ops.profile_read()
            ops.profile_enable(False)
            assert kinds.get("conv_head", (0, 0))[1] == 2, kinds
            assert sal.shape == (B, H, W, 1) and sal.dtype == torch.float32
            sel.head = "torch"
            with torch.no_grad():
                ref_logit = sel(feat, return_logits=True)
            sel.head = "tcgen05"
            cpu = KeypointSelector(384, hidden).eval()
            cpu.load_state_dict(sel.state_dict())
            with torch.no_grad():
                cpu_logit = cpu.conv(feat.cpu().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
                f64_logit = cpu.double().conv(feat.cpu().double().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
            e_gpu = float((logit - ref_logit).abs().max())
            e_cpu = float((logit.cpu() - cpu_logit).abs().max())
            e_f64 = float((logit.cpu().double() - f64_logit).abs().max())
            scale = max(1.0, float(ref_logit.abs().max()))
            record(f"selector_head.{B}x{H}x{W}x{hidden}", {"max_abs_logit_err_vs_cudnn_fp32": e_gpu,
                                                           "max_abs_logit_err_vs_cpu": e_cpu,
                                                           "max_abs_logit_err_vs_fp64": e_f64,
                                                           "cudnn_fp32_err_vs_fp64": float((ref_logit.cpu().double() - f64_logit).abs().max()),
                                                           "cpu_fp32_err_vs_fp64": float((cpu_logit.double() - f64_logit).abs().max()),
                                                           "logit_abs_max": float(ref_logit.abs().max())})
            tol = 1e-5 * scale
            assert e_gpu < tol and e_cpu < tol and e_f64 < tol, (e_gpu, e_cpu, e_f64, tol)
            assert torch.allclose(sal, torch.sigmoid(ref_logit), rtol=0, atol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_selector_head_c0_fixture(dev):
    """c0 end to end without a library kernel: seeded selector weights + seeded 28x28 features ->
    saliency within 1e-5 of the reference's (tests/golden/decode.npz: c0.sal, written by the
    reference's own KeypointSelector on CPU) and the same 500 keypoints (branch B, duplicates)."""
    from models.keypoint_selector import KeypointSelector
    from test_oracle_golden import DEC
    torch.manual_seed(0)
    sel = KeypointSelector(384, 256).to(dev).eval()
    g = torch.Generator().manual_seed(7)
    feats = torch.randn(2, 28, 28, 384, generator=g).to(dev)
    with torch.no_grad():
        sal = sel(feats)
        kp, sc = sel.select_keypoints(sal, num_keypoints=500)
    err = np.abs(sal[..., 0].cpu().numpy() - DEC["c0.sal"]).max()
    assert err < 1e-5, err
    same = np.array_equal(kp.cpu().numpy(), DEC["c0.kpts"])
    record("selector_head.c0", {"max_abs_saliency_err": float(err), "keypoints_identical": bool(same)})
    assert np.allclose(sc.cpu().numpy(), DEC["c0.scores"], rtol=0, atol=1e-5)
    if not same:
        # a keypoint may only move where two saliency values are closer than the head's arithmetic error
        okp, osc, _ = oracle.select_keypoints(DEC["c0.sal"], 500)
        diff = np.nonzero((kp.cpu().numpy() != okp).any(-1))
        assert np.abs(sc.cpu().numpy()[diff] - osc[diff]).max() < 2e-5
        assert len(diff[0]) < 10
