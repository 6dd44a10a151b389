"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures written by the
reference's own functions, and against the oracle run live on identical seeded inputs.

Bar: bit-exact for keypoint coordinates, scores, indices and match pairs (top-k tie order and
similarity near-ties < 1e-6 excepted and counted); floating-point outputs within the tolerance
written next to each assertion.
"""

import json

import numpy as np
import pytest
import torch

import oracle
from oracle import recipes
from parity import compare_keypoints, compare_matches, load_golden, record
from test_oracle_golden import (DEC, DEC_CASES, DEC_META, MAT, MAT_META, decode_case_input,
                                match_case_input)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from sslam_b200 import _lib
    assert _lib.load().sslam_device_check() == 0, _lib.last_error()
    return torch.device("cuda", 0)


def cu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


# ------------------------------------------------------------------------------------ decode
@pytest.fixture(params=["stream", "register"])
def scan_path(request):
    """Run a decode test with the streaming scan + histogram top-k (default where eligible) and again
    with the register-prefetching scan + radix-select top-k forced (the fallback kernels)."""
    from sslam_b200 import _lib
    _lib.load().sslam_debug_decode_stream(1 if request.param == "stream" else 0)
    yield request.param
    _lib.load().sslam_debug_decode_stream(1)


@pytest.mark.parametrize("name", DEC_CASES)
def test_decode_golden(name, dev, scan_path):
    from sslam_b200 import ops
    sal, m = decode_case_input(name)
    kp, sc, info = ops.decode_topk(cu(sal[None], dev), m["K"], m["nms_radius"], m["pct"])
    kp, sc, info = kp.cpu().numpy()[0], sc.cpu().numpy()[0], info.cpu().numpy()[0]
    _, _, oinfo = oracle.select_keypoints(sal[None], m["K"], m["nms_radius"], m["pct"])
    assert info[0] == oinfo[0, 0], "branch differs from oracle"
    if info[1] >= 0:
        assert info[1] == oinfo[0, 1]
    assert info[2] == oinfo[0, 2], "tie count differs from oracle"
    n = int(oinfo[0, 1])
    compare_keypoints(sal, DEC[name + ".kpts"], DEC[name + ".scores"], kp, sc,
                      sorted_from=n if info[0] in (1, 2) else 0)
    # our own tie rule is deterministic: identical to the oracle, position by position
    okp, osc, _ = oracle.select_keypoints(sal[None], m["K"], m["nms_radius"], m["pct"])
    assert np.array_equal(okp[0], kp) and np.array_equal(osc[0], sc)


def test_decode_mixed_batch(dev, scan_path):
    """Several maps of one size in a single call take different branches independently."""
    from sslam_b200 import ops
    maps = [recipes.spread_saliency(48, 64, 101), np.full((48, 64), 0.05, np.float32),
            recipes.box_saliency(48, 64, 102, quant=16), recipes.spread_saliency(48, 64, 103)]
    three = np.full((48, 64), 0.2, np.float32)
    three[5, 5], three[20, 30], three[40, 60] = 0.9, 0.8, 0.7
    maps.append(three)
    batch = np.stack(maps)
    for K in (16, 64, 600):
        kp, sc, info = ops.decode_topk(cu(batch, dev), K)
        okp, osc, oinfo = oracle.select_keypoints(batch, K)
        assert np.array_equal(info.cpu().numpy()[:, 0], oinfo[:, 0])
        assert np.array_equal(kp.cpu().numpy(), okp)
        assert np.array_equal(sc.cpu().numpy(), osc)
    assert len(set(oinfo[:, 0].tolist())) >= 2


def test_decode_native_grid_c0(dev):
    from models.keypoint_selector import KeypointSelector
    sel = KeypointSelector(384, 256).to(dev)
    kp, sc = sel.select_keypoints(cu(DEC["c0.sal"], dev)[..., None], num_keypoints=500)
    assert kp.shape == (2, 500, 2) and kp.dtype == torch.float32
    assert np.array_equal(kp.cpu().numpy(), DEC["c0.kpts"])
    assert np.array_equal(sc.cpu().numpy(), DEC["c0.scores"])


def test_decode_raises_like_reference(dev):
    from models.keypoint_selector import KeypointSelector
    sel = KeypointSelector(384, 256).to(dev)
    sal = cu(recipes.spread_saliency(30, 40, 5)[None, :, :, None], dev)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        sel.select_keypoints(sal, num_keypoints=2048)


def test_decode_from_logits(dev):
    from sslam_b200 import ops
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(2, 64, 96, generator=g) * 2).to(dev)
    kp, sc, _ = ops.decode_topk(logits, 50, from_logits=True)
    sal = torch.sigmoid(logits)
    # every score is the sigmoid of the logit at its keypoint (<= 2 ulp: device expf vs torch)
    xs, ys = kp[..., 0].long(), kp[..., 1].long()
    picked = torch.stack([sal[b, ys[b], xs[b]] for b in range(2)])
    assert torch.allclose(picked, sc, rtol=3e-7, atol=0)
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())


def test_nms_golden(dev):
    from models.keypoint_selector import KeypointSelector
    sel = KeypointSelector(8, 8)
    x = cu(DEC["nms.in"][None], dev)
    for r in range(4):
        out = sel._apply_nms(x, r)
        assert np.array_equal(out.cpu().numpy()[0], DEC[f"nms.r{r}"])


def test_decode_speculative_count_is_only_a_hint(dev):
    """The scan counts pixels below a per-slot hint left by the previous call (0.95 x that slot's last
    K-th score) so that the second pass over the map can be skipped.  Whatever the hint holds — nothing
    (first call), a good value (same data again), a value far too high or too low (different data in the
    same workspace slot) — keypoints, scores and the branch taken must equal the oracle's."""
    from sslam_b200 import ops
    H, W, K = 96, 256, 64
    ws = ops.Workspace()
    lib_need = ops._lib.load().sslam_decode_workspace_bytes(3, H, W, K)
    buf = ws.get(lib_need, dev)
    buf.zero_()
    hi = np.stack([recipes.spread_saliency(H, W, 301 + i, lo=0.6, hi=0.99) for i in range(3)])
    lo = np.stack([recipes.spread_saliency(H, W, 311 + i, lo=0.02, hi=0.3) for i in range(3)])
    mid = np.stack([recipes.box_saliency(H, W, 321 + i) for i in range(3)])
    for maps in (hi, hi, lo, lo, mid, hi, mid, mid):
        kp, sc, info = ops.decode_topk(cu(maps, dev), K, workspace=buf)
        okp, osc, oinfo = oracle.select_keypoints(maps, K)
        assert np.array_equal(info.cpu().numpy()[:, 0], oinfo[:, 0])
        assert np.array_equal(kp.cpu().numpy(), okp) and np.array_equal(sc.cpu().numpy(), osc)
    buf.fill_(0xff)                                              # garbage (NaN) hints
    kp, sc, info = ops.decode_topk(cu(mid, dev), K, workspace=buf)
    okp, osc, oinfo = oracle.select_keypoints(mid, K)
    assert np.array_equal(kp.cpu().numpy(), okp) and np.array_equal(sc.cpu().numpy(), osc)


def test_decode_ring_and_band_shape_do_not_change_results(dev):
    """The streaming scan's ring depth and band height are scheduling choices: whatever they are — two
    stages and 64-row bands (the default), a deep ring, bands of a few rows whose halos overlap, one band per
    image, a staging list that overflows into the global list (plateau maps, tall bands) — keypoints,
    scores and the branch taken equal the oracle's."""
    from sslam_b200 import ops, _lib
    lib = _lib.load()
    H, W, K = 400, 512, 256                               # ~8 000 local maxima per map: one band overflows the staging list
    maps = np.stack([recipes.spread_saliency(H, W, 401, lo=0.2, hi=0.99), recipes.box_saliency(H, W, 402),
                     recipes.spread_saliency(H, W, 403, lo=0.02, hi=0.4)])
    okp, osc, oinfo = oracle.select_keypoints(maps, K)
    try:
        for stages, band in ((0, 0), (2, 64), (7, 128), (16, 8), (3, 5), (2, 4096), (5, 33)):
            lib.sslam_debug_decode_tune(stages, band)
            kp, sc, info = ops.decode_topk(cu(maps, dev), K)
            assert np.array_equal(info.cpu().numpy()[:, 0], oinfo[:, 0]), (stages, band)
            assert np.array_equal(sc.cpu().numpy(), osc), (stages, band)
            cmp_kp = kp.cpu().numpy()
            same = np.array_equal(cmp_kp, okp)
            if not same:                                  # equal scores may be ordered by index on either side
                for b in range(maps.shape[0]):
                    assert sorted(map(tuple, cmp_kp[b])) == sorted(map(tuple, okp[b])), (stages, band, b)
    finally:
        lib.sslam_debug_decode_tune(0, 0)


def test_decode_full_size_properties(dev, scan_path):
    """BASELINE sizes (640x480 K=2048, 1280x960 K=8192) on the seeded synthetic sequence:
    identical to the oracle, plus size-independent properties."""
    from sslam_b200 import ops, synth
    for (H, W, K, T) in ((480, 640, 2048, 3), (960, 1280, 8192, 1)):
        sal, _ = synth.make_sequence(T, seq_id=0, height=H, width=W)
        kp, sc, info = ops.decode_topk(sal.to(dev), K)
        kp, sc, info = kp.cpu().numpy(), sc.cpu().numpy(), info.cpu().numpy()
        okp, osc, oinfo = oracle.select_keypoints(sal.numpy(), K)
        assert np.array_equal(info[:, 0], oinfo[:, 0]) and (info[:, 0] == 0).all()
        assert np.array_equal(kp, okp) and np.array_equal(sc, osc)
        s = sal.numpy()[..., 0]
        for b in range(T):
            assert (np.diff(sc[b]) <= 0).all()                          # sorted
            lin = kp[b, :, 1].astype(np.int64) * W + kp[b, :, 0].astype(np.int64)
            assert np.unique(lin).size == K                              # no duplicates
            assert np.array_equal(s[b].ravel()[lin], sc[b])              # scores are map values
            assert (sc[b] > max(np.median(s[b]), 0.1)).all()             # above the threshold
            nms = oracle.apply_nms(s[b], 2).ravel()
            assert (nms[lin] > 0).all()                                  # all are NMS survivors


# ------------------------------------------------------------------------------------ gather / norm
def test_gather_golden(dev):
    from models.dino_backbone import DinoBackbone
    z, _ = load_golden("gather")
    bb = DinoBackbone(load_vit=False).to(dev)
    out = bb.extract_at_keypoints(cu(z["small.feat"], dev), cu(z["small.kpts"], dev)).cpu().numpy()
    assert np.abs(out - z["small.out"]).max() <= 1e-6
    assert np.array_equal(out, z["small.out"]), "gather is expected to be bit-identical to ATen CPU"
    feat = recipes.int_features(1, 30, 40, 384, 43)
    pc = bb.pixel_to_patch(cu(z["tum.pix"], dev))
    assert np.array_equal(pc.cpu().numpy(), z["tum.patch"])
    assert np.array_equal(bb.patch_to_pixel(pc).cpu().numpy(), z["tum.back"])
    out = bb.extract_at_keypoints(cu(feat, dev), pc).cpu().numpy()
    assert np.array_equal(out, z["tum.out"])
    fused = bb.extract_at_pixel_keypoints(cu(feat, dev), cu(z["tum.pix"], dev)).cpu().numpy()
    assert np.array_equal(fused, z["tum.out"]), "fused pixel_to_patch differs"
    out = bb.extract_at_keypoints(cu(z["native.feat"], dev), cu(z["native.kpts"], dev)).cpu().numpy()
    assert np.array_equal(out, z["native.out"])


def test_gather_unaligned_channels(dev):
    from sslam_b200 import ops
    feat = recipes.int_features(2, 5, 6, 10, 7)          # C % 4 != 0 -> scalar path
    rng = np.random.Generator(np.random.PCG64(8))
    kp = np.stack([rng.integers(-32, 6 * 32 + 32, size=(2, 33)) / 32.0,
                   rng.integers(-32, 5 * 32 + 32, size=(2, 33)) / 32.0], -1).astype(np.float32)
    out = ops.gather_bilinear(cu(feat, dev), cu(kp, dev)).cpu().numpy()
    assert np.array_equal(out, oracle.extract_at_keypoints(feat, kp))


def test_refiner_and_normalize_golden(dev):
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import ops
    z, _ = load_golden("refiner")
    m = DescriptorRefiner(input_dim=48, hidden_dim=64, output_dim=32, num_layers=4)
    m.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")})
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m(cu(z["x"], dev)).cpu().numpy()
    assert out.shape == z["out"].shape
    assert np.abs(out - z["out"]).max() < 1e-5                          # descriptors within 1e-5 abs
    nrm, nrm16 = ops.l2norm_rows(cu(z["norm.in"], dev), want_bf16=True)
    assert np.allclose(nrm.cpu().numpy(), z["norm.out"], rtol=1e-6, atol=1e-7)
    assert bool((nrm[3] == 0).all())
    assert torch.allclose(nrm16.float(), nrm, rtol=8e-3, atol=1e-6)     # bf16 rounding
    odd = cu(recipes.int_features(1, 1, 5, 10, 9).reshape(5, 10), dev)  # D % 4 != 0
    assert np.allclose(ops.l2norm_rows(odd).cpu().numpy(), oracle.l2_normalize(odd.cpu().numpy()),
                       rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------ matching
def _top_cpu(top, p=0):
    return {k: v[p].cpu().numpy() for k, v in top.items()}


@pytest.mark.parametrize("name", list(MAT_META))
def test_matchers_golden(name, dev):
    from sslam_b200 import matchers
    d1, d2, m = match_case_input(name)
    if d1.shape[1] % 4:
        pytest.skip("D % 4 != 0")
    S = (d1.astype(np.float64) @ d2.astype(np.float64).T)
    s1, s2, i1, i2 = (MAT[f"{name}.{k}"] for k in ("s1", "s2", "i1", "i2"))
    exc = 0

    got = matchers.find_matches(d1, d2, 0.8)
    assert all(isinstance(t, tuple) and len(t) == 3 for t in got)
    gp = np.array([(i, j) for i, j, _ in got], dtype=np.int64).reshape(-1, 2)
    assert (np.diff(gp[:, 0]) > 0).all()                                  # ascending i
    exc += compare_matches(S, MAT[name + ".m1"], gp)
    if gp.shape == MAT[name + ".m1"].shape and np.array_equal(gp, MAT[name + ".m1"]):
        # default arithmetic = f16x3 on tcgen05: similarities within 3e-6 abs of fp64 (the tensor core
        # truncates when it accumulates in fp32; tests/test_gpu_tc.py::test_f16x3_top2)
        assert np.allclose([s for _, _, s in got], MAT[name + ".m1s"], rtol=0, atol=3e-6)
    second = lambda i: np.partition(S[i], -2)[-2] if S.shape[1] > 1 else -1.0  # noqa: E731
    got = matchers.find_matches(d1, d2, 1.02)
    exc += compare_matches(S, MAT[name + ".m1b"], np.array([(i, j) for i, j, _ in got]).reshape(-1, 2),
                           threshold_margin=lambda i, j: abs(S[i, j] - 1.02 * max(second(i), -1.0)))

    mm, qq = matchers.match_with_quality(d1, d2, s1, s2)
    assert mm.dtype == np.int64 and qq.dtype == np.float32 and mm.ndim == 2 and mm.shape[1] == 2
    exc += compare_matches(S, MAT[name + ".m2"], mm, threshold_margin=lambda i, j: abs(S[i, j] - 0.7))
    if np.array_equal(mm, MAT[name + ".m2"]):
        assert np.allclose(qq, MAT[name + ".m2q"], rtol=0, atol=3e-6)
    mi, _ = matchers.match_with_quality(d1, d2, s1, s2, intensity1=i1, intensity2=i2,
                                        min_intensity=0.15, min_saliency=0.5)
    exc += compare_matches(S, MAT[name + ".m2i"], mi, threshold_margin=lambda i, j: abs(S[i, j] - 0.7))
    me, qe = matchers.match_with_quality(d1, d2, s1, s2, min_descriptor_sim=1.5)
    assert me.shape == (0, 2) and me.dtype == np.int64 and qe.shape == (0,) and qe.dtype == np.float32

    if m["m"] >= 2:
        m3, dist = matchers.find_mutual_nearest_neighbors(d1, d2, 0.9)
        ratio_margin = lambda thr: (lambda i, j: abs(second(i) / (S[i, j] + 1e-8) - thr))  # noqa: E731
        exc += compare_matches(S, MAT[name + ".m3"], m3, threshold_margin=ratio_margin(0.9))
        if np.array_equal(m3, MAT[name + ".m3"]):
            assert np.allclose(dist, MAT[name + ".m3d"], rtol=0, atol=3e-6)
        m3b, _ = matchers.find_mutual_nearest_neighbors(d1, d2, 0.98)
        exc += compare_matches(S, MAT[name + ".m3b"], m3b, threshold_margin=ratio_margin(0.98))
    if m["n"] == m["m"]:
        b1 = cu(np.stack([d1, d2[::-1].copy()]), dev)
        b2 = cu(np.stack([d2, d1]), dev)
        o4 = matchers.find_matches_batched(b1, b2).cpu().numpy()
        ref4 = MAT[name + ".m4"]
        assert o4.dtype == np.int64
        if o4.shape == ref4.shape and np.array_equal(o4, ref4):
            pass
        else:
            for b in range(2):
                Sb = (np.stack([d1, d2[::-1]])[b].astype(np.float64) @ np.stack([d2, d1])[b].astype(np.float64).T)
                strip = lambda a: a[(a != 0).any(axis=1) | (np.arange(a.shape[0]) == 0)]  # noqa: E731
                exc += compare_matches(Sb, strip(ref4[b]), strip(o4[b]))
    assert matchers.tracking_count(d1, d2, 0.8) == int(MAT[name + ".m5"])
    assert matchers.tracking_count(d1, d2, 0.5) == int(MAT[name + ".m5lo"])
    record("matchers_golden." + name, {"near_tie_exceptions": int(exc)})


def test_match_top2_primitive_vs_oracle(dev):
    """Row top-2 / column argmax incl. ragged sizes, exact ties and pair indexing."""
    from sslam_b200 import ops
    cases = [(130, 257, 32), (128, 128, 256), (1, 5, 16), (300, 1, 64), (77, 129, 100)]
    for (n, m, d) in cases:
        d1, d2, _ = recipes.descriptor_pair(n, m, d, 200 + n, noise=2, dup_every=9)
        top = ops.match_top2(cu(d1[None], dev), cu(d2[None], dev), mode=ops.SIM_F32)
        t = _top_cpu(top)
        nn12, best12, second12, nn21, best21, S = oracle.similarity_top2(d1, d2)
        S64 = d1.astype(np.float64) @ d2.astype(np.float64).T
        # values within fp32 accumulation error; indices equal unless a near tie
        assert np.allclose(t["best12"], best12, rtol=0, atol=2e-6)
        assert np.allclose(t["best21"], best21, rtol=0, atol=2e-6)
        if m > 1:
            assert np.allclose(t["second12"], second12, rtol=0, atol=2e-6)
        else:
            assert np.isneginf(t["second12"]).all()
        for i in np.nonzero(t["nn12"] != nn12)[0]:
            assert abs(S64[i, t["nn12"][i]] - S64[i, nn12[i]]) < 1e-6
        for j in np.nonzero(t["nn21"] != nn21)[0]:
            assert abs(S64[t["nn21"][j], j] - S64[nn21[j], j]) < 1e-6
    # exact duplicates: lowest index must win in both directions
    d1 = np.zeros((6, 8), np.float32); d1[:, 0] = 1
    d2 = np.zeros((5, 8), np.float32); d2[:, 0] = 1
    for mode in (ops.SIM_F32, ops.SIM_F16X3):                 # the default (f16x3) keeps the tie rule
        t = _top_cpu(ops.match_top2(cu(d1[None], dev), cu(d2[None], dev), mode=mode))
        assert (t["nn12"] == 0).all() and (t["nn21"] == 0).all()
        assert (t["best12"] == 1).all() and (t["second12"] == 1).all()
    # pair_index into banks
    bank1 = np.stack([recipes.descriptor_pair(64, 64, 32, s)[0] for s in range(4)])
    bank2 = np.stack([recipes.descriptor_pair(64, 80, 32, s)[1] for s in range(3)])
    idx = torch.tensor([[3, 0], [1, 2], [0, 0]], dtype=torch.int32, device=dev)
    top = ops.match_top2(cu(bank1, dev), cu(bank2, dev), pair_index=idx, mode=ops.SIM_F32)
    for p, (a, b) in enumerate(idx.cpu().tolist()):
        nn12, best12, *_ = oracle.similarity_top2(bank1[a], bank2[b])
        assert np.array_equal(top["nn12"][p].cpu().numpy(), nn12)
        assert np.allclose(top["best12"][p].cpu().numpy(), best12, atol=2e-6)


# ------------------------------------------------------------------------------------ end to end
def test_sequence_pipeline_vs_oracle(dev):
    """c2-shaped slice: 4 frames 640x480, K=2048, D=256, fp32 mode; keypoints identical, descriptors
    within 1e-5 abs, consecutive-pair M1 and M2 lists identical up to counted near ties."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 256, 4).to(dev)
    T, K = 4, 2048
    sal, feat = synth.make_sequence(T, seq_id=0)
    fe = FrontEnd(refiner, num_keypoints=K, grid="pixel")
    feats, pairs, pscores, counts = fe.run_sequence(sal.to(dev), feat.to(dev), matchers.M1, chunk=3)
    kp = feats["keypoints_pixel"].cpu().numpy()
    okp, osc, oinfo = oracle.select_keypoints(sal.numpy(), K)
    assert np.array_equal(kp, okp) and np.array_equal(feats["scores"].cpu().numpy(), osc)
    g = oracle.extract_at_keypoints(feat.numpy(), oracle.pixel_to_patch(okp))
    w = oracle.RefinerWeights.from_state_dict(refiner.state_dict())
    od = oracle.refiner_forward(w, g)
    d = feats["descriptors"].cpu().numpy()
    assert np.abs(od - d).max() < 1e-5
    assert np.allclose(np.linalg.norm(d, axis=-1), 1.0, atol=1e-5)
    p2, q2, c2 = fe.match_consecutive(feats, matchers.M2)
    total_exc = 0
    for p in range(T - 1):
        S = d[p].astype(np.float64) @ d[p + 1].astype(np.float64).T
        ref = oracle.match_m1(d[p], d[p + 1], 0.8)
        got = pairs[p, :int(counts[p])].cpu().numpy()
        assert (pairs[p, int(counts[p]):] == -1).all()
        total_exc += compare_matches(S, np.array([(i, j) for i, j, _ in ref]).reshape(-1, 2), got)
        rm, rq = oracle.match_m2(d[p], d[p + 1], osc[p], osc[p + 1])
        gm = p2[p, :int(c2[p])].cpu().numpy()
        total_exc += compare_matches(S, rm, gm, threshold_margin=lambda i, j: abs(S[i, j] - 0.7))
        if np.array_equal(rm, gm):
            assert np.allclose(q2[p, :int(c2[p])].cpu().numpy(), rq, rtol=0, atol=3e-6)
        assert int(counts[p]) > K // 4, "overlapping frames should match plentifully"
    record("sequence_pipeline_vs_oracle", {"pairs": T - 1, "near_tie_exceptions": int(total_exc)})


def test_abi_error_codes(dev):
    import ctypes
    from sslam_b200 import _lib
    l = _lib.load()
    z = ctypes.c_void_p(0)
    buf = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    p = ctypes.c_void_p(buf.data_ptr())
    assert l.sslam_decode_topk_f32(z, 0, 1, 8, 8, 4, 2, 0.5, 0.1, z, z, z, z, 0, z) == -1       # null
    assert l.sslam_decode_topk_f32(p, 0, 1, 8, 8, 4, 99, 0.5, 0.1, p, p, p, p, 1 << 20, z) == -2  # radius
    assert l.sslam_decode_topk_f32(p, 0, 1, 8, 8, 4, 2, 0.5, 0.1, p, p, p, p, 8, z) == -3        # workspace
    assert l.sslam_decode_topk_f32(p, 0, 1, 8, 8, 4, 2, 0.5, -1.0, p, p, p, p, 1 << 20, z) == -1  # floor
    assert "floor" in _lib.last_error()
    assert l.sslam_decode_topk_f32(p, 0, 0, 8, 8, 4, 2, 0.5, 0.1, p, p, p, p, 0, z) == 0         # empty batch
    assert l.sslam_match_top2(p, z, 1, p, z, 1, z, 0, 1, 4, 4, 6, p, p, p, p, p, p, 1 << 20, z) == -2  # D % 4
    assert l.sslam_match_top2(p, z, 1, p, z, 1, z, 7, 1, 4, 4, 8, p, p, p, p, p, p, 1 << 20, z) == -1  # dtype
    assert l.sslam_match_top2(p, p, 1, p, z, 1, z, 3, 1, 4, 4, 8, p, p, p, p, p, p, 1 << 20, z) == -1  # one lo only
    assert l.sslam_match_top2(p, p, 1, p, p, 1, z, 0, 1, 4, 4, 8, p, p, p, p, p, p, 1 << 20, z) == -1  # pairs need f16x3
    assert l.sslam_launch_count() > 0
    torch.cuda.synchronize()


def test_cuda_graph_replay_matches_eager(dev):
    """FrontEnd.capture_sequence: a replayed CUDA graph of the whole step gives the eager results,
    and follows new input data copied into the static buffers."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 128, 4).to(dev)
    fe = FrontEnd(refiner, num_keypoints=256, grid="pixel")
    salA, featA = synth.make_sequence(5, seq_id=3, height=96, width=128)
    salB, featB = synth.make_sequence(5, seq_id=4, height=96, width=128)
    sal, feat = salA.to(dev).clone(), featA.to(dev).clone()
    replay, feats, pairs, pscores, counts = fe.capture_sequence(sal, feat, matchers.M1, chunk=2)
    for s_src, f_src in ((salA, featA), (salB, featB), (salA, featA)):
        sal.copy_(s_src.to(dev)); feat.copy_(f_src.to(dev))
        replay()
        torch.cuda.synchronize()
        e_feats, e_pairs, e_pscores, e_counts = fe.run_sequence(s_src.to(dev), f_src.to(dev), matchers.M1, chunk=2)
        assert torch.equal(feats["keypoints_pixel"], e_feats["keypoints_pixel"])
        assert torch.equal(feats["descriptors"], e_feats["descriptors"])
        assert torch.equal(counts, e_counts) and torch.equal(pairs, e_pairs)
        assert torch.equal(pscores, e_pscores)


def test_host_streaming_pipeline_matches_device_pipeline(dev):
    """FrontEnd.run_sequence_host (pinned host buffers streamed chunk by chunk, pairs matched as
    their frames arrive, lists copied back) returns exactly what the device-resident pass returns,
    for chunk sizes that do and do not divide the sequence."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 128, 4).to(dev)
    fe = FrontEnd(refiner, num_keypoints=256, grid="pixel")
    T = 7
    sal, feat = synth.make_sequence(T, seq_id=5, height=96, width=128)
    _, pairs, pscores, counts = fe.run_sequence(sal.to(dev), feat.to(dev), matchers.M1, chunk=T, ratio_thresh=0.8)
    sal_h, feat_h = sal.pin_memory(), feat.pin_memory()
    for chunk in (1, 3, 7, 16):
        hp, hs, hc = fe.run_sequence_host(sal_h, feat_h, matchers.M1, chunk=chunk, ratio_thresh=0.8)
        assert hp.shape == (T - 1, 256, 2) and hc.shape == (T - 1,)
        assert torch.equal(hc, counts.cpu()), chunk
        assert torch.equal(hp, pairs.cpu()), chunk
        assert torch.equal(hs, pscores.cpu()), chunk


def test_refiner_back_to_back_launches_are_stable(dev):
    """Regression test for the cluster pipeline of the refiner GEMM: 400 forwards of a 50-frame chunk
    (102 400 rows, the e2e staging size) enqueued back to back must all complete and give bit-identical
    descriptors.  (A producer that could be overtaken by two phases of a stage's `empty` barrier used
    to deadlock about one launch in 300 in the residual layers.)"""
    from models.descriptor_refiner import DescriptorRefiner
    torch.manual_seed(0)
    m = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
    x = torch.randn(1, 102400, 384, device=dev)
    with torch.no_grad():
        ref = m.forward_fused(x).clone()
        for it in range(400):
            y = m.forward_fused(x)
            if it % 100 == 99:
                assert torch.equal(y, ref), it
    torch.cuda.synchronize()


def test_empty_and_degenerate_shapes(dev):
    """Empty batches / zero keypoints / tiny maps go through the C ABI without launching."""
    from sslam_b200 import ops
    kp, sc, info = ops.decode_topk(torch.rand(0, 16, 16, device=dev), 8)
    assert kp.shape == (0, 8, 2) and sc.shape == (0, 8)
    kp, sc, info = ops.decode_topk(torch.rand(2, 16, 16, device=dev), 0)
    assert kp.shape == (2, 0, 2)
    out = ops.gather_bilinear(torch.rand(1, 4, 4, 8, device=dev), torch.zeros(1, 0, 2, device=dev))
    assert out.shape == (1, 0, 8)
    # 1x1 map: everything is the single pixel (reference: NMS keeps it, branch by threshold)
    one = torch.full((1, 1, 1), 0.7, device=dev)
    kp, sc, info = ops.decode_topk(one, 1)
    okp, osc, oinfo = oracle.select_keypoints(one.cpu().numpy(), 1)
    assert np.array_equal(kp.cpu().numpy(), okp) and np.array_equal(sc.cpu().numpy(), osc)
    assert int(info[0, 0]) == int(oinfo[0, 0])
    # a map with rows shorter than a warp strip and K equal to the pixel count
    m = cu(recipes.spread_saliency(5, 7, 91)[None], dev)
    kp, sc, info = ops.decode_topk(m, 35)
    okp, osc, oinfo = oracle.select_keypoints(m.cpu().numpy(), 35)
    assert np.array_equal(kp.cpu().numpy(), okp) and np.array_equal(sc.cpu().numpy(), osc)
    with pytest.raises(RuntimeError):
        ops.decode_topk(torch.rand(1, 8, 8), 4)                       # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        ops.match_top2(torch.rand(1, 4, 6, device=dev), torch.rand(1, 4, 6, device=dev), mode=ops.SIM_F32)   # D % 4
    with pytest.raises(RuntimeError):
        ops.match_top2(torch.rand(1, 4, 12, device=dev), torch.rand(1, 4, 12, device=dev))   # f16x3 default: D % 8


def test_decode_all_radii_and_percentiles(dev, scan_path):
    """Strip kernel (radius 1..3) and tiled kernel (radius 0, 4..8) against the oracle."""
    from sslam_b200 import ops
    sal = np.stack([recipes.spread_saliency(70, 150, 200 + i) for i in range(3)] +
                   [recipes.box_saliency(70, 150, 210, quant=24)])
    for r in (0, 1, 2, 3, 4, 6):
        for pct in (0.5, 0.9, 0.25):
            kp, sc, info = ops.decode_topk(cu(sal, dev), 60, nms_radius=r, min_score_percentile=pct)
            okp, osc, oinfo = oracle.select_keypoints(sal, 60, r, pct)
            assert np.array_equal(info.cpu().numpy()[:, 0], oinfo[:, 0]), (r, pct)
            assert np.array_equal(kp.cpu().numpy(), okp), (r, pct)
            assert np.array_equal(sc.cpu().numpy(), osc), (r, pct)


def test_end_to_end_index_agreement_oracle_own_descriptors(dev):
    """north_star: "100 % index agreement in fp32 mode (near ties within 1e-6 excepted and counted)".

    The oracle runs END TO END on its own descriptors (its own decode, sampling, refiner MLP and
    matcher) for 24 frames of the c2 workload; the GPU runs the default pipeline (f16x3 refiner +
    f16x3 matcher).  Every (i, j) that appears in one M1 list and not the other is reported with
    its deciding margin in the oracle's similarity matrix.  A decision can only move if that margin
    is below twice the similarity error, which is MEASURED here as max |S_gpu - S_oracle| per pair
    (descriptor error of the refiner) plus the matcher's own arithmetic error (2e-6, f16x3 with
    truncating fp32 accumulation, tests/test_gpu_tc.py::test_f16x3_top2)."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 256, 4).to(dev)
    T, K = 24, 2048
    sal, feat = synth.make_sequence(T, seq_id=0)
    fe = FrontEnd(refiner, num_keypoints=K, grid="pixel")                  # defaults: f16x3
    feats, pairs, pscores, counts = fe.run_sequence(sal.to(dev), feat.to(dev), matchers.M1, chunk=8,
                                                    ratio_thresh=0.8)
    w = oracle.RefinerWeights.from_state_dict(refiner.state_dict())
    okp, osc, _ = oracle.select_keypoints(sal.numpy(), K)
    assert np.array_equal(feats["keypoints_pixel"].cpu().numpy(), okp)     # keypoints bit-exact
    od = oracle.refiner_forward(w, oracle.extract_at_keypoints(feat.numpy(), oracle.pixel_to_patch(okp)))
    gd = feats["descriptors"].cpu().numpy()
    desc_err = float(np.abs(od - gd).max())
    assert desc_err < 1e-5
    MATCHER_ERR = 2e-6
    total_ref = total_diff = 0
    worst_margin = 0.0
    worst_bound = 0.0
    details = []
    for p in range(T - 1):
        So = od[p].astype(np.float64) @ od[p + 1].astype(np.float64).T
        Sg = gd[p].astype(np.float64) @ gd[p + 1].astype(np.float64).T
        eps = float(np.abs(So - Sg).max())
        bound = 2 * (eps + MATCHER_ERR)
        ref = {(i, j) for i, j, _ in oracle.match_m1(od[p], od[p + 1], 0.8)}
        got = {tuple(r) for r in pairs[p, :int(counts[p])].cpu().numpy().tolist()}
        total_ref += len(ref)
        part = -np.partition(-So, 1, axis=1)
        row_margin = part[:, 0] - part[:, 1]
        partc = -np.partition(-So.T, 1, axis=1)
        col_margin = partc[:, 0] - partc[:, 1]
        for (i, j) in ref ^ got:
            jj = int(np.argmax(So[i]))
            ii = int(np.argmax(So[:, j]))
            margin = float(min(row_margin[i], col_margin[j], col_margin[jj], row_margin[ii]))
            details.append({"pair": p, "i": int(i), "j": int(j), "margin": margin, "bound": bound})
            assert margin < bound, f"pair {p}: ({i},{j}) differs with margin {margin:.3g} >= bound {bound:.3g}"
            worst_margin = max(worst_margin, margin)
            total_diff += 1
        worst_bound = max(worst_bound, bound)
    agreement = 1.0 - total_diff / max(total_ref, 1)
    record("end_to_end_oracle_own_descriptors", {
        "frames": T, "pairs": T - 1, "K": K, "matcher": "M1 ratio 0.8, f16x3 (default)",
        "oracle_matches": total_ref, "disagreements": total_diff, "index_agreement": agreement,
        "max_descriptor_abs_err": desc_err, "max_disagreement_margin": worst_margin,
        "max_allowed_margin": worst_bound, "disagreements_detail": details[:50]})
    assert agreement > 0.9995
