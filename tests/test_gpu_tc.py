"""GPU parity of the tensor-core similarity kernels (tcgen05 / TMEM / TMA) against the oracle.

fp32 mode (SSLAM_SIM_TF32X3): match index pairs identical to the oracle except similarity
near-ties < 1e-6 (counted); similarity values within 3e-6 abs.
bf16 mode (SSLAM_SIM_BF16): similarity values within 1e-5 abs of the fp64 product of the same
bf16-rounded inputs, indices identical except near ties of that product; against the fp32 oracle on
the un-rounded inputs the scores agree to bf16 input rounding.
"""

import numpy as np
import pytest
import torch

import oracle
from oracle import recipes
from parity import compare_matches, record

pytestmark = pytest.mark.gpu

CASES = [(128, 128, 256), (130, 257, 32), (2048, 2048, 256), (1, 5, 64), (300, 1, 64), (77, 129, 96),
         (500, 500, 128)]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def cu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def check_top(t, S64, tol_val, tol_tie):
    n, m = S64.shape
    nn12 = S64.argmax(1)
    nn21 = S64.argmax(0)
    best12 = S64[np.arange(n), nn12]
    best21 = S64[nn21, np.arange(m)]
    assert np.abs(t["best12"] - best12).max() <= tol_val
    assert np.abs(t["best21"] - best21).max() <= tol_val
    if m > 1:
        second = -np.partition(-S64, 1, axis=1)[:, 1]
        assert np.abs(t["second12"] - second).max() <= tol_val
    else:
        assert np.isneginf(t["second12"]).all()
    exc = 0
    for i in np.nonzero(t["nn12"] != nn12)[0]:
        assert abs(S64[i, t["nn12"][i]] - S64[i, nn12[i]]) < tol_tie, "row argmax differs beyond a near tie"
        exc += 1
    for j in np.nonzero(t["nn21"] != nn21)[0]:
        assert abs(S64[t["nn21"][j], j] - S64[nn21[j], j]) < tol_tie, "column argmax differs beyond a near tie"
        exc += 1
    return exc


@pytest.mark.parametrize("n,m,d", CASES)
def test_tf32x3_top2(n, m, d, dev):
    from sslam_b200 import ops
    d1, d2, _ = recipes.descriptor_pair(n, m, d, 300 + n + m, noise=3, dup_every=11)
    top = ops.match_top2(cu(d1[None], dev), cu(d2[None], dev), mode=ops.SIM_TF32X3)
    t = {k: v[0].cpu().numpy() for k, v in top.items()}
    S64 = d1.astype(np.float64) @ d2.astype(np.float64).T
    # tensor-core fp32 accumulation truncates (~2e-6 low at |S|~1, D=256); decisions are unaffected
    # beyond the 1e-6 near-tie band because the bias is common to neighbouring values
    exc = check_top(t, S64, 3e-6, 1e-6)
    record(f"tf32x3_top2.{n}x{m}x{d}", {"near_tie_index_exceptions": int(exc),
                                        "max_abs_best_err": float(np.abs(t['best12'] - S64.max(1)).max())})


@pytest.mark.parametrize("n,m,d", [c for c in CASES if c[2] % 8 == 0])
def test_f16x3_top2(n, m, d, dev):
    """fp32 mode on the 16-bit tensor pipe: x = hi + lo*2^-11 (fp16 pair), three MMAs."""
    from sslam_b200 import ops
    d1, d2, _ = recipes.descriptor_pair(n, m, d, 300 + n + m, noise=3, dup_every=11)
    top = ops.match_top2(cu(d1[None], dev), cu(d2[None], dev), mode=ops.SIM_F16X3)
    t = {k: v[0].cpu().numpy() for k, v in top.items()}
    S64 = d1.astype(np.float64) @ d2.astype(np.float64).T
    exc = check_top(t, S64, 3e-6, 1e-6)
    record(f"f16x3_top2.{n}x{m}x{d}", {"near_tie_index_exceptions": int(exc),
                                       "max_abs_best_err": float(np.abs(t['best12'] - S64.max(1)).max())})


def test_match_back_to_back_launches_are_stable(dev):
    """The CTA-pair matcher enqueued 300 times back to back (odd strip count, so one CTA of the last
    pair of every set idles through the same barrier protocol) completes and stays bit-identical."""
    from sslam_b200 import ops
    F, N, D = 6, 1100, 256                                   # 9 strips -> 5 strip pairs per set
    g = torch.Generator(device="cpu").manual_seed(5)
    bank = torch.nn.functional.normalize(torch.randn(F, N, D, generator=g), dim=-1).to(dev)
    ref = ops.match_top2(bank, bank[1:], mode=ops.SIM_F16X3, num_pairs=F - 1)
    ref = {k: v.clone() for k, v in ref.items()}
    for it in range(300):
        out = ops.match_top2(bank, bank[1:], mode=ops.SIM_F16X3, num_pairs=F - 1)
        if it % 100 == 99:
            for k in ref:
                assert torch.equal(out[k], ref[k]), (it, k)
    torch.cuda.synchronize()
    chk = ops.match_top2(bank, bank[1:], mode=ops.SIM_F32, num_pairs=F - 1)
    assert (chk["nn12"] == ref["nn12"]).float().mean().item() > 0.999


def test_f16x3_presplit_banks_equal_internal_split(dev):
    """The fp16 (hi, lo) pair written by l2norm_rows (what the pipeline hands to the matcher) gives
    bit-identical top-2 results to the matcher splitting the fp32 bank itself, and the pair
    reconstructs the fp32 descriptors to 2^-22 relative."""
    from sslam_b200 import ops
    F, N, D = 4, 300, 256
    g = torch.Generator(device="cpu").manual_seed(11)
    raw = torch.randn(F, N, D, generator=g).to(dev)
    hi = torch.empty(F, N, D, dtype=torch.float16, device=dev)
    lo = torch.empty_like(hi)
    d32 = ops.l2norm_rows(raw, pair=(hi, lo))
    assert torch.equal(d32, ops.l2norm_rows(raw))
    rec = hi.float() + lo.float() * (2.0 ** -11)
    assert (rec - d32).abs().max().item() <= 2.0 ** -22
    a = ops.match_top2(d32, d32[1:], mode=ops.SIM_F16X3, num_pairs=F - 1)
    b = ops.match_top2((hi, lo), (hi[1:], lo[1:]), mode=ops.SIM_F16X3, num_pairs=F - 1)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    idx = torch.tensor([[3, 0], [1, 1]], dtype=torch.int32, device=dev)
    a = ops.match_top2(d32, d32, pair_index=idx, mode=ops.SIM_F16X3)
    b = ops.match_top2((hi, lo), (hi, lo), pair_index=idx, mode=ops.SIM_F16X3)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    with pytest.raises(RuntimeError):
        ops.match_top2((hi, lo), (hi, lo), mode=ops.SIM_TF32X3)


@pytest.mark.parametrize("n,m,d", [c for c in CASES if c[2] % 8 == 0])
def test_bf16_top2(n, m, d, dev):
    from sslam_b200 import ops
    d1, d2, _ = recipes.descriptor_pair(n, m, d, 400 + n + m, noise=3)
    b1 = cu(d1[None], dev).to(torch.bfloat16)
    b2 = cu(d2[None], dev).to(torch.bfloat16)
    top = ops.match_top2(b1, b2, mode=ops.SIM_BF16)
    t = {k: v[0].cpu().numpy() for k, v in top.items()}
    S64 = b1[0].double().cpu().numpy() @ b2[0].double().cpu().numpy().T
    check_top(t, S64, 1e-5, 1e-5)
    _, best12, *_ = oracle.similarity_top2(d1, d2)
    err = np.abs(t["best12"] - best12)
    assert err.max() < 4e-3, err.max()          # bf16 input rounding: 2^-9 per operand, |S| <= 1
    print(f"bf16 {n}x{m}x{d}: max abs err vs fp32 oracle {err.max():.2e}")


def test_tc_pair_index_and_aliasing(dev):
    """Sequence-style aliasing (bank2 = bank1[1:]) and explicit pair lists, both modes."""
    from sslam_b200 import ops
    F, N, D = 5, 200, 64
    bank = np.stack([recipes.descriptor_pair(N, N, D, 500 + f, noise=2)[f % 2] for f in range(F)])
    b32 = cu(bank, dev)
    ref = ops.match_top2(b32, b32[1:], mode=ops.SIM_F32, num_pairs=F - 1)
    for mode, bk in ((ops.SIM_TF32X3, b32), (ops.SIM_F16X3, b32), (ops.SIM_BF16, b32.to(torch.bfloat16))):
        top = ops.match_top2(bk, bk[1:], mode=mode, num_pairs=F - 1)
        tol = 2e-6 if mode != ops.SIM_BF16 else 2e-2
        assert torch.allclose(top["best12"], ref["best12"], atol=tol, rtol=0)
        assert torch.allclose(top["best21"], ref["best21"], atol=tol, rtol=0)
        if mode != ops.SIM_BF16:
            agree = (top["nn12"] == ref["nn12"]).float().mean().item()
            assert agree > 0.999
        idx = torch.tensor([[4, 0], [2, 2], [0, 3]], dtype=torch.int32, device=dev)
        a = ops.match_top2(bk, bk, pair_index=idx, mode=mode)
        r = ops.match_top2(b32, b32, pair_index=idx, mode=ops.SIM_F32)
        assert torch.allclose(a["best12"], r["best12"], atol=tol, rtol=0)


@pytest.mark.parametrize("mode_name", ["tf32x3", "f16x3"])
def test_fp32_mode_sequence_matches_vs_oracle(mode_name, dev):
    """c2-shaped: 3 frames 640x480, K=2048, D=256, consecutive pairs, fp32 modes on tensor cores."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, ops, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 256, 4).to(dev)
    T, K = 3, 2048
    sal, feat = synth.make_sequence(T, seq_id=1)
    fe = FrontEnd(refiner, num_keypoints=K, grid="pixel",
                  sim_mode={"tf32x3": ops.SIM_TF32X3, "f16x3": ops.SIM_F16X3}[mode_name])
    feats, pairs, pscores, counts = fe.run_sequence(sal.to(dev), feat.to(dev), matchers.M1)
    d = feats["descriptors"].cpu().numpy()
    sc = feats["scores"].cpu().numpy()
    p2, q2, c2 = fe.match_consecutive(feats, matchers.M2)
    exc = 0
    for p in range(T - 1):
        S = d[p].astype(np.float64) @ d[p + 1].astype(np.float64).T
        ref = oracle.match_m1(d[p], d[p + 1], 0.8)
        got = pairs[p, :int(counts[p])].cpu().numpy()
        exc += compare_matches(S, np.array([(i, j) for i, j, _ in ref]).reshape(-1, 2), got)
        rm, rq = oracle.match_m2(d[p], d[p + 1], sc[p], sc[p + 1])
        exc += compare_matches(S, rm, p2[p, :int(c2[p])].cpu().numpy(),
                               threshold_margin=lambda i, j: abs(S[i, j] - 0.7))
        assert int(counts[p]) > K // 4
    record(f"fp32_mode_sequence.{mode_name}", {"pairs": T - 1, "near_tie_exceptions": int(exc)})


def test_bf16_sequence_scores(dev):
    """c3-shaped slice: K=4096, bf16 similarity, ratio test 0.8 (M1)."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, ops, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 256, 4).to(dev)
    sal, feat = synth.make_sequence(2, seq_id=2)
    fe = FrontEnd(refiner, num_keypoints=4096, grid="pixel", sim_mode=ops.SIM_BF16)
    feats, pairs, pscores, counts = fe.run_sequence(sal.to(dev), feat.to(dev), matchers.M1)
    d = feats["descriptors"].cpu().numpy()
    ref = oracle.match_m1(d[0], d[1], 0.8)
    refd = {(i, j): s for i, j, s in ref}
    n = int(counts[0])
    got = pairs[0, :n].cpu().numpy()
    gs = pscores[0, :n].cpu().numpy()
    common = [(k, refd[tuple(r)]) for k, r in enumerate(got.tolist()) if tuple(r) in refd]
    agree = len(common) / max(len(ref), 1)
    rel = max(abs(gs[k] - s) / abs(s) for k, s in common)
    record("bf16_sequence_K4096", {"ref_matches": len(ref), "index_agreement": agree, "max_rel_score_err": float(rel)})
    assert agree > 0.98
    assert rel < 1e-2


def test_refiner_fused_vs_oracle_and_torch(dev):
    """DescriptorRefiner on tcgen05 (tf32x3 GEMMs + fused epilogues): descriptors within 1e-5 abs of
    the fp32 oracle (north_star / SURVEY.md §8(d)) and of the cuBLAS-fp32 PyTorch body."""
    from models.descriptor_refiner import DescriptorRefiner
    from parity import load_golden
    torch.manual_seed(0)
    for (C, Hd, D, layers, rows) in ((384, 384, 256, 4, 2048 + 77), (384, 384, 128, 4, 500), (64, 96, 32, 2, 130),
                                     (48, 64, 32, 5, 40)):
        m = DescriptorRefiner(C, Hd, D, layers).to(dev).eval()
        g = torch.Generator().manual_seed(rows)
        x = torch.randn(1, rows, C, generator=g).to(dev)
        with torch.no_grad():
            fused = m(x)
            m.mlp = "torch"
            ref_t = m(x)
            m.mlp = "tcgen05"
        w = oracle.RefinerWeights.from_state_dict(m.state_dict())
        ref_o = oracle.refiner_forward(w, x.cpu().numpy())
        err_o = np.abs(fused.cpu().numpy() - ref_o).max()
        err_t = (fused - ref_t).abs().max().item()
        print(f"refiner {C}-{Hd}-{D} x{layers} rows={rows}: max|fused-oracle| {err_o:.2e}  max|fused-torch| {err_t:.2e}")
        assert err_o < 1e-5 and err_t < 1e-5
        assert torch.allclose(fused.norm(dim=-1), torch.ones_like(fused[..., 0]), atol=1e-5)
    # golden weights written by the reference's own module
    z, _ = load_golden("refiner")
    m = DescriptorRefiner(input_dim=48, hidden_dim=64, output_dim=32, num_layers=4)
    m.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")})
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m(cu(z["x"], dev)).cpu().numpy()
    assert np.abs(out - z["out"]).max() < 1e-5


def test_refiner_layer_fused_equals_per_layer(dev):
    """The layer-fused persistent kernel (refiner_fused_kernel: activations exchanged through L2, weights
    switched per layer) performs the arithmetic of the per-layer GEMM launches in the same order: outputs are
    bit-identical for every chunk depth S = 1..4, ragged row counts, several chunks per cluster, 0..3
    residual blocks and partial column tiles."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(1)
    try:
        for (C, Hd, D, layers, rows) in ((384, 384, 256, 4, 40), (384, 384, 256, 4, 2048 + 77), (384, 384, 256, 4, 30000),
                                         (384, 384, 128, 4, 6000), (64, 96, 32, 2, 130), (48, 64, 32, 5, 9000),
                                         (384, 384, 256, 5, 23000), (320, 264, 200, 3, 12345), (384, 384, 256, 4, 200000)):
            m = DescriptorRefiner(C, Hd, D, layers).to(dev).eval()
            with torch.no_grad():
                for blk in m.residual_blocks:                 # non-trivial LayerNorm affine parameters
                    for ln in (blk.norm1, blk.norm2):
                        ln.weight.uniform_(0.5, 1.5)
                        ln.bias.uniform_(-0.3, 0.3)
            g = torch.Generator().manual_seed(rows)
            x = torch.randn(1, rows, C, generator=g).to(dev)
            lib.sslam_debug_refiner_fused(0)
            with torch.no_grad():
                ref = m(x).clone()
            for S in (1, 2, 3, 4):
                lib.sslam_debug_refiner_fused(S)
                with torch.no_grad():
                    out = m(x)
                torch.cuda.synchronize()
                assert torch.equal(out, ref), (C, Hd, D, layers, rows, S, float((out - ref).abs().max()))
    finally:
        lib.sslam_debug_refiner_fused(3)


def test_refiner_layer_fused_back_to_back(dev):
    """200 launches of the fused kernel in a row on c2-sized chunks give the same bits every time (the
    cluster protocol — multicast stages, weight switch, ready barriers — leaves no state behind)."""
    from models.descriptor_refiner import DescriptorRefiner
    torch.manual_seed(2)
    m = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
    x = torch.randn(1, 50 * 2048, 384, device=dev)
    with torch.no_grad():
        ref = m(x).clone()
        for _ in range(200):
            out = m(x)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)


def test_gather_pair_output(dev):
    """gather_bilinear(pair=True): fp16 (hi, lo) operands reproduce the fp32 sample to 22 bits."""
    from sslam_b200 import ops
    feat = cu(recipes.int_features(2, 30, 40, 384, 71), dev)
    kp = cu(recipes.pixel_keypoints(2, 300, 480, 640, 72), dev)
    ref = ops.gather_bilinear(feat, kp, pixel_coords=True)
    hi, lo = ops.gather_bilinear(feat, kp, pixel_coords=True, pair=True)
    assert hi.dtype == torch.float16 and lo.dtype == torch.float16 and hi.shape == ref.shape
    rec = hi.float() + lo.float() / 2048.0
    err = (rec - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -21 + 1e-7).all())


def test_profile_api(dev):
    from sslam_b200 import ops
    ops.profile_enable(True)
    x = torch.rand(3, 64, 96, device=dev)
    ops.decode_topk(x, 20)
    ops.l2norm_rows(torch.rand(10, 32, device=dev))
    torch.cuda.synchronize()
    prof = ops.profile_read()
    ops.profile_enable(False)
    assert prof["decode_scan"][1] == 1 and prof["decode_topk"][1] == 1 and prof["l2norm"][1] == 1
    assert all(ms >= 0 for ms, _ in prof.values())
    assert ops.profile_read() == {}


def test_c3_bf16_pairs_with_ratio_rules(dev):
    """c3-shaped: independent pairs, K=4096, bf16 similarity; M1 (ratio 0.8) and the rejecting M3
    rule (0.9) against the fp32 oracle on the same descriptors: index agreement reported, scores
    within 1e-3 relative."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, ops, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 256, 4).to(dev)
    sal, feat = synth.make_pairs(2)                       # (P, 2, ...)
    fe = FrontEnd(refiner, num_keypoints=4096, grid="pixel", sim_mode=ops.SIM_BF16)
    P = sal.shape[0]
    f = fe.extract(sal.reshape(-1, *sal.shape[2:]).to(dev), feat.reshape(-1, *feat.shape[2:]).to(dev))
    idx = torch.tensor([[2 * p, 2 * p + 1] for p in range(P)], dtype=torch.int32, device=dev)
    pairs, sc, cnt = fe.match_pairs(f, idx, matchers.M3, ratio_threshold=0.9)
    d = f["descriptors"].cpu().numpy()
    for p in range(P):
        ref, dist = oracle.match_m3(d[2 * p], d[2 * p + 1], 0.9)
        got = pairs[p, :int(cnt[p])].cpu().numpy()
        refset, gotset = {tuple(r) for r in ref.tolist()}, {tuple(r) for r in got.tolist()}
        agree = len(refset & gotset) / max(len(refset | gotset), 1)
        refd = {tuple(r): v for r, v in zip(ref.tolist(), dist.tolist())}
        rel = max(abs((1 - sc[p, k].item()) - (1 - refd[tuple(r)])) / abs(1 - refd[tuple(r)])
                  for k, r in enumerate(got.tolist()) if tuple(r) in refd)
        record(f"c3_bf16_M3.pair{p}", {"ref_matches": len(refset), "index_agreement": agree, "max_rel_score_err": float(rel)})
        assert agree > 0.97 and rel < 1e-2


def test_c4_all_pairs_keyframes(dev):
    """c4-shaped: all unordered pairs of a keyframe set matched from one resident bank via
    pair_index, matcher M2, dealt to 2 ranks block-cyclically; union equals the full list."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import dist as sdist, matchers, ops, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 256, 4).to(dev)
    KF, K = 6, 512
    sal, feat = synth.make_sequence(KF, seq_id=5, stride=8)
    fe = FrontEnd(refiner, num_keypoints=K, grid="pixel", sim_mode=ops.SIM_F16X3)
    f = fe.extract(sal.to(dev), feat.to(dev))
    idx = sdist.all_pairs_index(KF)
    assert idx.shape[0] == KF * (KF - 1) // 2
    full = fe.match_pairs(f, idx.to(dev), matchers.M2)
    d, s = f["descriptors"].cpu().numpy(), f["scores"].cpu().numpy()
    exc = 0
    for p, (a, b) in enumerate(idx.tolist()):
        S = d[a].astype(np.float64) @ d[b].astype(np.float64).T
        rm, rq = oracle.match_m2(d[a], d[b], s[a], s[b])
        exc += compare_matches(S, rm, full[0][p, :int(full[2][p])].cpu().numpy(),
                               threshold_margin=lambda i, j: abs(S[i, j] - 0.7))
    owned = [sdist.deal_pairs_block_cyclic(idx, 2, r, block=2) for r in range(2)]
    assert torch.equal(torch.cat(owned).sort().values, torch.arange(idx.shape[0]))
    for r in range(2):
        part = fe.match_pairs(f, idx[owned[r]].to(dev), matchers.M2)
        assert torch.equal(part[2], full[2][owned[r].to(dev)])
        assert torch.equal(part[0], full[0][owned[r].to(dev)])
    record("c4_all_pairs_keyframes", {"pairs": int(idx.shape[0]), "near_tie_exceptions": int(exc)})


def test_c5_hires_pair(dev):
    """c5-shaped: 1280x960 frames, K=8192, D=256, fp32 (f16x3) mode, one consecutive pair."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import matchers, ops, synth
    from sslam_b200.pipeline import FrontEnd
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 256, 4).to(dev)
    sal, feat = synth.make_sequence(2, seq_id=6, height=960, width=1280)
    K = 8192
    fe = FrontEnd(refiner, num_keypoints=K, grid="pixel", sim_mode=ops.SIM_F16X3)
    feats, pairs, pscores, counts = fe.run_sequence(sal.to(dev), feat.to(dev), matchers.M1)
    okp, osc, oinfo = oracle.select_keypoints(sal.numpy(), K)
    assert np.array_equal(feats["keypoints_pixel"].cpu().numpy(), okp)
    d = feats["descriptors"].cpu().numpy()
    S = d[0].astype(np.float64) @ d[1].astype(np.float64).T
    ref = oracle.match_m1(d[0], d[1], 0.8)
    exc = compare_matches(S, np.array([(i, j) for i, j, _ in ref]).reshape(-1, 2),
                          pairs[0, :int(counts[0])].cpu().numpy())
    assert int(counts[0]) > K // 4
    record("c5_hires_pair", {"matches": int(counts[0]), "near_tie_exceptions": int(exc)})


def _refiner_fp64(m, x):
    """DescriptorRefiner forward in float64 (models/descriptor_refiner.py:73-86, 108-126) from the module's own weights."""
    sd = {k: v.detach().cpu().double() for k, v in m.state_dict().items()}
    h = torch.relu(x.cpu().double().reshape(-1, x.shape[-1]) @ sd["input_proj.weight"].T + sd["input_proj.bias"])
    i = 0
    while f"residual_blocks.{i}.fc1.weight" in sd:
        pre = f"residual_blocks.{i}."
        def ln(t, w, b):
            mu = t.mean(-1, keepdim=True)
            var = ((t - mu) ** 2).mean(-1, keepdim=True)
            return (t - mu) / torch.sqrt(var + 1e-5) * w + b
        u = torch.relu(ln(h, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"]) @ sd[pre + "fc1.weight"].T + sd[pre + "fc1.bias"])
        u = ln(u, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"]) @ sd[pre + "fc2.weight"].T + sd[pre + "fc2.bias"]
        h = torch.relu(u + h)
        i += 1
    d = h @ sd["output_proj.weight"].T + sd["output_proj.bias"]
    return d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)


def test_refiner_trained_like_weights_and_range_guard(dev):
    """Weights and inputs unlike a fresh initialisation: LayerNorm gains 0.3..2.5 and shifts
    +-1, biases that put the row mean several standard deviations from zero (the one-pass variance
    E[x^2]-mu^2 of the folded LayerNorm cancels there), scaled weights, inputs with an offset and with
    outlier channels of magnitude ~400 as DINO features have.  Descriptors stay within 1e-5 of a float64
    evaluation.  And the range guard: inputs that push an activation past the fp16 range make forward()
    raise SSLAM_ERANGE instead of returning NaN, after which the flag is clear again."""
    from models.descriptor_refiner import DescriptorRefiner
    from sslam_b200 import ops, _lib
    torch.manual_seed(7)
    m = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
    with torch.no_grad():
        m.input_proj.weight.mul_(1.7)
        m.input_proj.bias.add_(2.0)
        for blk in m.residual_blocks:
            for lnm in (blk.norm1, blk.norm2):
                lnm.weight.uniform_(0.3, 2.5)
                lnm.bias.uniform_(-1.0, 1.0)
            blk.fc1.weight.mul_(1.5); blk.fc1.bias.add_(1.0)
            blk.fc2.weight.mul_(0.7); blk.fc2.bias.add_(3.0)
        m.output_proj.weight.mul_(2.0)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 3000, 384, generator=g) * 4.0 + 1.5
    x[..., 7] *= 50.0                                       # outlier channels
    x[..., 200] = x[..., 200] * 30.0 + 100.0
    x = x.to(dev)
    with torch.no_grad():
        out = m(x)[0].double().cpu()
    ref = _refiner_fp64(m, x)
    err = float((out - ref).abs().max())
    # row statistics of the LayerNorm inputs, for the record: how far the mean is from zero in units of std
    with torch.no_grad():
        h = torch.relu(m.input_proj(x[0]))
    ratio = float((h.mean(-1).abs() / h.std(-1)).max())
    record("refiner_trained_like", {"max_abs_err_vs_fp64": err, "max_row_mean_over_std": ratio})
    print(f"trained-like refiner: max|gpu - fp64| = {err:.2e}, max |row mean| / std of the first LayerNorm input = {ratio:.1f}")
    assert err < 1e-5
    ops.refiner_range_check()                               # nothing left the range
    # rows whose mean is tens of standard deviations from zero (bias shift 40 on every LayerNorm input)
    m2 = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
    with torch.no_grad():
        m2.input_proj.bias.add_(40.0)
        for blk in m2.residual_blocks:
            blk.fc2.bias.add_(40.0)
    xs = torch.randn(1, 2000, 384, generator=torch.Generator().manual_seed(3)).to(dev)
    with torch.no_grad():
        o2 = m2(xs)[0].double().cpu()
        h2 = torch.relu(m2.input_proj(xs[0]))
    err2 = float((o2 - _refiner_fp64(m2, xs)).abs().max())
    ratio2 = float((h2.mean(-1).abs() / h2.std(-1)).max())
    record("refiner_shifted_rows", {"max_abs_err_vs_fp64": err2, "max_row_mean_over_std": ratio2})
    print(f"shifted rows: max|gpu - fp64| = {err2:.2e} at |row mean| / std up to {ratio2:.0f}")
    assert err2 < 1e-5 and ratio2 > 30
    # out of range: hidden activations ~ 1e5
    with torch.no_grad():
        try:
            m(x * 3.0e3)
            raised = False
        except _lib.SslamError as e:
            raised = e.code == -6
    assert raised, "forward() must raise SSLAM_ERANGE when an activation leaves the fp16 range"
    with torch.no_grad():
        out2 = m(x)[0].double().cpu()                       # flag cleared, results unchanged
    assert torch.equal(out2, out)
