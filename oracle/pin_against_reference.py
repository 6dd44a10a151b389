#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions (build container only).

Usage (from the repo root, in the container that mounts /root/reference):

    python oracle/pin_against_reference.py [--reference /root/reference] [--timing]

The reference has no golden vectors of its own for this path, so its own
functions, imported — never copied — from ``<reference>/semantic-slam``, are executed on inputs
made by ``oracle/recipes.py`` and the results are stored.  ``timm`` and ``matplotlib`` are absent
from the image; empty stub modules satisfy the imports (the ViT and the plotting code are never
called).  Library versions are recorded in every fixture.

This script is test infrastructure: it is the only file that reads /root/reference, and nothing
at test/bench/smoke time on the GPU box depends on it — only on the fixtures it wrote.
"""

import argparse
import json
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import recipes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_reference(ref_root):
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    os.environ.setdefault("WANDB_MODE", "disabled")
    sys.dont_write_bytecode = True
    for name in ("timm", "matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.patches"].Circle = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    pkg = os.path.join(ref_root, "semantic-slam")
    sys.path.insert(0, pkg)
    sys.path.insert(0, os.path.join(pkg, "test"))
    import torch  # noqa: F401
    from models.keypoint_selector import KeypointSelector
    from models.dino_backbone import DinoBackbone
    from models.descriptor_refiner import DescriptorRefiner
    from visualize_matches import MatchVisualizer
    from visualize_matches_sequence import SequenceMatcher
    from train import SemanticSLAMTrainer
    from test_descriptor_quality import DescriptorQualityTester
    from test_repeatability import RepeatabilityTester
    return dict(RepeatabilityTester=RepeatabilityTester, KeypointSelector=KeypointSelector, DinoBackbone=DinoBackbone,
                DescriptorRefiner=DescriptorRefiner, MatchVisualizer=MatchVisualizer,
                SequenceMatcher=SequenceMatcher, SemanticSLAMTrainer=SemanticSLAMTrainer,
                DescriptorQualityTester=DescriptorQualityTester)


def versions():
    import torch
    return json.dumps({"torch": torch.__version__, "numpy": np.__version__})


# --------------------------------------------------------------------------- decode
# name -> (recipe, kwargs, K, nms_radius, percentile, store_input)
DECODE_CASES = {
    # reference-native grid, always branch B (duplicates, unsorted first group)
    "grid28_k500": ("spread", dict(H=28, W=28, seed=11), 500, 2, 0.5, True),
    # branch B ending in raw padding: 60x80 grid, K=2048
    "grid60x80_k2048": ("spread", dict(H=60, W=80, seed=12), 2048, 2, 0.5, True),
    # branch B satisfied by a lower percentile
    "grid48x64_k80_lower": ("box", dict(H=48, W=64, seed=13), 80, 2, 0.5, True),
    # main branch, mid size
    "map96x128_k128": ("box", dict(H=96, W=128, seed=14), 128, 2, 0.5, True),
    # quantised map: plateaus + boundary ties
    "map96x128_q64_k200": ("boxq", dict(H=96, W=128, seed=15, quant=64), 200, 2, 0.5, True),
    # other radii / percentiles
    "map64x64_r0_k300": ("box", dict(H=64, W=64, seed=16), 300, 0, 0.5, True),
    "map64x64_r1_p70_k100": ("box", dict(H=64, W=64, seed=17), 100, 1, 0.7, True),
    "map64x64_r3_k40": ("box", dict(H=64, W=64, seed=18), 40, 3, 0.5, True),
    # odd sizes (no alignment anywhere)
    "map37x53_k60": ("spread", dict(H=37, W=53, seed=19), 60, 2, 0.5, True),
    # full size, inputs regenerated from the recipe (hash checked)
    "tum480x640_k1024": ("spread", dict(H=480, W=640, seed=21), 1024, 2, 0.5, False),
    "tum480x640_k2048": ("spread", dict(H=480, W=640, seed=22), 2048, 2, 0.5, False),
    "tum480x640_k4096": ("spread", dict(H=480, W=640, seed=23), 4096, 2, 0.5, False),
    "hires960x1280_k8192": ("spread", dict(H=960, W=1280, seed=24), 8192, 2, 0.5, False),
}


def decode_input(kind, kw):
    if kind == "spread":
        return recipes.spread_saliency(**kw)
    if kind == "box":
        return recipes.box_saliency(**kw)
    if kind == "boxq":
        return recipes.box_saliency(**kw)
    raise KeyError(kind)


def special_decode_inputs():
    """Degenerate maps: branch C, three maxima, all-constant high."""
    out = {}
    out["const005_k50"] = (np.full((28, 28), 0.05, dtype=np.float32), 50)          # n = 0
    three = np.full((28, 28), 0.2, dtype=np.float32)
    three[3, 4], three[10, 20], three[25, 7] = 0.9, 0.8, 0.7
    out["three_maxima_k20"] = (three, 20)
    out["const09_k30"] = (np.full((20, 24), 0.9, dtype=np.float32), 30)            # nothing > median
    ramp = (np.arange(32 * 32, dtype=np.float32).reshape(32, 32) / np.float32(1024)).astype(np.float32)
    out["ramp_k64"] = (ramp, 64)
    # branch B satisfied by a lower percentile (keypoint_selector.py:139-156): isolated peaks on a
    # dim background; 20 peaks above the 0.1 floor, 36 more in (0.05, 0.1]
    rng = np.random.Generator(np.random.PCG64(77))
    low = (rng.integers(10, 41, size=(40, 48)) / 1000.0).astype(np.float32)
    k = 0
    for yy in range(2, 40, 6):
        for xx in range(2, 48, 6):
            low[yy, xx] = np.float32((150 + 10 * k) / 1000.0) if k < 20 else np.float32((52 + k) / 1000.0)
            k += 1
    out["peaks_lower_k40"] = (low, 40)
    return out


def gen_decode(ref):
    import torch
    sel = ref["KeypointSelector"](384, 256)
    blob = {"versions": versions()}
    meta = {}
    for name, (kind, kw, K, r, p, store) in DECODE_CASES.items():
        sal = decode_input(kind, kw)
        t = torch.from_numpy(sal)[None, :, :, None]
        kp, sc = sel.select_keypoints(t, num_keypoints=K, nms_radius=r, min_score_percentile=p)
        blob[name + ".kpts"] = kp[0].numpy()
        blob[name + ".scores"] = sc[0].numpy()
        if store:
            blob[name + ".sal"] = sal
        meta[name] = dict(kind=kind, kw=kw, K=K, nms_radius=r, pct=p, stored=store,
                          sha256=recipes.sha256(sal))
    for name, (sal, K) in special_decode_inputs().items():
        t = torch.from_numpy(sal)[None, :, :, None]
        kp, sc = sel.select_keypoints(t, num_keypoints=K)
        blob[name + ".kpts"] = kp[0].numpy()
        blob[name + ".scores"] = sc[0].numpy()
        blob[name + ".sal"] = sal
        meta[name] = dict(kind="stored", kw={}, K=K, nms_radius=2, pct=0.5, stored=True,
                          sha256=recipes.sha256(sal))
    # error behaviour: k > H*W in a fallback raises (keypoint_selector.py:166)
    try:
        sel.select_keypoints(torch.from_numpy(recipes.spread_saliency(30, 40, 5))[None, :, :, None],
                             num_keypoints=2048)
        meta["raises_k_gt_hw"] = False
    except RuntimeError as e:
        meta["raises_k_gt_hw"] = str(e)
    # NMS alone
    nms_in = recipes.box_saliency(40, 56, 31, quant=32)
    for r in (0, 1, 2, 3):
        blob[f"nms.r{r}"] = sel._apply_nms(torch.from_numpy(nms_in)[None], r)[0].numpy()
    blob["nms.in"] = nms_in
    # head + sigmoid on the native grid (c0): saliency from seeded weights
    torch.manual_seed(0)
    sel0 = ref["KeypointSelector"](384, 256)
    g = torch.Generator().manual_seed(7)
    feats = torch.randn(2, 28, 28, 384, generator=g)
    with torch.no_grad():
        sal0 = sel0(feats)
        kp0, sc0 = sel0.select_keypoints(sal0, num_keypoints=500)
    blob["c0.sal"] = sal0[..., 0].numpy()
    blob["c0.kpts"] = kp0.numpy()
    blob["c0.scores"] = sc0.numpy()
    blob["meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(GOLDEN, "decode.npz"), **blob)
    print("decode.npz:", len(meta), "cases")


def gen_quantile():
    import torch
    rng = np.random.Generator(np.random.PCG64(99))
    blob = {"versions": versions()}
    cases = []
    i = 0
    for n in (2, 3, 17, 784, 1200, 4799, 4800):
        for q in (0.5, 0.4, 0.3, 0.2, 0.1, 0.37, 0.0, 1.0):
            v = (rng.integers(0, 1 << 24, size=n).astype(np.float64) / float(1 << 24)).astype(np.float32)
            out = torch.quantile(torch.from_numpy(v), q).numpy()
            blob[f"v{i}"] = v
            blob[f"o{i}"] = out
            cases.append((i, n, q))
            i += 1
    blob["cases"] = json.dumps(cases)
    np.savez_compressed(os.path.join(GOLDEN, "quantile.npz"), **blob)
    print("quantile.npz:", len(cases), "cases")


# --------------------------------------------------------------------------- gather / refiner
def gen_gather(ref):
    import torch
    from types import SimpleNamespace
    DB = ref["DinoBackbone"]
    ns = SimpleNamespace(patch_size=16)
    blob = {"versions": versions()}
    # small map, arbitrary float coordinates incl. out-of-range (zero padding taps)
    feat = recipes.int_features(2, 7, 9, 32, 41)
    rng = np.random.Generator(np.random.PCG64(42))
    kp = np.stack([rng.integers(-64, 9 * 64 + 64, size=(2, 96)) / 64.0,
                   rng.integers(-64, 7 * 64 + 64, size=(2, 96)) / 64.0], -1).astype(np.float32)
    blob["small.feat"], blob["small.kpts"] = feat, kp
    blob["small.out"] = DB.extract_at_keypoints(None, torch.from_numpy(feat), torch.from_numpy(kp)).numpy()
    # pipeline-P shape: 30x40x384 map, integer pixel keypoints through pixel_to_patch
    feat = recipes.int_features(1, 30, 40, 384, 43)
    pix = recipes.pixel_keypoints(1, 160, 480, 640, 44)
    pc = DB.pixel_to_patch(ns, torch.from_numpy(pix))
    blob["tum.pix"] = pix
    blob["tum.patch"] = pc.numpy()
    blob["tum.back"] = DB.patch_to_pixel(ns, pc).numpy()
    blob["tum.out"] = DB.extract_at_keypoints(None, torch.from_numpy(feat), pc).numpy()
    blob["tum.feat_sha256"] = recipes.sha256(feat)
    # native grid: integer patch coordinates (not an exact gather)
    feat = recipes.int_features(1, 28, 28, 64, 45)
    kpi = recipes.pixel_keypoints(1, 200, 28, 28, 46)
    blob["native.feat"], blob["native.kpts"] = feat, kpi
    blob["native.out"] = DB.extract_at_keypoints(None, torch.from_numpy(feat), torch.from_numpy(kpi)).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "gather.npz"), **blob)
    print("gather.npz written")


def gen_refiner(ref):
    import torch
    blob = {"versions": versions()}
    torch.manual_seed(0)
    m = ref["DescriptorRefiner"](input_dim=48, hidden_dim=64, output_dim=32, num_layers=4).eval()
    for k, v in m.state_dict().items():
        blob["w." + k] = v.numpy()
    x = recipes.int_features(2, 5, 8, 48, 51).reshape(2, 40, 48)
    with torch.no_grad():
        blob["out"] = m(torch.from_numpy(x)).numpy()
    blob["x"] = x
    # F.normalize alone, incl. an all-zero row (eps clamp)
    z = recipes.int_features(1, 1, 12, 256, 52).reshape(12, 256)
    z[3] = 0
    blob["norm.in"] = z
    blob["norm.out"] = torch.nn.functional.normalize(torch.from_numpy(z), p=2, dim=-1).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "refiner.npz"), **blob)
    print("refiner.npz written")


# --------------------------------------------------------------------------- matchers
MATCH_CASES = {
    # name: (n, m, d, seed, noise, dup_every, stored[, near_dup_every])
    "small": (192, 160, 64, 61, 1, 0, True),
    "ties": (96, 96, 32, 62, 0, 7, True),
    "wide": (64, 300, 48, 63, 2, 0, True),
    "single_col": (8, 1, 16, 64, 1, 0, True),
    "hard": (256, 256, 64, 67, 6, 0, True, 5),
    "hard_k2048": (2048, 2048, 256, 68, 10, 0, False, 3),
    "k1024": (1024, 1024, 256, 65, 3, 0, False),
    "k2048": (2048, 2048, 256, 66, 3, 0, False),
}


def gen_match(ref):
    import torch
    MV, SM = ref["MatchVisualizer"], ref["SequenceMatcher"]
    TR, DQ = ref["SemanticSLAMTrainer"], ref["DescriptorQualityTester"]
    blob = {"versions": versions()}
    meta = {}
    for name, spec in MATCH_CASES.items():
        n, m, d, seed, noise, dup, stored = spec[:7]
        near = spec[7] if len(spec) > 7 else 0
        d1, d2, _ = recipes.descriptor_pair(n, m, d, seed, noise=noise, dup_every=dup,
                                            near_dup_every=near)
        rng = np.random.Generator(np.random.PCG64(seed + 1000))
        s1 = (rng.integers(0, 1024, size=n) / 1024.0).astype(np.float32)
        s2 = (rng.integers(0, 1024, size=m) / 1024.0).astype(np.float32)
        i1 = (rng.integers(0, 256, size=n) / 255.0).astype(np.float32)
        i2 = (rng.integers(0, 256, size=m) / 255.0).astype(np.float32)
        if stored:
            blob[name + ".d1"], blob[name + ".d2"] = d1, d2
        blob[name + ".s1"], blob[name + ".s2"] = s1, s2
        blob[name + ".i1"], blob[name + ".i2"] = i1, i2
        m1 = MV.find_matches(None, d1, d2, 0.8)
        blob[name + ".m1"] = np.array([(i, j) for i, j, _ in m1], dtype=np.int64).reshape(-1, 2)
        blob[name + ".m1s"] = np.array([s for _, _, s in m1], dtype=np.float32)
        m1b = MV.find_matches(None, d1, d2, 1.02)            # a ratio that can reject
        blob[name + ".m1b"] = np.array([(i, j) for i, j, _ in m1b], dtype=np.int64).reshape(-1, 2)
        for tag, kw in (("m2", {}), ("m2i", dict(intensity1=i1, intensity2=i2, min_intensity=0.15,
                                                 min_saliency=0.5)),
                        ("m2none", dict(min_descriptor_sim=1.5))):
            mm, qq = SM.match_with_quality(d1, d2, s1, s2, **kw)
            blob[f"{name}.{tag}"], blob[f"{name}.{tag}q"] = mm, qq
        if m >= 2:
            m3, dist = DQ.find_mutual_nearest_neighbors(None, d1, d2, 0.9)
            blob[name + ".m3"], blob[name + ".m3d"] = m3.astype(np.int64).reshape(-1, 2), dist
            m3b, _ = DQ.find_mutual_nearest_neighbors(None, d1, d2, 0.98)
            blob[name + ".m3b"] = m3b.astype(np.int64).reshape(-1, 2)
        if n == m:
            b2 = np.stack([d1, d2[::-1].copy()]), np.stack([d2, d1])
            blob[name + ".m4"] = TR._find_matches(None, torch.from_numpy(b2[0]), torch.from_numpy(b2[1])).numpy()
        sim = d1 @ d2.T                                       # test/test_tracking.py:159-161
        blob[name + ".m5"] = np.int64((sim.max(axis=1) > 0.8).sum())
        blob[name + ".m5lo"] = np.int64((sim.max(axis=1) > 0.5).sum())
        meta[name] = dict(n=n, m=m, d=d, seed=seed, noise=noise, dup_every=dup, stored=stored,
                          near_dup_every=near, sha256=[recipes.sha256(d1), recipes.sha256(d2)])
    # all-empty M4 (train.py:440)
    z = np.zeros((2, 4, 8), dtype=np.float32)
    z[:, :, 0] = 1.0
    e1 = z.copy()
    e2 = z.copy()
    e2[:, :, 0] = -1.0
    blob["m4.degenerate"] = TR._find_matches(None, torch.from_numpy(e1), torch.from_numpy(e2)).numpy()
    blob["meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(GOLDEN, "match.npz"), **blob)
    print("match.npz:", len(meta), "cases")


# --------------------------------------------------------------------------- evaluation adaptors
def eval_case(n, m, seed, H_kind):
    """Integer-valued pixel keypoints (as the decode emits) of two frames related by a homography."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k1 = np.stack([rng.integers(0, 640, size=n), rng.integers(0, 480, size=n)], 1).astype(np.float32)
    if H_kind == "shift":
        H = np.array([[1.0, 0.0, 16.0], [0.0, 1.0, -3.0], [0.0, 0.0, 1.0]])
    elif H_kind == "persp":
        H = np.array([[1.02, 0.015, 5.5], [-0.01, 0.98, 2.25], [1e-5, -2e-5, 1.0]])
    else:
        H = np.eye(3)
    homo = np.concatenate([k1.astype(np.float64), np.ones((n, 1))], 1)
    w = (H @ homo.T).T
    w = w[:, :2] / w[:, 2:3]
    # frame-2 keypoints: a shuffled subset of the warped points rounded to pixels (+ jitter), plus fresh ones
    keep = rng.permutation(n)[: min(n, m) * 2 // 3]
    k2 = np.round(w[keep] + rng.integers(-2, 3, size=(keep.size, 2))).astype(np.float32)
    extra = np.stack([rng.integers(0, 640, size=m - keep.size), rng.integers(0, 480, size=m - keep.size)], 1)
    k2 = np.concatenate([k2, extra.astype(np.float32)], 0)[rng.permutation(m)]
    return k1, np.ascontiguousarray(k2), H


EVAL_CASES = {"small_shift": (200, 180, 71, "shift"), "persp": (500, 640, 72, "persp"),
              "k2048": (2048, 2048, 73, "persp"), "ident": (64, 64, 74, "ident")}


def gen_evaluation(ref):
    DQ, RT = ref["DescriptorQualityTester"], ref["RepeatabilityTester"]
    blob = {"versions": versions()}
    meta = {}
    for name, (n, m, seed, kind) in EVAL_CASES.items():
        k1, k2, H = eval_case(n, m, seed, kind)
        blob[name + ".k1"], blob[name + ".k2"], blob[name + ".H"] = k1, k2, H
        for thr in (3.0, 1.0):
            gt = DQ.compute_ground_truth_matches(None, k1, k2, H, thr)
            blob[f"{name}.gt{thr:g}"] = gt.astype(np.int64).reshape(-1, 2)
        gt = blob[name + ".gt3"]
        # a "prediction": half of the gt pairs, plus wrong pairs on other rows
        rng = np.random.Generator(np.random.PCG64(seed + 5))
        pred = gt[::2].copy()
        free = np.setdiff1d(np.arange(n), gt[:, 0])[:40]
        wrong = np.stack([free, rng.integers(0, m, size=free.size)], 1)
        pred = np.concatenate([pred, wrong], 0).astype(np.int64)
        blob[name + ".pred"] = pred
        ev = DQ.evaluate_matches(None, pred, gt, n, m)
        blob[name + ".eval"] = np.array([ev["tp"], ev["fp"], ev["fn"]], dtype=np.int64)
        blob[name + ".evalf"] = np.array([ev["precision"], ev["recall"], ev["f1"], ev["inlier_ratio"]])
        for tag, HH in (("repH", H), ("rep0", None)):
            r = RT.compute_repeatability(None, k1, k2, HH, 3.0)
            blob[f"{name}.{tag}"] = np.array([r["repeatability"], float(r["repeatable_count"]),
                                              float(r["mean_nn_distance"]), float(r["median_nn_distance"])])
        meta[name] = dict(n=n, m=m, seed=seed, H=kind)
    blob["meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(GOLDEN, "evaluation.npz"), **blob)
    print("evaluation.npz:", len(meta), "cases")


# --------------------------------------------------------------------------- reference CPU timing
def gen_timing(ref, frames=6):
    """Times the unmodified reference functions on c1/c2-shaped inputs in THIS container
    (method of test/test_performance.py:88-131: warm-ups, perf_counter).  Orientation only —
    the GPU box's host is timed by bench.py with the oracle port."""
    import torch
    from types import SimpleNamespace
    sys.path.insert(0, os.path.join(ROOT, "semantic-slam-master_b200"))
    from sslam_b200 import synth
    sel = ref["KeypointSelector"](384, 256)
    DB, ns = ref["DinoBackbone"], SimpleNamespace(patch_size=16)
    res = {"versions": json.loads(versions()), "threads": torch.get_num_threads(),
           "cpu_count": os.cpu_count(), "frames": frames}
    for K in (1024, 2048):
        torch.manual_seed(0)
        refiner = ref["DescriptorRefiner"](384, 384, 256, 4).eval()
        sal, feat = synth.make_sequence(frames, seq_id=0)
        t = dict(select=[], gather=[], refiner=[], m1=[], m2=[])
        outs = []
        with torch.no_grad():
            for it in range(frames):
                t0 = time.perf_counter()
                kp, sc = sel.select_keypoints(sal[it:it + 1], num_keypoints=K)
                t1 = time.perf_counter()
                f = DB.extract_at_keypoints(None, feat[it:it + 1], DB.pixel_to_patch(ns, kp))
                t2 = time.perf_counter()
                d = refiner(f)
                t3 = time.perf_counter()
                outs.append((d[0].numpy(), sc[0].numpy()))
                if it >= 1:
                    t["select"].append(t1 - t0); t["gather"].append(t2 - t1); t["refiner"].append(t3 - t2)
            for it in range(1, frames - 1):
                a, b = outs[it], outs[it + 1]
                t0 = time.perf_counter()
                ref["MatchVisualizer"].find_matches(None, a[0], b[0], 0.8)
                t1 = time.perf_counter()
                ref["SequenceMatcher"].match_with_quality(a[0], b[0], a[1], b[1])
                t2 = time.perf_counter()
                t["m1"].append(t1 - t0); t["m2"].append(t2 - t1)
        ms = {k: 1e3 * float(np.median(v)) for k, v in t.items()}
        ms["pairs_per_s_m1"] = 1e3 / (ms["select"] + ms["gather"] + ms["refiner"] + ms["m1"])
        res[f"K{K}"] = ms
    with open(os.path.join(GOLDEN, "reference_timing.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("SSLAM_REFERENCE", "/root/reference"))
    ap.add_argument("--timing", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    ref = load_reference(args.reference)
    todo = args.only.split(",") if args.only else ["quantile", "decode", "gather", "refiner", "match", "evaluation"]
    if "quantile" in todo:
        gen_quantile()
    if "decode" in todo:
        gen_decode(ref)
    if "gather" in todo:
        gen_gather(ref)
    if "refiner" in todo:
        gen_refiner(ref)
    if "match" in todo:
        gen_match(ref)
    if "evaluation" in todo:
        gen_evaluation(ref)
    if args.timing:
        gen_timing(ref)


if __name__ == "__main__":
    main()
