"""Matcher variants M1..M5 on top of the fused top-2 primitive.

Device API: ``match(...)`` keeps everything on the GPU (padded pair lists + counts).
Host API: ``find_matches`` / ``match_with_quality`` / ``find_mutual_nearest_neighbors`` /
``find_matches_batched`` / ``tracking_count`` take and return what the reference functions take
and return (NumPy arrays / lists), moving data over PCIe around one kernel sequence.
"""

import numpy as np
import torch

from . import ops
from .ops import M1, M2, M3, M4, M5, SIM_F32, SIM_BF16, SIM_TF32X3, SIM_F16X3  # noqa: F401


def _f32(x):
    return float(np.float32(x))


def variant_params(variant, **kw):
    """fp32-rounded scalar parameters, as the reference's fp32 comparisons see them."""
    if variant == M1:
        # "nep50": `sim > second * ratio` in fp32 (NumPy >= 2 casts the Python float to fp32);
        # "legacy": in double, as NumPy < 2 (pinned by the reference's requirements.txt:1) promotes it
        promo = kw.get("promotion", "nep50")
        if promo not in ("nep50", "legacy"):
            raise ValueError("promotion must be 'nep50' or 'legacy'")
        r = kw.get("ratio_thresh", 0.8)
        return [float(r), 1.0] if promo == "legacy" else [_f32(r), 0.0]
    if variant == M2:
        w = kw.get("saliency_weight", 0.3)
        return [_f32(w), _f32(kw.get("min_saliency", 0.2)), _f32(kw.get("min_descriptor_sim", 0.7)),
                _f32(kw.get("min_intensity", 0.1)), _f32(1 - w)]
    if variant == M3:
        return [_f32(kw.get("ratio_threshold", 0.9))]
    if variant == M4:
        return [0.0]
    if variant == M5:
        return [_f32(kw.get("match_threshold", 0.8))]
    raise ValueError(f"unknown matcher variant {variant}")


def auto_mode(D):
    """Default similarity arithmetic: the tcgen05 f16x3 tile GEMM (fp32-level accuracy) whenever its
    operand layout allows (D % 8 == 0), else the CUDA-core exact kernel."""
    return SIM_F16X3 if D % 8 == 0 else SIM_F32


def match(bank1, bank2, variant=M1, pair_index=None, num_pairs=None, mode=None, scores1=None,
          scores2=None, inten1=None, inten2=None, top=None, **kw):
    """Device-resident matching of P pairs.  Returns (pairs (P,N,2) int32 -1-padded,
    pair_scores (P,N) fp32, counts (P,) int32, top-dict).  mode None = auto_mode(D)."""
    if mode is None:
        b = bank1[0] if isinstance(bank1, (tuple, list)) else bank1
        mode = SIM_F16X3 if isinstance(bank1, (tuple, list)) else (
            SIM_BF16 if b.dtype == torch.bfloat16 else auto_mode(b.shape[-1]))
    if top is None:
        top = ops.match_top2(bank1, bank2, pair_index=pair_index, mode=mode, num_pairs=num_pairs)
    pairs, pscores, counts = ops.match_finalize(variant, top, variant_params(variant, **kw),
                                                pair_index=pair_index, scores1=scores1,
                                                scores2=scores2, inten1=inten1, inten2=inten2)
    return pairs, pscores, counts, top


def _dev(x, device, dtype=torch.float32):
    t = torch.as_tensor(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype, non_blocking=True)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("sslam_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _bank(d, mode, device):
    t = _dev(d, device)[None]
    if mode is None:
        mode = auto_mode(t.shape[-1])
    if mode == SIM_BF16:
        t = t.to(torch.bfloat16)
    return t.contiguous()


def find_matches(desc1, desc2, ratio_thresh=0.8, mode=None, promotion="nep50"):
    """M1 — visualize_matches.py:102-124 there.  NumPy (N,D),(M,D) -> list of (i, j, sim)."""
    dev = _device()
    pairs, sc, cnt, _ = match(_bank(desc1, mode, dev), _bank(desc2, mode, dev), M1, mode=mode,
                              ratio_thresh=ratio_thresh, promotion=promotion)
    n = int(cnt[0])
    p = pairs[0, :n].cpu().numpy()
    s = sc[0, :n].cpu().numpy()
    return [(int(p[k, 0]), int(p[k, 1]), s[k]) for k in range(n)]


def match_with_quality(desc1, desc2, scores1, scores2, saliency_weight=0.3, min_saliency=0.2,
                       min_descriptor_sim=0.7, intensity1=None, intensity2=None, min_intensity=0.1,
                       mode=None):
    """M2 — visualize_matches_sequence.py:106-197 there.  Returns (K',2) int64, (K',) fp32."""
    dev = _device()
    i1 = _dev(intensity1, dev)[None].contiguous() if intensity1 is not None and intensity2 is not None else None
    i2 = _dev(intensity2, dev)[None].contiguous() if i1 is not None else None
    pairs, sc, cnt, top = match(_bank(desc1, mode, dev), _bank(desc2, mode, dev), M2, mode=mode,
                                scores1=_dev(scores1, dev)[None].contiguous(),
                                scores2=_dev(scores2, dev)[None].contiguous(), inten1=i1, inten2=i2,
                                saliency_weight=saliency_weight, min_saliency=min_saliency,
                                min_descriptor_sim=min_descriptor_sim, min_intensity=min_intensity)
    n = int(cnt[0])
    if n == 0:
        # the reference prints this only when mutual matches existed but none passed (:178-180)
        mutual = int((torch.gather(top["nn21"][0], 0, top["nn12"][0].long())
                      == torch.arange(top["nn12"].shape[1], device=dev)).sum())
        if mutual:
            print(f"⚠️  No matches above thresholds (min_sal={min_saliency}, min_desc={min_descriptor_sim})")
        return np.zeros((0, 2), dtype=np.int64), np.zeros((0,), dtype=np.float32)
    return pairs[0, :n].cpu().numpy().astype(np.int64), sc[0, :n].cpu().numpy()


def find_mutual_nearest_neighbors(desc1, desc2, ratio_threshold=0.9, mode=None):
    """M3 — test/test_descriptor_quality.py:97-142 there.  Returns (K',2) int64, distances fp32."""
    dev = _device()
    pairs, sc, cnt, _ = match(_bank(desc1, mode, dev), _bank(desc2, mode, dev), M3, mode=mode,
                              ratio_threshold=ratio_threshold)
    n = int(cnt[0])
    return pairs[0, :n].cpu().numpy().astype(np.int64), sc[0, :n].cpu().numpy()


def find_matches_batched(desc1, desc2, mode=None):
    """M4 — train.py:410-449 there.  Tensors (B,N,D) on device -> int64 (B,maxM,2) padded with
    (0,0) rows; all-empty -> zeros (B,1,2)."""
    d1, d2 = desc1.contiguous(), desc2.contiguous()
    if mode is None:
        mode = auto_mode(d1.shape[-1])
    if mode == SIM_BF16:
        d1, d2 = d1.to(torch.bfloat16), d2.to(torch.bfloat16)
    pairs, _, cnt, _ = match(d1, d2, M4, mode=mode)
    mx = max(int(cnt.max()), 1)                            # host sync, as the reference has (:438)
    out = pairs[:, :mx].to(torch.int64)
    return torch.where(out < 0, torch.zeros_like(out), out)


def tracking_count(desc_prev, desc_curr, match_threshold=0.8, mode=None):
    """M5 — test/test_tracking.py:159-161 there.  NumPy in, int out."""
    dev = _device()
    _, _, cnt, _ = match(_bank(desc_prev, mode, dev), _bank(desc_curr, mode, dev), M5, mode=mode,
                         match_threshold=match_threshold)
    return int(cnt[0])
