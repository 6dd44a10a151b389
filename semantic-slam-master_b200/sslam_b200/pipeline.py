"""Batched, device-resident extract + match (the hot path as one object).

Mirrors the orchestration of ``MatchVisualizer.extract_features`` /
``SequenceMatcher.extract`` (visualize_matches.py:70-100, visualize_matches_sequence.py:69-104
there) without the per-frame ``.cpu().numpy()`` and without re-extracting shared frames
(visualize_matches_sequence.py:306-307): every frame is decoded once, its descriptors stay in HBM
and consecutive (or listed) pairs are matched from the resident bank.

Two coordinate conventions:
  * ``grid="pixel"`` — pipeline P: saliency at image resolution (H, W), features at (H/16, W/16);
    keypoints are pixel coordinates and ``pixel_to_patch`` is fused into the sampler.
  * ``grid="patch"`` — the reference-native case: saliency and features on the same patch grid;
    keypoints are patch coordinates, ``patch_to_pixel`` gives the pixel output.
"""

import torch

from . import matchers, ops
from .ops import SIM_BF16, SIM_F32, SIM_F16X3


class FrontEnd:
    def __init__(self, refiner, num_keypoints=2048, nms_radius=2, min_score_percentile=0.5,
                 grid="pixel", sim_mode=SIM_F16X3, patch_size=16):
        self.refiner = refiner.eval()
        self.K, self.r, self.pct = int(num_keypoints), int(nms_radius), float(min_score_percentile)
        self.grid, self.sim_mode, self.patch = grid, sim_mode, patch_size

    def _mark(self, timers, name):
        if timers is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timers.append((name, ev))

    @torch.no_grad()
    def extract(self, saliency, features, timers=None, out=None):
        """saliency (B,H,W,1)|(B,H,W), features (B,h,w,C) -> dict of device tensors:
        keypoints_pixel (B,K,2), scores (B,K), descriptors (B,K,D) [, descriptors_bf16], info.
        `timers`, when a list, receives (stage, cuda event) marks on the current stream.
        `out` (dict of preallocated slices: keypoints, scores, info, descriptors[, descriptors_bf16])
        makes the kernels write into a resident bank instead of fresh tensors."""
        self._mark(timers, "begin")
        dec_out = (out["keypoints"], out["scores"], out["info"]) if out is not None else None
        kp, sc, info = ops.decode_topk(saliency, self.K, self.r, self.pct, out=dec_out)
        self._mark(timers, "decode")
        fused = getattr(self.refiner, "mlp", "torch") == "tcgen05"
        # the tensor-core refiner consumes fp16 (hi, lo) pairs: let the sampler write them directly
        sampled = ops.gather_bilinear(features, kp, pixel_coords=(self.grid == "pixel"), pair=fused)
        self._mark(timers, "gather")
        B = kp.shape[0]
        res = dict(scores=sc, info=info)
        want16 = self.sim_mode == SIM_BF16
        o32 = out["descriptors"] if out is not None else None
        o16 = out.get("descriptors_bf16") if out is not None else None
        # f16x3 matcher: the normalisation kernel also writes the descriptors as the fp16 (hi, lo)
        # pair the matcher multiplies, so no split pass runs later
        opair = None
        if self.sim_mode == SIM_F16X3:
            if out is not None:
                opair = (out["descriptors_hi"], out["descriptors_lo"])
            else:
                D = self.refiner.output_dim
                opair = tuple(torch.empty(B * self.K, D, dtype=torch.float16, device=kp.device) for _ in range(2))
        if fused:
            res_d = self.refiner.forward_fused(sampled, want_bf16=want16, out=o32, out16=o16,
                                               out_pair=opair)                                  # MLP + L2 norm
            self._mark(timers, "refiner_mlp")
        else:
            raw = self.refiner.forward_unnormalized(sampled)
            self._mark(timers, "refiner_mlp")
            res_d = ops.l2norm_rows(raw, out=o32, out_bf16=o16, want_bf16=want16, pair=opair)
        d32, d16 = res_d if want16 else (res_d, None)
        if want16:
            res["descriptors_bf16"] = d16.reshape(B, self.K, -1)
        if opair is not None:
            res["descriptors_hi"] = opair[0].reshape(B, self.K, -1)
            res["descriptors_lo"] = opair[1].reshape(B, self.K, -1)
        self._mark(timers, "l2norm")
        res["descriptors"] = d32.reshape(B, self.K, -1)
        res["keypoints_pixel"] = kp if self.grid == "pixel" else kp * self.patch + self.patch / 2
        return res

    def bank(self, feats, start=0, stop=None):
        """Descriptor bank (frames start:stop) in the operand format of the configured matcher."""
        sl = slice(start, stop)
        if self.sim_mode == SIM_BF16:
            return feats["descriptors_bf16"][sl]
        if self.sim_mode == SIM_F16X3 and "descriptors_hi" in feats:
            return (feats["descriptors_hi"][sl], feats["descriptors_lo"][sl])
        return feats["descriptors"][sl]

    @staticmethod
    def _shift(bank, k=1):
        return tuple(t[k:] for t in bank) if isinstance(bank, tuple) else bank[k:]

    @torch.no_grad()
    def match_consecutive(self, feats, variant=matchers.M1, timers=None, **kw):
        """Match frame t with frame t+1 for every t of an extracted batch (P = B-1 pairs)."""
        bank = self.bank(feats)
        F = feats["descriptors"].shape[0]
        sc = feats["scores"]
        self._mark(timers, "begin")
        top = ops.match_top2(bank, self._shift(bank), mode=self.sim_mode, num_pairs=F - 1)
        self._mark(timers, "match_top2")
        out = matchers.match(bank, self._shift(bank), variant, num_pairs=F - 1, mode=self.sim_mode,
                             scores1=sc, scores2=sc[1:], top=top, **kw)[:3]
        self._mark(timers, "match_finalize")
        return out

    @torch.no_grad()
    def match_pairs(self, feats, pair_index, variant=matchers.M1, **kw):
        """Match the listed (a, b) frame pairs of an extracted batch (loop-closure style)."""
        bank = self.bank(feats)
        sc = feats["scores"]
        return matchers.match(bank, bank, variant, pair_index=pair_index, mode=self.sim_mode,
                              scores1=sc, scores2=sc, **kw)[:3]

    @torch.no_grad()
    def run_sequence(self, saliency, features, variant=matchers.M1, chunk=256, timers=None, **kw):
        """Extract every frame once (in chunks) and match consecutive pairs.  Returns padded pair
        lists for the T-1 pairs, all on device."""
        T = saliency.shape[0]
        feats = self._alloc_bank(T, saliency.device)
        for s in range(0, T, chunk):
            e = min(T, s + chunk)
            self.extract(saliency[s:e], features[s:e], timers=timers, out=self._bank_slice(feats, s, e))
        if self.grid != "pixel":
            feats["keypoints_pixel"] = feats["keypoints"] * self.patch + self.patch / 2
        pairs, pscores, counts = self.match_consecutive(feats, variant, timers=timers, **kw)
        return feats, pairs, pscores, counts

    def check_range(self):
        """Raises if a refiner forward since the last check left the fp16 range of the f16x3 arithmetic
        (ops.refiner_range_check; synchronises).  The device-resident entry points stay asynchronous and
        capturable: call this once after run_sequence / run_pairs when the weights are not known to be
        safe; the host entry points call it themselves."""
        if getattr(self.refiner, "mlp", "torch") == "tcgen05":
            ops.refiner_range_check()

    @torch.no_grad()
    def run_pairs(self, saliency, features, pair_index, variant=matchers.M1, chunk=256, timers=None, **kw):
        """Extract every frame once (in chunks) into a resident bank and match the listed (a, b)
        frame pairs (independent pairs, all-pairs / loop-closure lists).  Returns
        (feats, pairs, pair_scores, counts), all on device."""
        T = saliency.shape[0]
        feats = self._alloc_bank(T, saliency.device)
        for s in range(0, T, chunk):
            e = min(T, s + chunk)
            self.extract(saliency[s:e], features[s:e], timers=timers, out=self._bank_slice(feats, s, e))
        if self.grid != "pixel":
            feats["keypoints_pixel"] = feats["keypoints"] * self.patch + self.patch / 2
        self._mark(timers, "begin")
        pairs, pscores, counts = self.match_pairs(feats, pair_index, variant, **kw)
        self._mark(timers, "match")
        return feats, pairs, pscores, counts

    @torch.no_grad()
    def capture(self, fn, *args, **kw):
        """Capture ``fn(*args, **kw)`` (run_sequence / run_pairs on static input tensors) into a CUDA
        graph after two warm-up calls on a side stream.  Returns (replay, results of fn)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn(*args, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn(*args, **kw)
        return graph.replay, out

    @torch.no_grad()
    def run_pairs_host(self, saliency_host, features_host, pair_index, variant=matchers.M1, chunk=64,
                       out_host=None, **kw):
        """End-to-end entry point for HOST data and a pair LIST: pinned (T,H,W,1) saliency and
        (T,h,w,C) features stream to the device chunk by chunk (double-buffered on a copy stream,
        overlapping the extraction of the previous chunk); when every frame is resident the listed
        pairs are matched and the lists copied back.  Returns host tensors (pairs, scores, counts)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        T = saliency_host.shape[0]
        compute = torch.cuda.current_stream()
        st, copy = self._staging(saliency_host, features_host, chunk, dev)
        feats = self._alloc_bank(T, dev)
        copy.wait_stream(compute)
        for ci, s in enumerate(range(0, T, chunk)):
            e = min(T, s + chunk)
            b = ci & 1
            with torch.cuda.stream(copy):
                if ci >= 2:
                    copy.wait_event(st["free"][b])
                st["sal"][b][:e - s].copy_(saliency_host[s:e], non_blocking=True)
                st["feat"][b][:e - s].copy_(features_host[s:e], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy)
            compute.wait_event(ready)
            self.extract(st["sal"][b][:e - s], st["feat"][b][:e - s], out=self._bank_slice(feats, s, e))
            st["free"][b].record(compute)
        res = self.match_pairs(feats, pair_index, variant, **kw)
        P = pair_index.shape[0]
        if out_host is None:
            out_host = (torch.empty((P, self.K, 2), dtype=torch.int32, pin_memory=True),
                        torch.empty((P, self.K), dtype=torch.float32, pin_memory=True),
                        torch.empty((P,), dtype=torch.int32, pin_memory=True))
        for dst, src in zip(out_host, res):
            dst.copy_(src, non_blocking=True)
        compute.synchronize()
        self.check_range()
        return out_host

    def _staging(self, saliency_host, features_host, chunk, dev):
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._stage = {}
        key = (tuple(saliency_host.shape[1:]), tuple(features_host.shape[1:]), chunk)
        if self._stage.get("key") != key:
            self._stage = dict(key=key,
                               sal=[torch.empty((chunk,) + tuple(saliency_host.shape[1:]), device=dev) for _ in range(2)],
                               feat=[torch.empty((chunk,) + tuple(features_host.shape[1:]), device=dev) for _ in range(2)],
                               free=[torch.cuda.Event(), torch.cuda.Event()])
        return self._stage, self._copy_stream

    def _alloc_bank(self, T, device):
        """Resident per-sequence bank the chunked extraction writes into (no concatenation)."""
        D = self.refiner.output_dim
        bank = dict(keypoints=torch.empty(T, self.K, 2, device=device),
                    scores=torch.empty(T, self.K, device=device),
                    info=torch.empty(T, 4, dtype=torch.int32, device=device),
                    descriptors=torch.empty(T, self.K, D, device=device))
        if self.sim_mode == SIM_BF16:
            bank["descriptors_bf16"] = torch.empty(T, self.K, D, dtype=torch.bfloat16, device=device)
        if self.sim_mode == SIM_F16X3:
            bank["descriptors_hi"] = torch.empty(T, self.K, D, dtype=torch.float16, device=device)
            bank["descriptors_lo"] = torch.empty(T, self.K, D, dtype=torch.float16, device=device)
        bank["keypoints_pixel"] = bank["keypoints"]     # pixel grid: the decode output is already pixels
        return bank

    def _bank_slice(self, bank, s, e):
        return {k: v[s:e] for k, v in bank.items() if k != "keypoints_pixel"}

    @torch.no_grad()
    def capture_sequence(self, saliency, features, variant=matchers.M1, chunk=256, **kw):
        """Capture run_sequence on the given (static) input tensors into a CUDA graph.

        Every launch of the step (≈ 600 kernels of this library plus a few torch copies) is recorded
        once; ``replay()`` re-issues them with one graph launch, which removes the per-launch host
        cost (it matters when 8 ranks share the host cores).  New data is fed by copying into
        ``saliency`` / ``features`` before ``replay()``; results land in the returned tensors.
        Returns (replay, feats, pairs, pair_scores, counts)."""
        # warm-up on a side stream (lazy initialisation, workspaces, packed weights)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.run_sequence(saliency, features, variant, chunk=chunk, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            feats, pairs, pscores, counts = self.run_sequence(saliency, features, variant, chunk=chunk, **kw)
        return graph.replay, feats, pairs, pscores, counts

    @torch.no_grad()
    def run_sequence_host(self, saliency_host, features_host, variant=matchers.M1, chunk=64,
                          out_host=None, **kw):
        """End-to-end entry point for HOST data: pinned (T,H,W,1) saliency and (T,h,w,C) features
        are streamed to the device chunk by chunk on a copy stream (double-buffered, overlapping the
        kernels of the previous chunk), every frame is extracted once, consecutive pairs are matched,
        pairs are matched as soon as both of their frames are resident, and the match lists are
        copied back as they are produced.  Returns host tensors (pairs, pair_scores, counts)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        T = saliency_host.shape[0]
        compute = torch.cuda.current_stream()
        st, copy = self._staging(saliency_host, features_host, chunk, dev)
        feats = self._alloc_bank(T, dev)
        if out_host is None:
            out_host = (torch.empty((T - 1, self.K, 2), dtype=torch.int32, pin_memory=True),
                        torch.empty((T - 1, self.K), dtype=torch.float32, pin_memory=True),
                        torch.empty((T - 1,), dtype=torch.int32, pin_memory=True))
        copy.wait_stream(compute)
        for ci, s in enumerate(range(0, T, chunk)):
            e = min(T, s + chunk)
            b = ci & 1
            with torch.cuda.stream(copy):
                if ci >= 2:
                    copy.wait_event(st["free"][b])                   # previous user of this buffer done
                st["sal"][b][:e - s].copy_(saliency_host[s:e], non_blocking=True)
                st["feat"][b][:e - s].copy_(features_host[s:e], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy)
            compute.wait_event(ready)
            self.extract(st["sal"][b][:e - s], st["feat"][b][:e - s], out=self._bank_slice(feats, s, e))
            st["free"][b].record(compute)
            # match every pair whose second frame has just been extracted, while later chunks are
            # still crossing PCIe; the lists go back to the host as they are produced
            p0 = max(s - 1, 0)
            if e - 1 > p0:
                bank = self.bank(feats, p0, e)
                sc = feats["scores"][p0:e]
                res = matchers.match(bank, self._shift(bank), variant, num_pairs=e - 1 - p0, mode=self.sim_mode,
                                     scores1=sc, scores2=sc[1:], **kw)[:3]
                for dst, src in zip(out_host, res):
                    dst[p0:e - 1].copy_(src, non_blocking=True)
        compute.synchronize()
        self.check_range()
        return out_host
