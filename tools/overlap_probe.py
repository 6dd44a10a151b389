"""Does it pay to run the HBM-bound kernels of chunk c+1 (decode, sample) next to the tensor-bound refiner of
chunk c?  The refiner occupies 132 of 148 SMs with one CTA each; on a low-priority stream the other kernels'
CTAs can only take the SMs it leaves free.  Compares one c2 step (600 frames) serial vs overlapped."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
import torch
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import matchers, ops, synth
from sslam_b200.pipeline import FrontEnd

T, K = int(os.environ.get("T", 600)), 2048
dev = torch.device("cuda", 0)
torch.manual_seed(0)
refiner = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
sal, feat = synth.make_sequence(60, seq_id=0)
reps = (T + 59) // 60
sal = sal.to(dev).repeat(reps, 1, 1, 1)[:T].contiguous()
feat = feat.to(dev).repeat(reps, 1, 1, 1)[:T].contiguous()
fe = FrontEnd(refiner, num_keypoints=K, grid="pixel")
lo_prio, hi_prio = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
hi = torch.cuda.Stream(priority=-1)


def step(chunk, overlap):
    main = torch.cuda.current_stream()
    bank = fe._alloc_bank(T, dev)
    keep = []
    prev = None
    for s in range(0, T, chunk):
        e = min(T, s + chunk)
        out = fe._bank_slice(bank, s, e)
        kp, sc, info = ops.decode_topk(sal[s:e], K, fe.r, fe.pct, out=(out["keypoints"], out["scores"], out["info"]))
        sampled = ops.gather_bilinear(feat[s:e], kp, pixel_coords=True, pair=True)
        keep.append(sampled)
        if overlap:
            ev = torch.cuda.Event()
            ev.record(main)
            hi.wait_event(ev)
            with torch.cuda.stream(hi):
                refiner.forward_fused(sampled, out=out["descriptors"], out_pair=(out["descriptors_hi"], out["descriptors_lo"]))
        else:
            refiner.forward_fused(sampled, out=out["descriptors"], out_pair=(out["descriptors_hi"], out["descriptors_lo"]))
    if overlap:
        main.wait_stream(hi)
    res = fe.match_consecutive(bank, matchers.M1, ratio_thresh=0.8)
    return bank, res, keep


with torch.no_grad():
    ref = None
    for chunk, overlap in ((300, False), (150, False), (150, True), (100, True), (75, True), (300, True)):
        for _ in range(2):
            out = step(chunk, overlap)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = step(chunk, overlap)
            b.record()
            b.synchronize()
            ms.append(a.elapsed_time(b))
        counts = out[1][2]
        if ref is None:
            ref = counts.clone()
        print(f"chunk {chunk:4d} overlap {overlap!s:5}: {min(ms):7.3f} ms / step (min of 5), median {sorted(ms)[2]:7.3f}; "
              f"match counts identical to serial: {bool(torch.equal(counts, ref))}")
