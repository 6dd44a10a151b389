"""Single-pair latency (BASELINE.json configs[0]-shaped: ONE 640x480 frame pair, extract both frames + match):
reported as latency, not as a roofline fraction — two frames cannot fill 148 SMs.

  device-resident : inputs in HBM, FrontEnd.run_sequence on 2 frames; eager launches and one CUDA-graph replay
  host to host    : pinned host buffers in, match list back on the host (FrontEnd.run_sequence_host)
  reference-named : MatchVisualizer.extract_features x 2 + find_matches (patch-grid surface, host arrays out)
Median / p90 over REPS repetitions, CUDA events (device) or perf_counter around synchronised calls (host)."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
import torch
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import matchers, synth
from sslam_b200.pipeline import FrontEnd

REPS = int(os.environ.get("REPS", 200))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
refiner = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
sal_h, feat_h = synth.make_sequence(2, seq_id=0)
sal_h, feat_h = sal_h.pin_memory(), feat_h.pin_memory()
sal, feat = sal_h.to(dev), feat_h.to(dev)


def stats(v):
    v = sorted(v)
    return "median %.1f us, p90 %.1f us" % (statistics.median(v) * 1e3, v[int(0.9 * len(v))] * 1e3)


for K in (1024, 2048):
    fe = FrontEnd(refiner, num_keypoints=K, grid="pixel")
    for _ in range(5):
        fe.run_sequence(sal, feat, matchers.M1, ratio_thresh=0.8)
    torch.cuda.synchronize()
    ms = []
    for _ in range(REPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fe.run_sequence(sal, feat, matchers.M1, ratio_thresh=0.8)
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    print(f"K={K}  device-resident, eager launches      : {stats(ms)}")
    replay, feats, pairs, pscores, counts = fe.capture_sequence(sal, feat, matchers.M1, ratio_thresh=0.8)
    for _ in range(5):
        replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(REPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        replay()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    print(f"K={K}  device-resident, CUDA-graph replay   : {stats(ms)}   (matches: {int(counts[0])})")
    for _ in range(5):
        fe.run_sequence_host(sal_h, feat_h, matchers.M1, ratio_thresh=0.8)
    ms = []
    for _ in range(REPS):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fe.run_sequence_host(sal_h, feat_h, matchers.M1, ratio_thresh=0.8)      # synchronises before returning
        ms.append((time.perf_counter() - t0) * 1e3)
    print(f"K={K}  host to host (H2D 6.1 MB, lists back): {stats(ms)}")
