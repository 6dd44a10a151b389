"""Times the DescriptorRefiner forward (384-384-384-256, 2 blocks) on F c2 frames of 2048 keypoints:
per-layer launches against the layer-fused kernel at chunk depths S = 1..4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
import torch
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import ops, _lib

F = int(os.environ.get("F", 300))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
CIN, HD = int(os.environ.get('CIN', 384)), int(os.environ.get('HD', 384))
m = DescriptorRefiner(CIN, HD, 256, 4).to(dev).eval()
x = torch.randn(1, F * 2048, CIN, device=dev)
lib = _lib.load()
MODES = [int(v) for v in os.environ.get("MODES", "0,1,2,3,4").split(",")]
REPS = int(os.environ.get("REPS", 5))
for mode in MODES:
    lib.sslam_debug_refiner_fused(mode)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        ops.profile_enable(True)
        for _ in range(REPS):
            m(x)
        torch.cuda.synchronize()
    prof = ops.profile_read()
    ops.profile_enable(False)
    print("per-layer launches" if mode == 0 else f"layer-fused, S = {mode}")
    for k, (ms, n) in prof.items():
        print(f"  {k:14s} {ms / REPS:8.3f} ms/call over {n // REPS} launches")
lib.sslam_debug_refiner_fused(3)
