"""Times the descriptor sampler alone on c2-shaped inputs (F frames, 30x40x384 maps, 2048 keypoints)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
import torch
from sslam_b200 import ops

F = int(os.environ.get("F", 300))
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
feat = torch.randn(F, 30, 40, 384, generator=g).to(dev)
kp = torch.stack([torch.randint(0, 640, (F, 2048), generator=g), torch.randint(0, 480, (F, 2048), generator=g)], -1).float().to(dev)
for _ in range(3):
    ops.gather_bilinear(feat, kp, pixel_coords=True, pair=True)
torch.cuda.synchronize()
ops.profile_enable(True)
for _ in range(5):
    ops.gather_bilinear(feat, kp, pixel_coords=True, pair=True)
torch.cuda.synchronize()
for k, (ms, n) in ops.profile_read().items():
    byt = feat.numel() * 4 + kp.numel() * 4 + F * 2048 * 384 * 4
    print(f"{k:16s} {ms / 5:8.3f} ms/call  {byt / (ms / 5 * 1e-3) / 1e9:8.0f} GB/s algorithmic")
