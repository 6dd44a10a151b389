import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
import torch
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
for rows in (16384, 33792, 67584, 135168):
    x = torch.randn(1, rows, 384, device=dev)
    try:
        with torch.no_grad():
            for it in range(4):
                n0 = lib.sslam_launch_count()
                y = m.forward_fused(x)
                torch.cuda.synchronize()
        print(rows, "ok", float(y.float().abs().sum()))
    except Exception as e:
        print(rows, "FAILED at iteration", it, "launches in call", lib.sslam_launch_count() - n0, str(e)[:100])
        break
