"""Where does the layer-fused refiner differ from the per-layer path?  Prints the error structure."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import _lib

dev = torch.device("cuda", 0)
lib = _lib.load()
cases = [(384, 384, 256, 4, 2125, 3), (384, 384, 256, 4, 30000, 1), (384, 384, 256, 4, 30000, 2), (384, 384, 256, 4, 30000, 3)]
for (C, Hd, D, layers, rows, S) in cases:
    torch.manual_seed(1)
    m = DescriptorRefiner(C, Hd, D, layers).to(dev).eval()
    x = torch.randn(1, rows, C, generator=torch.Generator().manual_seed(rows)).to(dev)
    lib.sslam_debug_refiner_fused(0)
    with torch.no_grad():
        ref = m(x).clone()[0]
    lib.sslam_debug_refiner_fused(S)
    for rep in range(2):
        with torch.no_grad():
            out = m(x)[0]
        torch.cuda.synchronize()
        bad = (out != ref)
        nb = int(bad.sum())
        print(f"C{C} Hd{Hd} D{D} layers{layers} rows{rows} S{S} rep{rep}: {nb} differing elements, max abs {float((out - ref).abs().max()):.3e}")
        if nb:
            br = bad.any(1).nonzero()[:, 0].cpu().numpy()
            bc = bad.any(0).nonzero()[:, 0].cpu().numpy()
            sp = np.unique(br // 256)
            print("   bad rows %d in %d strip pairs: first %s; strip pairs mod 22: %s" % (len(br), len(sp), br[:8], np.unique(sp % 22)[:22]))
            print("   strip-pair index // 22 (position in cluster's list):", np.unique(sp // 22))
            print("   rank (row%256//128):", np.unique(br % 256 // 128), " quarter:", np.unique(br % 128 // 32), " cols: %d bad, tiles %s" % (len(bc), np.unique(bc // 64)))
lib.sslam_debug_refiner_fused(3)
