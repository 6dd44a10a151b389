#!/usr/bin/env python
"""Micro-benchmark of sslam_refiner_forward_f32 (one 64-frame chunk: 131072 rows, 384-384-256 x 4 layers):
per-kernel CUDA-event times via the library profiler.  Usage: python tools/bench_refiner.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
import torch
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
x = torch.randn(1, rows, 384, device=dev)
with torch.no_grad():
    for _ in range(3):
        m.forward_fused(x)
    torch.cuda.synchronize()
    ops.profile_enable(True)
    for _ in range(10):
        m.forward_fused(x)
    torch.cuda.synchronize()
    prof = ops.profile_read()
    ops.profile_enable(False)
flops = 2.0 * rows * (384 * 384 * 5 + 384 * 256)
tot = 0.0
for k, (ms, n) in prof.items():
    print(f"{k:16s} {ms / 10:8.3f} ms/call  ({n // 10} launches)")
    tot += ms / 10
g = prof["gemm_f16x3"][0] / 10
print(f"total {tot:.3f} ms; GEMMs {flops / g / 1e9:.1f} TFLOP/s algorithmic; {g / 6 * 1e3:.1f} us per GEMM")
