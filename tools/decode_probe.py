"""Times the decode kernels alone on c2-shaped maps (F frames 640x480, K=2048)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
import torch
from sslam_b200 import ops, synth

F = int(os.environ.get("F", 300))
dev = torch.device("cuda", 0)
sal, _ = synth.make_sequence(min(F, 60), seq_id=0)
sal = sal.to(dev)
sal = sal.repeat((F + sal.shape[0] - 1) // sal.shape[0], 1, 1, 1)[:F].contiguous()
from sslam_b200 import _lib
TUNES = [tuple(int(v) for v in t.split(':')) for t in os.environ.get('TUNES', '0:0').split(',')]
for stream, tune in [(1, t) for t in TUNES] + [(0, (0, 0))]:
    _lib.load().sslam_debug_decode_stream(stream)
    _lib.load().sslam_debug_decode_tune(*tune)
    for _ in range(3):
        ops.decode_topk(sal, 2048, 2, 0.5)
    torch.cuda.synchronize()
    ops.profile_enable(True)
    for _ in range(5):
        ops.decode_topk(sal, 2048, 2, 0.5)
    torch.cuda.synchronize()
    print(f"streaming scan + histogram top-k, stages:band_rows = {tune}" if stream else "register-prefetch scan + radix top-k")
    for k, (ms, n) in ops.profile_read().items():
        print(f"  {k:16s} {ms / 5:8.3f} ms/call  {sal.numel() * 4 / (ms / 5 * 1e-3) / 1e9:8.0f} GB/s of map bytes")
    ops.profile_enable(False)
_lib.load().sslam_debug_decode_stream(1)
_lib.load().sslam_debug_decode_tune(0, 0)
