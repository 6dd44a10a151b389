"""Times the refiner GEMMs on one chunk and prints the stall breakdown of gemm_pair_kernel
(last launch = output projection; set LAYERS=1 to stop after the input projection is not supported —
the counters are those of the last launch)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import ops, _lib

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 147456
D = int(os.environ.get("D", 384))       # D = 384 makes the last launch a hidden-layer-shaped GEMM
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = DescriptorRefiner(384, 384, D, 4).to(dev).eval()
x = torch.randn(1, rows, 384, device=dev)
lib = _lib.load()
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
lib.sslam_debug_gemm_stalls.argtypes = [ctypes.c_void_p]
lib.sslam_debug_gemm_stalls.restype = None
with torch.no_grad():
    for _ in range(3):
        m.forward_fused(x)
    torch.cuda.synchronize()
    lib.sslam_debug_gemm_stalls(ctypes.c_void_p(dbg.data_ptr()))
    ops.profile_enable(True)
    for _ in range(5):
        m.forward_fused(x)
    torch.cuda.synchronize()
    prof = ops.profile_read()
for k, (ms, n) in prof.items():
    print(f"{k:16s} {ms / 5:8.3f} ms/call  ({n // 5} launches)")
d = dbg.cpu().numpy().reshape(148, 8).astype(np.float64)
lead = d[0::2]
lead = lead[lead[:, 0] > 0]
tiles = -(-(-(-rows // 256)) // ((148 // 2) // 3))   # strip pairs per CTA pair at N = 384
print("tiles per CTA %d" % tiles)
print("MMA thread (leaders): total %.0f  wait_full %.0f  wait_tempty %.0f  wait_weights %.0f   per tile: %.0f / %.0f / %.0f"
      % (*lead[:, :4].mean(0), *(lead[:, :3].mean(0) / tiles)))
