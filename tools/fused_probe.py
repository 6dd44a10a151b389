"""Stall breakdown of refiner_fused_kernel (cycle counters of the MMA, producer and epilogue threads)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "semantic-slam-master_b200")]
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import ops, _lib

F = int(os.environ.get("F", 300))
rows = F * 2048
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = DescriptorRefiner(384, 384, 256, 4).to(dev).eval()
x = torch.randn(1, rows, 384, device=dev)
lib = _lib.load()
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.sslam_debug_gemm_stalls.argtypes = [ctypes.c_void_p]
lib.sslam_debug_gemm_stalls.restype = None
for S in [int(v) for v in os.environ.get("MODES", "2,3,4").split(",")]:
    lib.sslam_debug_refiner_fused(S)
    with torch.no_grad():
        for _ in range(2):
            m(x)
        torch.cuda.synchronize()
        dbg.zero_()
        lib.sslam_debug_gemm_stalls(ctypes.c_void_p(dbg.data_ptr()))
        ops.profile_enable(True)
        m(x)
        torch.cuda.synchronize()
        prof = ops.profile_read()
        ops.profile_enable(False)
        lib.sslam_debug_gemm_stalls(ctypes.c_void_p(0))
    d = dbg.cpu().numpy().reshape(148, 16).astype(np.float64)
    used = d[d[:, 5] + d[:, 4] + d[:, 6] > 0]
    lead = used[used[:, 0] > 0]
    ntiles = -(-(-(-rows // 256)) // 22) * 6
    print(f"S = {S}: gemm {prof['gemm_f16x3'][0]:.3f} ms; tiles per cluster {ntiles}")
    print("  MMA thread  : total %.0f cyc = %.0f per tile; wait A %.0f, wait accumulator free %.0f, wait weights %.0f (per tile)"
          % (lead[:, 0].mean(), lead[:, 0].mean() / ntiles, lead[:, 1].mean() / ntiles, lead[:, 2].mean() / ntiles, lead[:, 3].mean() / ntiles))
    print("  A producer  : wait ready %.0f, wait stage empty %.0f (per tile)" % (used[:, 4].mean() / ntiles, used[:, 5].mean() / ntiles))
    names = ["wait acc full", "residual issue", "tmem ld wait", "math (+ LN scalars wait)", "-", "split + stores",
             "tail (coarse build: whole tile body)", "-"]
    print("  epilogue w2 : total %.0f per tile: " % (used[:, 6].mean() / ntiles)
          + ", ".join("%s %.0f" % (n, used[:, 7 + i].mean() / ntiles) for i, n in enumerate(names)))
lib.sslam_debug_refiner_fused(3)
