"""Times sslam_match_top2 alone on a c2-shaped bank (with and without the MMA-thread stall counters,
alternating, in one process) and prints the stall breakdown."""
import ctypes, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "semantic-slam-master_b200"))
from sslam_b200 import ops, _lib

F, K, D = int(os.environ.get("F", 300)), 2048, 256
mode = {"f16x3": ops.SIM_F16X3, "bf16": ops.SIM_BF16, "tf32x3": ops.SIM_TF32X3}[os.environ.get("MODE", "f16x3")]
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
bank = torch.nn.functional.normalize(torch.randn(F, K, D, generator=g), dim=-1).to(dev)
if mode == ops.SIM_BF16:
    bank = bank.to(torch.bfloat16)
lib = _lib.load()
dbg = torch.zeros(148 * 4, dtype=torch.int64, device=dev)
lib.sslam_debug_match_stalls.argtypes = [ctypes.c_void_p]
lib.sslam_debug_match_stalls.restype = None


def run(n, with_dbg):
    lib.sslam_debug_match_stalls(ctypes.c_void_p(dbg.data_ptr() if with_dbg else 0))
    lib.sslam_profile_enable(1)
    for _ in range(n):
        ops.match_top2(bank, bank[1:], mode=mode, num_pairs=F - 1)
    torch.cuda.synchronize()
    out = {}
    for k in range(lib.sslam_profile_kinds()):
        ms, cnt = ctypes.c_double(), ctypes.c_uint64()
        lib.sslam_profile_read(k, ctypes.byref(ms), ctypes.byref(cnt))
        if cnt.value:
            out[lib.sslam_profile_kind_name(k).decode()] = ms.value / cnt.value
    lib.sslam_profile_enable(0)
    return out


run(3, False)
for rep in range(3):
    for wd in (False, True):
        o = run(5, wd)
        print("counters %-3s  match_tc %.3f ms/launch  (%s)" % ("on" if wd else "off", o["match_tc"],
              ", ".join("%s %.3f" % kv for kv in o.items() if kv[0] != "match_tc")))
d = dbg.cpu().numpy().reshape(148, 4).astype(np.float64)
tiles = (F - 1) * 16 * 16 / 148
print("MMA thread cycles: total %.0f  wait_full %.0f  wait_tempty %.0f  wait_afull %.0f  (per tile: %.0f / %.0f / %.0f / %.0f)"
      % (*d.mean(0), *(d.mean(0) / tiles)))
