"""sslam_b200 — PyTorch host code over libsslam_b200.so (hand-written sm_100a CUDA kernels).

B200-native replacement for the per-frame learned-feature front-end of
Siverteh/semantic-slam-master: heatmap decode, bilinear descriptor sampling + L2 norm, and
mutual-nearest-neighbour matching.  PyTorch supplies device memory, streams and
``torch.distributed``; every hot-path operation is a kernel behind the C ABI in
``include/sslam_b200.h``.  There is no CPU fallback: operations raise if the library or an
sm_100 device is missing.
"""

from . import _lib  # noqa: F401
from .ops import (decode_topk, nms, gather_bilinear, l2norm_rows, match_top2, match_finalize,  # noqa: F401
                  RefinerPlan, refiner_forward,
                  SIM_F32, SIM_TF32X3, SIM_BF16, SIM_F16X3, M1, M2, M3, M4, M5, launch_count)
