"""CPU oracle for the extract+match hot path — TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference's per-frame learned-feature
front-end (heatmap decode, bilinear descriptor sampling + L2 norm, mutual-NN
matching).  It exists to check the CUDA path, never to serve it:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
  ``--impl reference`` legs may import it;
* nothing under ``semantic-slam-master_b200/`` imports it, and the product path raises
  when the CUDA library is missing instead of falling back to this code.

Parity status: the reference holds **no** golden vectors, known-answer tests or
fixtures for this path, so the oracle is pinned against outputs of
the reference's own functions executed in the build container
(``oracle/pin_against_reference.py`` imports ``/root/reference/semantic-slam`` with
stub ``timm``/``matplotlib`` modules and writes ``tests/golden/*.npz``).  Those
fixtures travel to the GPU box; ``/root/reference`` does not.

Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference/semantic-slam/``).  Library arithmetic the reference delegates to
(torch 2.11.0 ATen on CPU, NumPy 2.3.5) is restated from its observable behaviour:
``torch.quantile`` (fp32 rank, FMA lerp), ``max_pool2d`` (-inf padding),
``grid_sampler_2d`` (align_corners un-normalisation, zero padding), ``F.normalize``.

Tie order: ``torch.topk`` leaves the order of equal scores unspecified; the oracle
(and the CUDA kernels) define it as (score descending, linear index y*W+x ascending).
``argmax`` returns the lowest maximal index, as NumPy/torch-CPU do.
"""

from .decode import select_keypoints, apply_nms, quantile_f32, sigmoid_f32  # noqa: F401
from .gather import (extract_at_keypoints, patch_to_pixel, pixel_to_patch,  # noqa: F401
                     l2_normalize)
from .refiner import RefinerWeights, refiner_forward  # noqa: F401
from .match import (similarity_top2, match_m1, match_m2, match_m3, match_m4,  # noqa: F401
                    match_m5)
from . import evaluation  # noqa: F401,E402
