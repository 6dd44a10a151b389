"""Oracle: DescriptorRefiner forward pass in NumPy (test infrastructure, see oracle/__init__.py).

Restates ``DescriptorRefiner.forward`` and ``ResidualBlock.forward``
(models/descriptor_refiner.py:58-91, 108-126): Linear -> ReLU -> [pre-LN residual blocks] ->
Linear -> L2 normalise.  GEMM accumulation order is library specific, so results are compared
with a tolerance (descriptors within 1e-5 abs).
"""

import numpy as np

from .gather import l2_normalize

F32 = np.float32
LN_EPS = 1e-5            # torch.nn.LayerNorm default, models/descriptor_refiner.py:103,105


class RefinerWeights:
    """fp32 parameter arrays keyed exactly like the reference ``state_dict``
    (``input_proj.*``, ``residual_blocks.{i}.{norm1,fc1,norm2,fc2}.*``, ``output_proj.*``;
    models/descriptor_refiner.py:35-44, 103-106)."""

    def __init__(self, params):
        self.p = {k: np.ascontiguousarray(np.asarray(v, dtype=F32)) for k, v in params.items()}
        n = 0
        while f"residual_blocks.{n}.fc1.weight" in self.p:
            n += 1
        self.num_blocks = n

    @classmethod
    def from_state_dict(cls, state_dict):
        return cls({k: (v.detach().cpu().numpy() if hasattr(v, "detach") else v)
                    for k, v in state_dict.items()})


def _linear(x, w, b):
    return (x @ w.T + b).astype(F32)


def _layer_norm(x, g, b):
    x64 = x.astype(np.float64)
    mu = x64.mean(axis=-1, keepdims=True)
    var = ((x64 - mu) ** 2).mean(axis=-1, keepdims=True)
    return (((x64 - mu) / np.sqrt(var + LN_EPS)) * g + b).astype(F32)


def refiner_forward(weights, feats, normalize=True):
    """feats (B, N, C) fp32 -> descriptors (B, N, D) fp32, unit norm
    (models/descriptor_refiner.py:70-89)."""
    p = weights.p
    B, N, C = feats.shape
    x = np.asarray(feats, dtype=F32).reshape(B * N, C)                       # :73
    x = np.maximum(_linear(x, p["input_proj.weight"], p["input_proj.bias"]), 0)   # :76
    for i in range(weights.num_blocks):                                      # :79-80, :108-126
        pre = f"residual_blocks.{i}."
        h = _layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        h = np.maximum(_linear(h, p[pre + "fc1.weight"], p[pre + "fc1.bias"]), 0)
        h = _layer_norm(h, p[pre + "norm2.weight"], p[pre + "norm2.bias"])
        h = _linear(h, p[pre + "fc2.weight"], p[pre + "fc2.bias"])
        x = np.maximum(h + x, 0).astype(F32)
    d = _linear(x, p["output_proj.weight"], p["output_proj.bias"])           # :83
    if normalize:
        d = l2_normalize(d)                                                  # :86
    return d.reshape(B, N, -1)                                               # :89
