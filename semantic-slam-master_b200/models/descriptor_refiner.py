"""DescriptorRefiner running on the sm_100a kernels.

Same surface and ``state_dict`` layout as the reference (models/descriptor_refiner.py:11-126
there): ``input_proj``, ``residual_blocks.{i}.{norm1,fc1,norm2,fc2}``, ``output_proj``.  Under
``torch.no_grad()`` on a CUDA device the whole forward — six Linear layers as tcgen05 f16x3 GEMMs (fp16 hi/lo pairs, three MMAs per product)
with fused bias/ReLU/residual epilogues, LayerNorms, and the final ``F.normalize`` — is one call
into ``sslam_refiner_forward_f32``.  When autograd is recording (training is
out of scope) the PyTorch ops are used so gradients exist; ``mlp="torch"`` forces that path for
A/B comparisons.
"""

import torch
import torch.nn as nn
import torch.nn.functional as F

from sslam_b200 import ops


class ResidualBlock(nn.Module):
    """Pre-LayerNorm two-layer block: relu(x + fc2(norm2(relu(fc1(norm1(x))))))."""

    def __init__(self, dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.fc1 = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim)
        self.fc2 = nn.Linear(dim, dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = F.relu(self.fc1(self.norm1(x)))
        y = self.fc2(self.norm2(y))
        return F.relu(y + x)


class DescriptorRefiner(nn.Module):
    def __init__(self, input_dim: int = 384, hidden_dim: int = 384, output_dim: int = 128,
                 num_layers: int = 4):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.input_proj = nn.Linear(input_dim, hidden_dim)
        self.residual_blocks = nn.ModuleList(ResidualBlock(hidden_dim) for _ in range(num_layers - 2))
        self.output_proj = nn.Linear(hidden_dim, output_dim)
        for m in self.modules():                         # reference init: orthogonal, U(-.1,.1) bias
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, gain=1.0)
                nn.init.uniform_(m.bias, -0.1, 0.1)

    def forward_unnormalized(self, dino_features: torch.Tensor) -> torch.Tensor:
        """(B, N, C) -> (B*N, D) raw MLP output."""
        x = dino_features.reshape(-1, dino_features.shape[-1])
        x = F.relu(self.input_proj(x))
        for blk in self.residual_blocks:
            x = blk(x)
        return self.output_proj(x)

    mlp = "tcgen05"          # "tcgen05" (default) or "torch" (cuBLAS fp32 body + l2norm kernel)
    check_range = True       # forward(): synchronise and raise if an activation left the fp16 range of the
                             # f16x3 arithmetic (the pipeline's forward_fused callers check once per run instead)

    def _ordered_params(self):
        ps = [self.input_proj.weight, self.input_proj.bias]
        for blk in self.residual_blocks:
            ps += [blk.norm1.weight, blk.norm1.bias, blk.fc1.weight, blk.fc1.bias,
                   blk.norm2.weight, blk.norm2.bias, blk.fc2.weight, blk.fc2.bias]
        return ps + [self.output_proj.weight, self.output_proj.bias]

    def _plan(self):
        ps = self._ordered_params()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_plan_key", None) != key:
            for blk in self.residual_blocks:
                if blk.norm1.eps != 1e-5 or blk.norm2.eps != 1e-5:
                    raise RuntimeError("fused refiner assumes LayerNorm eps = 1e-5")
            self._plan_obj = ops.RefinerPlan(ps, self.input_dim, self.input_proj.out_features,
                                             self.output_dim, len(self.residual_blocks))
            self._plan_key = key
        return self._plan_obj

    def forward_fused(self, dino_features, want_bf16: bool = False, out=None, out16=None, out_pair=None):
        """(..., C) fp32 — or the fp16 (hi, lo) pair of ops.gather_bilinear(pair=True) —
        -> (rows, D) unit-norm fp32 [, bf16 copy] [, fp16 pair into out_pair] through
        sslam_refiner_forward_f32."""
        return ops.refiner_forward(self._plan(), dino_features, want_bf16=want_bf16, out=out, out16=out16,
                                   out_pair=out_pair)

    def forward(self, dino_features: torch.Tensor) -> torch.Tensor:
        """(B, N, C) features at keypoints -> (B, N, output_dim) unit-norm descriptors."""
        B, N, _ = dino_features.shape
        grad = torch.is_grad_enabled() and (dino_features.requires_grad or
                                            any(p.requires_grad for p in self.parameters()))
        if grad:
            # training needs autograd; the kernels have no backward
            raw = self.forward_unnormalized(dino_features)
            return F.normalize(raw, p=2, dim=-1).reshape(B, N, self.output_dim)
        if self.mlp == "torch":
            return ops.l2norm_rows(self.forward_unnormalized(dino_features)).reshape(B, N, self.output_dim)
        out = self.forward_fused(dino_features).reshape(B, N, self.output_dim)
        if self.check_range and not torch.cuda.is_current_stream_capturing():
            ops.refiner_range_check()      # no silent NaN: raises when an activation left the fp16 range
        return out
