"""DescriptorRefiner whose L2-normalising tail is a CUDA kernel.

Same surface and ``state_dict`` layout as the reference (models/descriptor_refiner.py:11-126
there): ``input_proj``, ``residual_blocks.{i}.{norm1,fc1,norm2,fc2}``, ``output_proj``.  The MLP
body runs on cuBLAS through PyTorch (SURVEY.md §8(f) N1 lists its fusion as the next step);
``F.normalize`` is replaced by ``sslam_l2norm_rows``, which can also emit the bf16 copy the
tensor-core matcher consumes.
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

    def forward(self, dino_features: torch.Tensor) -> torch.Tensor:
        """(B, N, C) features at keypoints -> (B, N, output_dim) unit-norm descriptors."""
        B, N, _ = dino_features.shape
        raw = self.forward_unnormalized(dino_features)
        if raw.requires_grad:
            # training needs autograd through the normalisation; the kernel has no backward
            return F.normalize(raw, p=2, dim=-1).reshape(B, N, self.output_dim)
        return ops.l2norm_rows(raw).reshape(B, N, self.output_dim)
