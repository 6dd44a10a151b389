"""DinoBackbone: the reference's sampling helpers on CUDA kernels; the ViT itself is optional.

Same surface as the reference class (models/dino_backbone.py:15-178 there).  Only
``extract_at_keypoints`` / ``patch_to_pixel`` / ``pixel_to_patch`` are on the hot path; they do not
need the ViT, so the class can be built with ``load_vit=False`` (synthetic feature maps, no
``timm``, no weights — the situation in this offline image).  With ``load_vit=True`` it behaves
like the reference and requires ``timm`` plus downloadable DINOv3 weights.
"""

import torch
import torch.nn as nn

from sslam_b200 import ops


class DinoBackbone(nn.Module):
    def __init__(self, model_name: str = "vit_small_patch16_dinov3.lvd1689m", input_size: int = 448,
                 freeze: bool = True, load_vit: bool = True, embed_dim: int = 384):
        super().__init__()
        self.model_name, self.input_size = model_name, input_size
        self.patch_size = 16
        self.grid_h = self.grid_w = input_size // self.patch_size
        self.num_patches = self.grid_h * self.grid_w
        self.n_storage_tokens = 4
        self.dino = None
        if load_vit:
            try:
                import timm
            except ImportError as e:                     # pragma: no cover - image has no timm
                raise ImportError("DinoBackbone(load_vit=True) needs `timm` and DINOv3 weights; "
                                  "use load_vit=False to work from precomputed feature maps") from e
            self.dino = timm.create_model(model_name, pretrained=True, dynamic_img_size=True)
            embed_dim = self.dino.embed_dim
            if freeze:
                self.dino.requires_grad_(False)
                self.dino.eval()
        self.embed_dim = embed_dim
        self.feature_norm = nn.BatchNorm1d(self.embed_dim, affine=True)

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """(B, 3, H, W) images -> (B, grid_h, grid_w, embed_dim) NHWC patch features."""
        if self.dino is None:
            raise RuntimeError("this DinoBackbone was built with load_vit=False; feed feature maps "
                               "to extract_at_keypoints directly")
        frozen = not next(self.dino.parameters()).requires_grad
        with torch.set_grad_enabled(self.training and not frozen):
            tokens = self.dino.forward_features(images)
        patches = tokens[:, 1 + self.n_storage_tokens:, :]
        assert patches.shape[1] == self.num_patches, \
            f"Expected {self.num_patches} patches, got {patches.shape[1]}"
        b, n, c = patches.shape
        normed = self.feature_norm(patches.reshape(b * n, c))
        return normed.reshape(b, self.grid_h, self.grid_w, self.embed_dim)

    def extract_at_keypoints(self, patch_features: torch.Tensor, keypoints: torch.Tensor) -> torch.Tensor:
        """Bilinear sampling of (B, H, W, C) features at (B, N, 2) PATCH coordinates -> (B, N, C),
        with grid_sample(align_corners=True, zero padding) arithmetic."""
        if torch.is_grad_enabled() and (patch_features.requires_grad or keypoints.requires_grad):
            # training (out of scope for the kernels, which have no backward) needs autograd: the
            # reference's own ops (models/dino_backbone.py:131-150 there), as DescriptorRefiner does
            H, W = patch_features.shape[1], patch_features.shape[2]
            g = keypoints.clone()
            g[..., 0] = 2.0 * g[..., 0] / (W - 1) - 1.0
            g[..., 1] = 2.0 * g[..., 1] / (H - 1) - 1.0
            out = torch.nn.functional.grid_sample(patch_features.permute(0, 3, 1, 2), g.unsqueeze(1), mode="bilinear",
                                                  align_corners=True)
            return out.squeeze(2).permute(0, 2, 1)
        return ops.gather_bilinear(patch_features, keypoints, pixel_coords=False)

    def extract_at_pixel_keypoints(self, patch_features, pixel_keypoints):
        """Additive API: extract_at_keypoints(pixel_to_patch(k)) in one kernel."""
        return ops.gather_bilinear(patch_features, pixel_keypoints, pixel_coords=True)

    def patch_to_pixel(self, patch_coords: torch.Tensor) -> torch.Tensor:
        return patch_coords * self.patch_size + self.patch_size / 2

    def pixel_to_patch(self, pixel_coords: torch.Tensor) -> torch.Tensor:
        return (pixel_coords - self.patch_size / 2) / self.patch_size
