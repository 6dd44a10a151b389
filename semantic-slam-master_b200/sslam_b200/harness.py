"""Shared construction of backbone / selector / refiner for the drop-in evaluation classes.

The reference's testers build the three modules from a YAML config and load a checkpoint
(e.g. test/test_descriptor_quality.py:32-67 there).  Here the same classes can also be handed
ready modules (keyword-only), which is how they are used without a ViT: ``backbone`` may be a
``DinoBackbone(load_vit=False)`` and the ``image`` arguments then carry patch-feature maps
(B,h,w,C) that are already past the backbone.
"""

import torch

DEFAULT_CONFIG = {"model": {"input_size": 448, "num_keypoints": 500, "selector_hidden": 256,
                            "descriptor_dim": 128, "refiner_hidden": 384,
                            "backbone": "vit_small_patch16_dinov3.lvd1689m"}}


class ModelHarness:
    def __init__(self, checkpoint_path: str = None, config_path: str = None, device: str = "cuda", *,
                 backbone=None, selector=None, refiner=None, config=None):
        from models.dino_backbone import DinoBackbone
        from models.keypoint_selector import KeypointSelector
        from models.descriptor_refiner import DescriptorRefiner
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 build needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device)
        if config is None and config_path is not None:
            import yaml
            with open(config_path, "r") as f:
                config = yaml.safe_load(f)
        self.config = config or DEFAULT_CONFIG
        m = self.config["model"]
        self.backbone = (backbone if backbone is not None else DinoBackbone(
            model_name=m["backbone"], input_size=m["input_size"], freeze=True)).to(self.device)
        self.selector = (selector if selector is not None else KeypointSelector(
            input_dim=self.backbone.embed_dim, hidden_dim=m["selector_hidden"])).to(self.device)
        self.refiner = (refiner if refiner is not None else DescriptorRefiner(
            input_dim=self.backbone.embed_dim, hidden_dim=m["refiner_hidden"],
            output_dim=m["descriptor_dim"])).to(self.device)
        if checkpoint_path is not None:
            ckpt = torch.load(checkpoint_path, map_location=self.device)
            self.selector.load_state_dict(ckpt["selector_state_dict"])
            self.refiner.load_state_dict(ckpt["refiner_state_dict"])
        self.backbone.eval()
        self.selector.eval()
        self.refiner.eval()

    def _features(self, image):
        """(B,3,H,W) images go through the ViT; (B,h,w,C) maps with C = embed_dim are taken as patch
        features that are already past it."""
        if image.dim() == 4 and image.shape[-1] == self.backbone.embed_dim and image.shape[1] != 3:
            return image
        return self.backbone(image)

    @torch.no_grad()
    def forward_pass(self, image: torch.Tensor):
        """PerformanceTester.forward_pass (test/test_performance.py:146-157 there): backbone ->
        selector -> select_keypoints -> extract_at_keypoints -> refiner; device tensors
        (keypoints_patch (B,K,2), descriptors (B,K,D), scores (B,K))."""
        dino_features = self._features(image)
        saliency_map = self.selector(dino_features)
        keypoints_patch, scores = self.selector.select_keypoints(
            saliency_map, num_keypoints=self.config["model"]["num_keypoints"])
        feat_at_kpts = self.backbone.extract_at_keypoints(dino_features, keypoints_patch)
        descriptors = self.refiner(feat_at_kpts)
        return keypoints_patch, descriptors, scores

    @torch.no_grad()
    def extract_features(self, image: torch.Tensor):
        """DescriptorQualityTester.extract_features / TrackingTester.extract_features
        (test/test_descriptor_quality.py:69-95, test/test_tracking.py:63-85 there): NumPy
        (keypoints_pixel (K,2), descriptors (K,D), scores (K,)) of the first batch element."""
        keypoints_patch, descriptors, scores = self.forward_pass(image)
        keypoints_pixel = self.backbone.patch_to_pixel(keypoints_patch)
        return (keypoints_pixel[0].cpu().numpy(), descriptors[0].cpu().numpy(), scores[0].cpu().numpy())
