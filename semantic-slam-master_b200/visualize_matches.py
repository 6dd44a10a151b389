"""Drop-in for the matching surface of the reference's ``visualize_matches.py``.

``MatchVisualizer.extract_features`` and ``MatchVisualizer.find_matches`` keep the reference's
names, arguments and return types (visualize_matches.py:70-100, 102-124 there); the arithmetic runs
in the sm_100a kernels of libsslam_b200.  Plotting (``visualize_matches`` drawing code, cv2 /
matplotlib) is presentation only and out of scope.
"""

import numpy as np
import torch

from models.dino_backbone import DinoBackbone
from models.keypoint_selector import KeypointSelector
from models.descriptor_refiner import DescriptorRefiner
from sslam_b200 import matchers


class MatchVisualizer:
    """Frame-pair extract + match.  Either pass ``checkpoint_path``/``config_path`` like the
    reference (needs PyYAML, ``timm`` and DINOv3 weights), or inject ready modules with the
    keyword-only arguments (used with synthetic feature maps)."""

    def __init__(self, checkpoint_path: str = None, config_path: str = None, device: str = "cuda",
                 *, backbone=None, selector=None, refiner=None, config=None):
        if not torch.cuda.is_available():
            raise RuntimeError("MatchVisualizer (B200 build) needs a CUDA device; no CPU fallback")
        self.device = torch.device(device)
        if config is None and config_path is not None:
            import yaml
            with open(config_path, "r") as f:
                config = yaml.safe_load(f)
        self.config = config or {"model": {"input_size": 448, "num_keypoints": 500,
                                           "selector_hidden": 256, "descriptor_dim": 128,
                                           "refiner_hidden": 384,
                                           "backbone": "vit_small_patch16_dinov3.lvd1689m"}}
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
        self.selector.eval()
        self.refiner.eval()

    @torch.no_grad()
    def _extract_device(self, dino_features):
        """selector -> select_keypoints -> extract_at_keypoints -> refiner -> patch_to_pixel
        (visualize_matches.py:79-95 there), everything still on the device.  The head runs once."""
        sal = self.selector(dino_features)
        kp, sc = self.selector.select_keypoints(sal, num_keypoints=self.config["model"]["num_keypoints"])
        desc = self.refiner(self.backbone.extract_at_keypoints(dino_features, kp))
        return {"saliency": sal, "keypoints_patch": kp, "keypoints_pixel": self.backbone.patch_to_pixel(kp),
                "scores": sc, "descriptors": desc}

    @torch.no_grad()
    def features_from_patch_map(self, dino_features):
        """The part of extract_features after the backbone: (1,h,w,C) -> dict of NumPy arrays
        (keypoints_pixel (K,2), scores (K,), descriptors (K,D); visualize_matches.py:97-99 there)."""
        f = self._extract_device(dino_features)
        return {"keypoints_pixel": f["keypoints_pixel"][0].cpu().numpy(),
                "scores": f["scores"][0].cpu().numpy(), "descriptors": f["descriptors"][0].cpu().numpy()}

    @torch.no_grad()
    def extract_features(self, image_path: str):
        from PIL import Image
        from torchvision import transforms
        size = self.config["model"]["input_size"]
        tf = transforms.Compose([transforms.Resize((size, size)), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406],
                                                      std=[0.229, 0.224, 0.225])])
        image = Image.open(image_path).convert("RGB")
        out = self.features_from_patch_map(self.backbone(tf(image).unsqueeze(0).to(self.device)))
        out["image"] = image
        return out

    def find_matches(self, desc1: np.ndarray, desc2: np.ndarray, ratio_thresh: float = 0.8):
        """Mutual nearest neighbours + ratio test -> list of (i, j, sim), ascending i."""
        return matchers.find_matches(desc1, desc2, ratio_thresh)

    def visualize_matches(self, image1_path, image2_path, output_path=None, max_matches=100):
        """Extract + match two images and return the ``max_matches`` best (by similarity).
        Drawing is out of scope in this build."""
        f1, f2 = self.extract_features(image1_path), self.extract_features(image2_path)
        matches = self.find_matches(f1["descriptors"], f2["descriptors"], ratio_thresh=0.8)
        matches = sorted(matches, key=lambda m: m[2], reverse=True)[:max_matches]
        return {"features1": f1, "features2": f2, "matches": matches}
