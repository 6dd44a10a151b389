"""Drop-in `models` package: same module/class/method names as the reference's
``semantic-slam/models`` (put ``semantic-slam-master_b200/`` on ``sys.path`` where the reference
puts ``semantic-slam/``).  The hot-path methods call the sm_100a kernels in libsslam_b200."""

from .keypoint_selector import KeypointSelector  # noqa: F401
from .descriptor_refiner import DescriptorRefiner, ResidualBlock  # noqa: F401
from .dino_backbone import DinoBackbone  # noqa: F401
