"""pixelspointspolygons_b200 -- B200 (sm_100a) LiDAR pillar-encode + early-fusion hot path of PixelsPointsPolygons.

Host side mirrors the reference's encoder plugin interface; the compute is libp3p.so (include/p3p.h).
"""
from .config import AttrDict, default_cfg  # noqa: F401
from .encoder import PointPillarsEncoder  # noqa: F401
from .fusion import ConvBnRelu3x3, EarlyFusionFrontEnd, PatchEmbed, ProjTail  # noqa: F401
from .las import LasPackedFrontEnd, las_packed_to_pixels, las_to_pixels, pack_las  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401

__all__ = ["AttrDict", "default_cfg", "PointPillarsEncoder", "EarlyFusionFrontEnd", "PatchEmbed", "ConvBnRelu3x3", "ProjTail",
           "las_to_pixels", "las_packed_to_pixels", "pack_las", "LasPackedFrontEnd", "HostPipeline"]
