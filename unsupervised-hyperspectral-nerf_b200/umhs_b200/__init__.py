"""umhs_b200: B200-native (sm_100a) hot path of UMHS / UnMix-NeRF behind the reference's plugin interfaces.

``UMHSField`` / ``SpectralRenderer`` (``UMHSRenderer``) / ``UMHSModel`` / ``UMHSConfig`` mirror
Factral/unsupervised-hyperspectral-nerf; every per-sample and per-ray computation runs in libumhs_sm100.so.
"""
from .umhs_field import UMHSField  # noqa: F401
from .umhs_renderer import SpectralRenderer, UMHSRenderer, get_weights_spectral  # noqa: F401
from .umhs_model import UMHSConfig, UMHSModel  # noqa: F401
from .spec_to_rgb import ColourSystem  # noqa: F401
from .clusterprobe import ClusterLookup  # noqa: F401
from .occ_grid import OccGridEstimator, VolumetricSampler  # noqa: F401
from .rays import Frustums, RayBundle, RaySamples  # noqa: F401
from .umhs_datamanager import UMHSDataManager  # noqa: F401

__all__ = ["UMHSField", "SpectralRenderer", "UMHSRenderer", "get_weights_spectral", "UMHSConfig", "UMHSModel",
           "ColourSystem", "ClusterLookup", "OccGridEstimator", "VolumetricSampler", "Frustums", "RayBundle",
           "RaySamples", "UMHSDataManager"]
