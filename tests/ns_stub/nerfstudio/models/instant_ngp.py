from dataclasses import dataclass, field
from typing import Optional, Type

from nerfstudio.models.base_model import Model, ModelConfig


@dataclass
class InstantNGPModelConfig(ModelConfig):
    _target: Type = field(default_factory=lambda: NGPModel)
    enable_collider: bool = False
    grid_resolution: int = 128
    grid_levels: int = 4
    max_res: int = 2048
    log2_hashmap_size: int = 19
    alpha_thre: float = 0.01
    cone_angle: float = 0.004
    render_step_size: Optional[float] = None
    near_plane: float = 0.05
    far_plane: float = 1e3
    use_gradient_scaling: bool = False
    use_appearance_embedding: bool = False
    background_color: str = "random"
    disable_scene_contraction: bool = False


class NGPModel(Model):
    config: InstantNGPModelConfig

    def __init__(self, config, **kwargs):
        super().__init__(config=config, **kwargs)
