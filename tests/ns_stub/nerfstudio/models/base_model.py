"""nerfstudio.models.base_model (1.1.5), reduced to what a model subclass relies on."""
from dataclasses import dataclass, field
from typing import Dict, Type

import torch
from torch import nn

from nerfstudio.configs.base_config import InstantiateConfig


@dataclass
class ModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Model)
    enable_collider: bool = True
    collider_params: Dict[str, float] = None
    loss_coefficients: Dict[str, float] = None
    eval_num_rays_per_chunk: int = 4096
    prompt: str = None


class Model(nn.Module):
    def __init__(self, config, scene_box, num_train_data, **kwargs):
        super().__init__()
        self.config = config
        self.scene_box = scene_box
        self.render_aabb = None
        self.num_train_data = num_train_data
        self.kwargs = kwargs
        self.collider = None
        self.populate_modules()
        self.callbacks = None
        self.device_indicator_param = nn.Parameter(torch.empty(0))

    @property
    def device(self):
        return self.device_indicator_param.device

    def populate_modules(self):
        pass

    def forward(self, ray_bundle):
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle)
