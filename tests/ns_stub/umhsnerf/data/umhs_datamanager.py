from dataclasses import dataclass
from typing import Any


class UMHSDataManager:  # ref umhsnerf/data/umhs_datamanager.py:60 (Generic[TDataset])
    def __class_getitem__(cls, item):
        return cls


@dataclass
class UMHSDataManagerConfig:  # ref umhsnerf/data/umhs_datamanager.py:48-57
    _target: Any = UMHSDataManager
    dataparser: Any = None
    train_num_rays_per_batch: int = 4096
    eval_num_rays_per_batch: int = 4096
    patch_size: int = 1
