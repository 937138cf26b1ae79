from dataclasses import dataclass, field
from typing import Any, Dict


@dataclass
class TrainerConfig:
    method_name: str = ""
    steps_per_eval_batch: int = 500
    steps_per_save: int = 1000
    max_num_iterations: int = 1000000
    mixed_precision: bool = False
    save_only_latest_checkpoint: bool = True
    pipeline: Any = None
    optimizers: Dict[str, Any] = field(default_factory=dict)
    viewer: Any = None
    vis: str = "viewer"
