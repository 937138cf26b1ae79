from dataclasses import dataclass
from typing import Any


@dataclass
class UMHSPipelineConfig:  # ref umhsnerf/umhs_pipeline.py:34-46
    datamanager: Any = None
    model: Any = None
    num_classes: int = 5
    check_nan: bool = False
