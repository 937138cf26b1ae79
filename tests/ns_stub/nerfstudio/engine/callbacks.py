from dataclasses import dataclass
from enum import Enum, auto
from typing import Callable, List


class TrainingCallbackLocation(Enum):
    BEFORE_TRAIN_ITERATION = auto()
    AFTER_TRAIN_ITERATION = auto()


@dataclass
class TrainingCallback:
    where_to_run: List[TrainingCallbackLocation]
    func: Callable
    update_every_num_iters: int = 1
