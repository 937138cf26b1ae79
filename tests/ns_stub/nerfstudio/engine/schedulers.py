from dataclasses import dataclass


@dataclass
class ExponentialDecaySchedulerConfig:
    lr_final: float = 1e-5
    max_steps: int = 100000
