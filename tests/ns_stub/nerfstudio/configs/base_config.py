from dataclasses import dataclass, field
from typing import Any, Type


@dataclass
class PrintableConfig:
    pass


@dataclass
class InstantiateConfig(PrintableConfig):
    _target: Type = object

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class ViewerConfig(PrintableConfig):
    num_rays_per_chunk: int = 32768
