"""Ray containers.  When nerfstudio is installed its own RayBundle / RaySamples / Frustums are used
(the plugin is a drop-in); otherwise these minimal stand-ins with the same attribute names are.
Only the attributes the UMHS hot path reads are modelled (ref umhsnerf/umhs_field.py:159-183,300-304;
ref umhsnerf/umhs_model.py:225-258)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from torch import Tensor

try:  # pragma: no cover - nerfstudio is absent in the build container
    from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples  # type: ignore  # noqa: F401

    HAVE_NERFSTUDIO = True
except Exception:  # ModuleNotFoundError or a broken install
    HAVE_NERFSTUDIO = False

    @dataclass
    class Frustums:  # type: ignore[no-redef]
        origins: Tensor  # [..., 3]
        directions: Tensor  # [..., 3]
        starts: Tensor  # [..., 1]
        ends: Tensor  # [..., 1]
        pixel_area: Optional[Tensor] = None

        def get_positions(self) -> Tensor:
            return self.origins + self.directions * (self.starts + self.ends) / 2

        @property
        def shape(self):
            return self.origins.shape[:-1]

    @dataclass
    class RaySamples:  # type: ignore[no-redef]
        frustums: Frustums
        camera_indices: Optional[Tensor] = None
        # set by the fused sampler: lets the field kernels rebuild positions from the ray batch
        packed: Optional[object] = None

    @dataclass
    class RayBundle:  # type: ignore[no-redef]
        origins: Tensor  # [R, 3]
        directions: Tensor  # [R, 3]
        pixel_area: Optional[Tensor] = None
        camera_indices: Optional[Tensor] = None
        nears: Optional[Tensor] = None
        fars: Optional[Tensor] = None
        metadata: Optional[dict] = None

        def __len__(self) -> int:
            return self.origins.shape[0]

        @property
        def shape(self):
            return self.origins.shape[:-1]

        def get_row_major_sliced_ray_bundle(self, start: int, end: int) -> "RayBundle":
            def s(t):
                return None if t is None else t.reshape(-1, t.shape[-1])[start:end]

            return RayBundle(s(self.origins), s(self.directions), s(self.pixel_area), s(self.camera_indices),
                             s(self.nears), s(self.fars), self.metadata)
