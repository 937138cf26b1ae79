"""nerfstudio method registration for the B200-native UMHS model (mirror of ref umhsnerf/umhs_config.py:34-69).

Entry point: ``[project.entry-points.'nerfstudio.method_configs'] umhsnerf = 'umhs_b200.umhs_config:umhs_method'``.
nerfstudio is imported lazily: it is absent from the build container; ``HYPERPARAMS`` holds the reference's trainer
settings for the in-repo trainer (``umhs_b200.trainer.UMHSTrainer``) either way.
"""
from __future__ import annotations

HYPERPARAMS = dict(method_name="umhsnerf", steps_per_eval_batch=500, steps_per_save=2000, max_num_iterations=30000,
                   mixed_precision=True, train_num_rays_per_batch=9216 * 4, eval_num_rays_per_batch=4096,
                   eval_num_rays_per_chunk=512, lr=2e-2, eps=1e-15, lr_final=1e-5, lr_max_steps=30000)


def _build():
    """The reference's ``umhs_method`` (ref umhs_config.py:34-69) with ``model=`` the B200-native ``UMHSConfig``.  The
    trainer, viewer and the DATA side (dataparser, datamanager, dataset, pipeline glue) are nerfstudio's and the
    reference's own ``umhsnerf`` package -- outside the hot path this package replaces."""
    from nerfstudio.configs.base_config import ViewerConfig  # type: ignore
    from nerfstudio.engine.optimizers import AdamOptimizerConfig  # type: ignore
    from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig  # type: ignore
    from nerfstudio.engine.trainer import TrainerConfig  # type: ignore
    from nerfstudio.plugins.types import MethodSpecification  # type: ignore
    from umhsnerf.data.umhs_datamanager import UMHSDataManager, UMHSDataManagerConfig  # type: ignore
    from umhsnerf.data.umhs_dataparser import UMHSDataParserConfig  # type: ignore
    from umhsnerf.data.utils.hs_dataloader import HyperspectralDataset  # type: ignore
    from umhsnerf.umhs_pipeline import UMHSPipelineConfig  # type: ignore

    from .umhs_model import UMHSConfig

    h = HYPERPARAMS
    return MethodSpecification(
        config=TrainerConfig(
            method_name=h["method_name"], steps_per_eval_batch=h["steps_per_eval_batch"],
            steps_per_save=h["steps_per_save"], max_num_iterations=h["max_num_iterations"],
            mixed_precision=h["mixed_precision"], save_only_latest_checkpoint=False,
            pipeline=UMHSPipelineConfig(
                datamanager=UMHSDataManagerConfig(_target=UMHSDataManager[HyperspectralDataset],
                                                  dataparser=UMHSDataParserConfig(),
                                                  train_num_rays_per_batch=h["train_num_rays_per_batch"],
                                                  eval_num_rays_per_batch=h["eval_num_rays_per_batch"]),
                model=UMHSConfig(eval_num_rays_per_chunk=h["eval_num_rays_per_chunk"])),
            optimizers={"fields": {"optimizer": AdamOptimizerConfig(lr=h["lr"], eps=h["eps"]),
                                   "scheduler": ExponentialDecaySchedulerConfig(lr_final=h["lr_final"],
                                                                                max_steps=h["lr_max_steps"])}},
            viewer=ViewerConfig(num_rays_per_chunk=1 << 12), vis="viewer"),
        description="UMHS (UnMix-NeRF) with the B200-native hot path (libumhs_sm100).")


def __getattr__(name):
    if name == "umhs_method":
        try:
            return _build()
        except ModuleNotFoundError as e:  # pragma: no cover
            raise ModuleNotFoundError(f"umhs_method needs nerfstudio and the reference's data modules: {e}") from e
    raise AttributeError(name)
