"""``load_experiment`` under the reference's module path (/root/reference/holo_diffusion/utils/checkpoint_utils.py:23-76).

With the full stack installed (pytorch3d + omegaconf + the reference's ``experiment.py``) the reference's own recipe is
followed: structured schema of the Experiment class merged with ``<exp_dir>/expconfig.yaml``, ``force_resume``, the
render size override, then ``experiment.model_factory(exp_dir=...)`` -- which resolves ``HoloDiffusionModel`` through
the registry to the B200 class of this package.  Without that stack (this image) the yaml is read with PyYAML, the
model is built from ``model_factory_ImplicitronModelFactory_args.model_HoloDiffusionModel_args`` and the newest
``model_epoch_*.pth`` is loaded; the data source is a stand-in that carries the one field generate_samples.py reads
(``data_loader_map_provider.batch_size``).
"""
import glob
import os
import types
from typing import Optional, Tuple

import torch

_MODEL_KEYS = ("resol", "volume_extent", "feature_size", "num_passes", "render_image_width", "render_image_height",
               "net_3d_enabled", "net_3d_class_type", "net_3d_SimpleUnet3D_args", "diffusion_enabled", "diffusion_args",
               "raysampler_class_type", "raysampler_AdaptiveRaySampler_args", "renderer_class_type",
               "renderer_HoloMultiPassEmissionAbsorptionRenderer_args", "implicit_function_class_type",
               "implicit_function_HoloVoxelGridImplicitFunction_args", "chunk_size_grid",
               # the view-pooling encoder (configs/base.yaml:160-168)
               "view_pooler_enabled", "image_feature_extractor_class_type", "image_feature_extractor_ResNetFeatureExtractor_args",
               "view_pooler_args", "mask_images", "mask_threshold", "bg_color")
_ENCODER_PREFIXES = ("image_feature_extractor.", "view_pooler.", "pooled_feature_mapper.")


def _get_config_from_experiment_directory(experiment_directory: str) -> dict:
    import yaml
    with open(os.path.join(experiment_directory, "expconfig.yaml")) as f:
        return yaml.safe_load(f)


def find_last_checkpoint(exp_dir: str) -> Optional[str]:
    fl = sorted(glob.glob(os.path.join(exp_dir, "model_epoch_" + "[0-9]" * 8 + ".pth")))
    return fl[-1] if fl else None


def load_experiment(ExperimentClass, exp_dir: str, restrict_sequence_name: Optional[str] = None,
                    render_size: Optional[Tuple[int, int]] = None, seed: int = 42,
                    device: torch.device = torch.device("cpu")):
    try:
        from omegaconf import OmegaConf
        import pytorch3d.implicitron.models.generic_model  # noqa: F401
        full_stack = ExperimentClass is not None
    except ImportError:
        full_stack = False
    if full_stack:
        schema = OmegaConf.structured(ExperimentClass)
        config = OmegaConf.merge(schema, OmegaConf.load(os.path.join(exp_dir, "expconfig.yaml")))
        config.exp_dir = exp_dir
        mf = config.model_factory_ImplicitronModelFactory_args
        mf.force_resume = True
        if render_size is not None:
            mf.model_HoloDiffusionModel_args.render_image_width = render_size[0]
            mf.model_HoloDiffusionModel_args.render_image_height = render_size[1]
        config.seed = seed
        experiment = ExperimentClass(**config)
        model = experiment.model_factory(exp_dir=exp_dir)
        model.to(device)
        return experiment, model, experiment.data_source
    from ..holo_diffusion_model import HoloDiffusionModel
    cfg = _get_config_from_experiment_directory(exp_dir)
    margs = dict(cfg["model_factory_ImplicitronModelFactory_args"]["model_HoloDiffusionModel_args"])
    if render_size is not None:
        margs["render_image_width"], margs["render_image_height"] = render_size[0], render_size[1]
    torch.manual_seed(seed)
    model = HoloDiffusionModel(**{k: v for k, v in margs.items() if k in _MODEL_KEYS})
    ckpt = find_last_checkpoint(exp_dir)
    if ckpt is not None:
        sd = torch.load(ckpt, map_location="cpu")
        missing, unexpected = model.load_state_dict(sd, strict=False)
        # the encoder side (image_feature_extractor.*, view_pooler.*, pooled_feature_mapper.*) loads when the config
        # enables the view pooler; a model built without it skips those keys, and a sampling-only checkpoint (no
        # encoder keys at all) may leave the encoder at its initialisation.  Anything else that does not match is an error
        ckpt_has_encoder = any(k.startswith(_ENCODER_PREFIXES) for k in sd)
        model_has_encoder = bool(getattr(model, "view_pooler_enabled", False))
        bad = [k for k in missing if not (k.startswith(_ENCODER_PREFIXES) and not ckpt_has_encoder)] + \
              [k for k in unexpected if not (k.startswith(_ENCODER_PREFIXES) and not model_has_encoder)]
        if bad:
            raise RuntimeError(f"checkpoint {ckpt} does not match the model: {bad[:8]} ...")
    model.to(device)
    model.n_train_target_views = int(margs.get("n_train_target_views", 1))
    bs = (cfg.get("data_source_ImplicitronDataSource_args", {})
          .get("data_loader_map_provider_SequenceDataLoaderMapProvider_args", {}).get("batch_size", 10))
    data_source = types.SimpleNamespace(data_loader_map_provider=types.SimpleNamespace(batch_size=int(bs)))
    return None, model, data_source
