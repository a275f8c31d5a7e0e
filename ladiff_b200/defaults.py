"""Shipped configurations (``ladiff_b200/configs``): the hot-path subset of the reference's YAML tree with the
``target:`` strings already pointing at the B200 classes."""
import os

from .config import Cfg, load_config

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


def default_config(dataset: str = "humanml3d", num_inference_timesteps: int = None, overrides: dict = None) -> Cfg:
    """dataset: 'humanml3d' (263 features / 22 joints) or 'kit' (251 / 21)."""
    if dataset not in ("humanml3d", "kit"):
        raise ValueError("dataset must be 'humanml3d' or 'kit'")
    cfg = load_config(os.path.join(CONFIG_DIR, f"config_ladiff_{dataset}.yaml"), CONFIG_DIR, overrides=overrides)
    if num_inference_timesteps is not None:
        cfg.model.scheduler.num_inference_timesteps = int(num_inference_timesteps)
    return cfg
