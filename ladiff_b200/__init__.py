"""ladiff_b200 -- B200-native (sm_100a) implementation of LADiff's sampling hot path behind the reference's
plugin API.  See DESIGN.md / INTEGRATION.md.  The CUDA extension is mandatory: there is no CPU fallback."""
from .config import Cfg, instantiate_from_config, load_config, retarget  # noqa: F401
from .scheduler import DDIMScheduler, DDPMScheduler  # noqa: F401

__all__ = ["Cfg", "instantiate_from_config", "load_config", "retarget", "DDIMScheduler", "DDPMScheduler",
           "LADiffDenoiser", "LADiffVae", "LADIFF", "MldTextEncoder", "default_config"]


def __getattr__(name):   # lazy: torch.nn modules are only needed once a model is built
    if name == "LADiffDenoiser":
        from .denoiser import LADiffDenoiser
        return LADiffDenoiser
    if name == "LADiffVae":
        from .vae import LADiffVae
        return LADiffVae
    if name == "LADIFF":
        from .modeltype import LADIFF
        return LADIFF
    if name == "MldTextEncoder":
        from .text_encoder import MldTextEncoder
        return MldTextEncoder
    if name == "default_config":
        from .defaults import default_config
        return default_config
    raise AttributeError(name)
