"""Import the UNMODIFIED reference modules from /root/reference (authoring container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing
that runs there may import this file; it is used by ``oracle/make_golden.py``
(to generate ``tests/golden/*``) and by the ``reference``-marked CPU tests,
which skip when the tree is absent.

``mdiff_transformer.py:10`` does ``import clip`` at import time (never used on
the path) -> a stub module is registered.  ``LADIFF`` itself cannot be imported
(pytorch_lightning / diffusers / torchmetrics / omegaconf are absent), so the
loop glue is taken from ``oracle/ladiff_oracle.py`` and only the learned
arithmetic (denoiser, VAE) comes from the reference classes.
"""
import os
import sys
import types
from types import SimpleNamespace

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "ladiff"))


def ablation():
    """configs/config_ladiff_humanml3d.yaml:50-64 (+ VAE_TYPE / MLP_DIST from configs/base.yaml)."""
    return SimpleNamespace(SKIP_CONNECT=True, PE_TYPE="mld", DIFF_PE_TYPE="mld", IDEA="ard", DVAE=False,
                           PERCENTAGE_NOISED=0.0, FINETUNE_DECODER=False, MAX_IT=5, FRAME_PER_LATENT=48,
                           MD_TRANS=True, PE_LATENT=False, JOINT_DISTRO_FIX=False, LAD=True,
                           TEST_EFFICIENCY=False, VAE_TYPE="ladiff", MLP_DIST=False)


def build_reference(nfeats: int = 263):
    """Returns (denoiser, vae): reference classes built with configs/modules/{denoiser,motion_vae}.yaml params."""
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    sys.modules.setdefault("clip", types.ModuleType("clip"))
    from ladiff.models.architectures.ladiff_denoiser import LADiffDenoiser
    from ladiff.models.architectures.ladiff_vae import LADiffVae
    den = LADiffDenoiser(ablation=ablation(), nfeats=nfeats, condition="text", latent_dim=[7, 256], ff_size=1024,
                         num_layers=9, num_heads=4, dropout=0.1, normalize_before=False, activation="gelu",
                         flip_sin_to_cos=True, return_intermediate_dec=False, position_embedding="learned",
                         arch="trans_enc", freq_shift=0, guidance_scale=7.5, guidance_uncondp=0.1,
                         text_encoded_dim=768, nclasses=10)
    vae = LADiffVae(ablation=ablation(), nfeats=nfeats, latent_dim=[7, 256], ff_size=1024, num_layers=9,
                    num_heads=4, dropout=0.1, arch="encoder_decoder", normalize_before=False,
                    activation="gelu", position_embedding="learned")
    return den.eval(), vae.eval()


def load_synthetic(den, vae, sd):
    """strict=True: proves the synthetic key layout equals the reference's."""
    from .ladiff_oracle import sub
    den.load_state_dict(sub(sd, "denoiser."), strict=True)
    vae.load_state_dict(sub(sd, "vae."), strict=True)
