"""Drop-in for ``ladiff.models.architectures.ladiff_vae.LADiffVae`` (reference file lines 33-362).

``decode(z, lengths)`` -- the hot path -- runs in the sm_100a kernels behind ``ladiff_vae_decode``: ragged (only the
L_i valid frames and m_i valid latents of each sequence are computed), padded frames come back as exact zeros.
``encode`` is outside the sampling path (training / reconstruction) and delegates to torch with the same parameters.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor, nn

from ._params import DecoderLayerParams, EngineBound, LearnedPE1D, PostNormEncoderLayerParams, SkipStack
from .denoiser import _abl
from .utils import lengths_to_mask


class LADiffVae(EngineBound):
    _prefix = "vae."
    _which = 2

    def __init__(self,
                 ablation,
                 nfeats: int,
                 latent_dim: list = [1, 256],
                 ff_size: int = 1024,
                 num_layers: int = 9,
                 num_heads: int = 4,
                 dropout: float = 0.1,
                 arch: str = "all_encoder",
                 normalize_before: bool = False,
                 activation: str = "gelu",
                 position_embedding: str = "learned",
                 precision: str = "bf16x3",
                 max_frames: int = 196,
                 **kwargs) -> None:
        super().__init__()
        self.latent_size = latent_dim[0]
        self.latent_dim = latent_dim[-1]
        self.nfeats = nfeats
        self.arch = arch
        self.mlp_dist = _abl(ablation, "MLP_DIST", False)
        self.pe_type = _abl(ablation, "PE_TYPE")
        self.dvae = _abl(ablation, "DVAE", False)
        self.max_it = int(_abl(ablation, "MAX_IT", 5))
        self.frame_per_latent = int(_abl(ablation, "FRAME_PER_LATENT", 48))
        self.joint_distro_fix = _abl(ablation, "JOINT_DISTRO_FIX", False)
        self.LAD = _abl(ablation, "LAD", True)
        self.test_efficiency = _abl(ablation, "TEST_EFFICIENCY", False)

        if self.pe_type != "mld":
            if self.pe_type == "actor":
                raise NotImplementedError("ladiff_b200 implements PE_TYPE 'mld' only")
            raise ValueError("Not Support PE type")
        if self.arch != "encoder_decoder":
            if self.arch == "all_encoder":
                raise NotImplementedError("ladiff_b200 implements arch 'encoder_decoder' only (configs/modules/motion_vae.yaml:4)")
            raise ValueError("Not support architecture!")
        if position_embedding not in ("v3", "learned"):
            raise ValueError(f"not supported {position_embedding}")
        if (self.latent_dim, num_layers, num_heads, ff_size, activation) != (256, 9, 4, 1024, "gelu") or normalize_before \
                or self.max_it == 0 or self.mlp_dist or self.test_efficiency or not self.LAD:
            raise NotImplementedError("the sm_100a kernels are specialised for latent 256, 9 layers, 4 heads, ff 1024, gelu, "
                                      "post-norm, MAX_IT>0, LAD=True (configs/modules/motion_vae.yaml, config_ladiff_humanml3d.yaml:50-64)")

        self.query_pos_encoder = LearnedPE1D(self.latent_dim)
        self.query_pos_decoder = LearnedPE1D(self.latent_dim)
        self.encoder = SkipStack(lambda: PostNormEncoderLayerParams(self.latent_dim, num_heads, ff_size, dropout, activation),
                                 num_layers, self.latent_dim)
        self.decoder = SkipStack(lambda: DecoderLayerParams(self.latent_dim, num_heads, ff_size, dropout),
                                 num_layers, self.latent_dim)
        self.global_motion_token = nn.Parameter(torch.randn(self.max_it * 2, self.latent_dim))
        self.skel_embedding = nn.Linear(nfeats, self.latent_dim)
        self.final_layer = nn.Linear(self.latent_dim, nfeats)
        self.max_frames = int(max_frames)
        self._init_engine_state(precision, dict(nfeats=nfeats, max_it=self.max_it, frame_per_latent=self.frame_per_latent,
                                                max_frames=max_frames))

    def forward(self, features: Tensor, lengths: Optional[List[int]] = None):
        z, dist, _ = self.encode(features, lengths)
        return self.decode(z, lengths), z, dist

    def dist_to_mask(self, max_iter_elements, z):
        """reference :152-159"""
        masks = torch.ones((len(max_iter_elements), z.shape[0]), dtype=torch.bool, device=z.device)
        for i, e in enumerate(max_iter_elements):
            masks[i, int(e):] = False
        return masks

    @torch.no_grad()
    def decode(self, z: Tensor, lengths: List[int], plot_att_map=None, latentwise_gen=None, max_len: Optional[int] = None):
        """z [MAX_IT, B, 256], lengths List[int] -> [B, max(lengths), nfeats]  (reference :288-362).
        ``max_len`` (extension): pad to this many frames instead of ``max(lengths)`` -- sharded callers need one shape on
        every rank before the all-gather (``parallel.gather_motions``)."""
        if plot_att_map is not None or latentwise_gen is not None:
            raise NotImplementedError("attention-map plotting / latent-wise generation are analysis tools outside the sampling path")
        lengths = [int(x) for x in lengths]
        out = self.engine().vae_decode(z, lengths, self.mode, max_len=max_len)
        return out.to(z.dtype)

    def encode(self, features: Tensor, lengths: Optional[List[int]] = None, eps: Optional[Tensor] = None,
               generator: Optional[torch.Generator] = None):
        """features [B, max(lengths), nfeats] -> (latent [MAX_IT,B,256], dist, max_iter_elements)  (reference :162-286, LAD
        branch, JOINT_DISTRO_FIX False).  CUDA tensors run in the sm_100a kernels behind ``ladiff_vae_encode`` (ragged
        token sequence mu | logvar | frames through the non-MD skip encoder); ``dist`` is the same
        ``torch.distributions.Normal(mu, std)`` on the valid rows (masked rows, whose values nothing reads in the reference,
        are N(0, 1) here) and ``latent = mu + std * eps`` with ``eps`` injected or drawn from ``generator``."""
        if self.joint_distro_fix or self.dvae:
            raise NotImplementedError("encode supports JOINT_DISTRO_FIX=False, DVAE=False")
        if lengths is None:
            lengths = [len(f) for f in features]
        lengths = [int(x) for x in lengths]
        mie = torch.ceil(torch.tensor(lengths) / self.frame_per_latent).to(torch.long)      # :198
        if not features.is_cuda:
            return self._encode_torch(features, lengths, mie, eps)
        if eps is None:
            eps = torch.randn((self.max_it, len(lengths), self.latent_dim), device=features.device, dtype=torch.float,
                              generator=generator)
        latent, mu, std = self.engine().vae_encode(features, lengths, self.mode, eps)
        return latent.to(features.dtype), torch.distributions.Normal(mu, std), mie

    def _encode_torch(self, features, lengths, mie, eps=None):
        """Plain-torch statement of the same computation for CPU tensors (host-side tests of the parameter holders; the CUDA
        path above is the product)."""
        device = features.device
        bs = features.shape[0]
        mask = lengths_to_mask(lengths, device)
        x = self.skel_embedding(features).permute(1, 0, 2)
        dist = torch.tile(self.global_motion_token[:, None, :], (1, bs, 1))
        dm = torch.ones((bs, self.max_it), dtype=torch.bool, device=device)
        for i, e in enumerate(mie):
            dm[i, int(e):] = False
        aug_mask = torch.cat((dm, dm, mask), 1)
        xseq = torch.cat((dist, x), 0)
        xseq = xseq + self.query_pos_encoder.pe[:xseq.shape[0]]
        dist = self.encoder.forward_encoder(xseq, src_key_padding_mask=~aug_mask)[:dist.shape[0]]
        mu, logvar = dist[:self.max_it], dist[self.max_it:]
        std = logvar.exp().pow(0.5)
        d = torch.distributions.Normal(mu, std)
        latent = d.rsample() if eps is None else mu + std * eps
        for i, e in enumerate(mie):
            latent[int(e):, i] = 0
        return latent, d, mie
