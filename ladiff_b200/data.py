"""Minimal stand-in for the one datamodule call on the sampling path: ``datamodule.feats2joints``
(``models/modeltype/ladiff.py:149,307`` -> ``data/HumanML3D.py:44-48``).  The reference de-normalises and runs
``recover_from_ric`` on the CPU after a ``.cpu()``; here it is one CUDA kernel (``ladiff_feats2joints``)."""
from __future__ import annotations

import torch


class SyntheticDataModule:
    """Holds (mean, std, nfeats, njoints) like ``HumanML3DDataModule.hparams``; no dataset on disk is needed."""
    accepts_cuda = True

    def __init__(self, nfeats: int = 263, njoints: int = 22, mean=None, std=None, engine=None, mean_eval=None, std_eval=None):
        self.nfeats, self.njoints = nfeats, njoints
        self.mean = torch.zeros(nfeats) if mean is None else torch.as_tensor(mean, dtype=torch.float32)
        self.std = torch.ones(nfeats) if std is None else torch.as_tensor(std, dtype=torch.float32)
        self.mean_eval = self.mean if mean_eval is None else torch.as_tensor(mean_eval, dtype=torch.float32)
        self.std_eval = self.std if std_eval is None else torch.as_tensor(std_eval, dtype=torch.float32)
        self.is_mm = False
        self._engine = engine

    def renorm4t2m(self, features: torch.Tensor) -> torch.Tensor:
        """data/HumanML3D.py:57-65: de-normalise with the dataset statistics, re-normalise with the T2M evaluators' ones."""
        f = features * self.std.to(features) + self.mean.to(features)
        return (f - self.mean_eval.to(features)) / self.std_eval.to(features)

    def bind_engine(self, engine):
        self._engine = engine

    def feats2joints(self, features: torch.Tensor) -> torch.Tensor:
        if not features.is_cuda:
            raise RuntimeError("feats2joints runs on the GPU (ladiff_feats2joints); pass a CUDA tensor")
        if self._engine is None:
            from ._lib import Engine
            self._engine = Engine(nfeats=self.nfeats)
        return self._engine.feats2joints(features, self.mean.to(features.device), self.std.to(features.device), self.njoints)
