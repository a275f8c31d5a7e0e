"""Small host helpers with the reference's names and semantics (``src/ladiff/utils/temos_utils.py:10-28``)."""
from __future__ import annotations

from typing import List, Sequence

import torch


def lengths_to_mask(lengths: Sequence[int], device: torch.device, max_len: int = None) -> torch.Tensor:
    """temos_utils.py:10-17"""
    lengths = torch.as_tensor(list(lengths), device=device)
    max_len = max_len if max_len else int(lengths.max())
    return torch.arange(max_len, device=device).expand(len(lengths), max_len) < lengths.unsqueeze(1)


def remove_padding(tensors, lengths) -> List[torch.Tensor]:
    """temos_utils.py:24-28"""
    return [tensor[:tensor_length] for tensor, tensor_length in zip(tensors, lengths)]


def max_iter_elements(lengths: Sequence[int], frame_per_latent: int) -> List[int]:
    """ceil(L / FRAME_PER_LATENT): models/modeltype/ladiff.py:379, architectures/ladiff_vae.py:292"""
    return [-(-int(L) // int(frame_per_latent)) for L in lengths]
