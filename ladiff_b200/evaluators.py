"""The three T2M evaluator networks ``LADIFF.t2m_eval`` feeds (reference ``models/modeltype/ladiff.py:179-219`` builds them,
``:1261-1270`` calls them; architectures ``models/architectures/t2m_motionenc.py:6-64`` and ``t2m_textenc.py:6-48``).

They sit AFTER the hot path (a strided-conv movement encoder and two bi-GRU co-embedding heads over <= 49 movement steps /
<= 20 words) and stay plain torch, with the reference's parameter names so ``finest.tar`` checkpoints load with
``strict=True``; without a checkpoint (none exists offline) they are random-init and ``t2m_eval`` still returns every key
with the right shapes, which is what the evaluation loop and the metric modules consume."""
from __future__ import annotations

import os

import torch
from torch import nn
from torch.nn.utils.rnn import pack_padded_sequence


class MovementConvEncoder(nn.Module):
    """[B, L, nfeats-4] -> [B, L/4, out]: two stride-2 convolutions (keys main.0 / main.3 / out_net)."""

    def __init__(self, input_size: int, hidden_size: int, output_size: int):
        super().__init__()
        self.main = nn.Sequential(nn.Conv1d(input_size, hidden_size, 4, 2, 1), nn.Dropout(0.2, inplace=True),
                                  nn.LeakyReLU(0.2, inplace=True), nn.Conv1d(hidden_size, output_size, 4, 2, 1),
                                  nn.Dropout(0.2, inplace=True), nn.LeakyReLU(0.2, inplace=True))
        self.out_net = nn.Linear(output_size, output_size)

    def forward(self, inputs):
        return self.out_net(self.main(inputs.transpose(1, 2)).transpose(1, 2))


class _BiGRUHead(nn.Module):
    """input_emb -> bi-GRU (learned initial state ``hidden``) -> output_net on the concatenated last states."""

    def __init__(self, input_size: int, hidden_size: int, output_size: int):
        super().__init__()
        self.input_emb = nn.Linear(input_size, hidden_size)
        self.gru = nn.GRU(hidden_size, hidden_size, batch_first=True, bidirectional=True)
        self.output_net = nn.Sequential(nn.Linear(hidden_size * 2, hidden_size), nn.LayerNorm(hidden_size),
                                        nn.LeakyReLU(0.2, inplace=True), nn.Linear(hidden_size, output_size))
        self.hidden_size = hidden_size
        self.hidden = nn.Parameter(torch.randn((2, 1, hidden_size)))

    def _run(self, x, lens):
        packed = pack_padded_sequence(self.input_emb(x), lens.data.tolist(), batch_first=True)   # lengths sorted descending
        _, last = self.gru(packed, self.hidden.repeat(1, x.shape[0], 1))
        return self.output_net(torch.cat([last[0], last[1]], dim=-1))


class MotionEncoderBiGRUCo(_BiGRUHead):
    def forward(self, inputs, m_lens):
        return self._run(inputs, m_lens)


class TextEncoderBiGRUCo(_BiGRUHead):
    def __init__(self, word_size: int, pos_size: int, hidden_size: int, output_size: int):
        super().__init__(word_size, hidden_size, output_size)
        self.pos_emb = nn.Linear(pos_size, word_size)

    def forward(self, word_embs, pos_onehot, cap_lens):
        return self._run(word_embs + self.pos_emb(pos_onehot), cap_lens)


def build_t2m_evaluators(cfg, nfeats: int):
    """``LADIFF._get_t2m_evaluator`` (ladiff.py:179-219): dims from ``cfg.model.t2m_*`` (configs/base.yaml:50-60); loads
    ``<t2m_path>/<dataset>/text_mot_match/model/finest.tar`` when it exists, frozen either way."""
    te, me = cfg.model.t2m_textencoder, cfg.model.t2m_motionencoder
    text = TextEncoderBiGRUCo(te.dim_word, te.dim_pos_ohot, te.dim_text_hidden, te.dim_coemb_hidden)
    move = MovementConvEncoder(nfeats - 4, me.dim_move_hidden, me.dim_move_latent)
    motion = MotionEncoderBiGRUCo(me.dim_move_latent, me.dim_motion_hidden, me.dim_motion_latent)
    root = cfg.model.get("t2m_path", None)
    if root:
        name = cfg.get("TEST", {}).get("DATASETS", ["humanml3d"])[0] if "TEST" in cfg else "humanml3d"
        path = os.path.join(root, "t2m" if name == "humanml3d" else name, "text_mot_match/model/finest.tar")
        if os.path.exists(path):
            ck = torch.load(path, map_location="cpu")
            text.load_state_dict(ck["text_encoder"])
            move.load_state_dict(ck["movement_encoder"])
            motion.load_state_dict(ck["motion_encoder"])
    for m in (text, move, motion):
        m.eval()
        for p in m.parameters():
            p.requires_grad = False
    return text, move, motion
