"""Runs the memory-bound kernels of the path once at the headline shapes (B = 128, 196 frames) for an ncu capture:
    LADIFF_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
        --clock-control none -k regex:"k_cfg_ddim|k_cross_ln|k_layernorm256|k_feats2joints|k_attn_ln|k_enc_|k_dec_init|k_pack_x" \
        --csv --log-file gpurun_out/hbm.csv python scripts/prof_hbm.py
    python scripts/summarize_hbm.py gpurun_out/hbm.csv"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ladiff_b200 as L
from ladiff_b200.data import SyntheticDataModule
from ladiff_b200.modeltype import LADIFF

torch.set_grad_enabled(False)
B = 128
cfg = L.default_config("humanml3d", num_inference_timesteps=2)
torch.manual_seed(1234)
model = LADIFF(cfg, SyntheticDataModule(263, 22))
model.text_encoder = None
model = model.cuda().eval()
g = torch.Generator().manual_seed(0)
text = torch.randn((2 * B, 1, 768), generator=g).cuda()
noise = torch.randn((B, 5, 256), generator=g).cuda()
lengths = [196] * B
for _ in range(2):
    z = model._diffusion_reverse(text, lengths, latents=noise)
    feats = model.vae.decode(z, lengths)
    joints = model.datamodule.feats2joints(feats)
    lat, dist, _ = model.vae.encode(feats, lengths)
torch.cuda.synchronize()
print("ok", tuple(feats.shape), tuple(joints.shape), tuple(lat.shape))
