"""One (or a few) full sampling steps at B=128 / 196 frames for ncu launch lists and A/B timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ladiff_b200 as L
from ladiff_b200.data import SyntheticDataModule
from ladiff_b200.modeltype import LADIFF

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
B = int(sys.argv[4]) if len(sys.argv) > 4 else 128
torch.set_grad_enabled(False)
cfg = L.default_config("humanml3d", num_inference_timesteps=n_steps)
cfg.model.clip_path = "synthetic://clip-vit-large-patch14"
torch.manual_seed(1234)
model = LADIFF(cfg, SyntheticDataModule(263, 22))
model.text_encoder = None
model = model.cuda().eval()
model.set_precision(mode)
g = torch.Generator().manual_seed(0)
text = torch.randn((2 * B, 1, 768), generator=g).cuda()
noise = torch.randn((B, 5, 256), generator=g).cuda()
lengths = [196] * B
def run():
    z = model._diffusion_reverse(text, lengths, latents=noise)
    return model.vae.decode(z, lengths), z
f, z = run(); torch.cuda.synchronize()
for name, fn in (("reverse", lambda: model._diffusion_reverse(text, lengths, latents=noise)), ("decode", lambda: model.vae.decode(z, lengths))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{mode} B={B} steps={n_steps} {name}: {e0.elapsed_time(e1)/reps:.3f} ms  launches {model._bind().last_launch_count}")
