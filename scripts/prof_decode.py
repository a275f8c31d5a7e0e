"""In-situ (warm L2, graph replay) per-kernel device times of one decode / encode / reverse loop via torch.profiler (CUPTI):
    python scripts/prof_decode.py [mode] [B] [what=decode|encode|reverse]"""
import os, sys, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import ladiff_b200 as L
from ladiff_b200.data import SyntheticDataModule
from ladiff_b200.modeltype import LADIFF

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
what = sys.argv[3] if len(sys.argv) > 3 else "decode"
torch.set_grad_enabled(False)
cfg = L.default_config("humanml3d", num_inference_timesteps=50 if what == "reverse" else 2)
torch.manual_seed(1234)
model = LADIFF(cfg, SyntheticDataModule(263, 22))
model.text_encoder = None
model = model.cuda().eval()
model.set_precision(mode)
g = torch.Generator().manual_seed(0)
text = torch.randn((2 * B, 1, 768), generator=g).cuda()
noise = torch.randn((B, 5, 256), generator=g).cuda()
lengths = [196] * B
z = model._diffusion_reverse(text, lengths, latents=noise)
feats = model.vae.decode(z, lengths)
fn = {"decode": lambda: model.vae.decode(z, lengths), "encode": lambda: model.vae.encode(feats, lengths),
      "reverse": lambda: model._diffusion_reverse(text, lengths, latents=noise)}[what]
for _ in range(3):
    fn()
torch.cuda.synchronize()
reps = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0:
        name = re.sub(r"\(.*", "", e.name).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"{mode} B={B} {what}: {tot / reps / 1e3:.3f} ms of kernel time per call (sum over kernels; in-situ, warm L2)")
print(f"{'share':>7} {'us/call':>9} {'launches/call':>13} {'avg us':>8}  kernel")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:18]:
    print(f"{100 * v[1] / tot:6.1f}% {v[1] / reps:9.1f} {v[0] / reps:13.1f} {v[1] / v[0]:8.2f}  {k[:90]}")
