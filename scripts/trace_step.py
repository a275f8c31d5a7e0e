"""Per-launch timeline of one reverse-loop replay (LADIFF_TRACE=1): python scripts/trace_step.py [mode] [steps] [B]"""
import os, sys
os.environ["LADIFF_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ladiff_b200 as L
from ladiff_b200.data import SyntheticDataModule
from ladiff_b200.modeltype import LADIFF

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.set_grad_enabled(False)
cfg = L.default_config("humanml3d", num_inference_timesteps=n_steps)
torch.manual_seed(1234)
model = LADIFF(cfg, SyntheticDataModule(263, 22))
model.text_encoder = None
model = model.cuda().eval()
model.set_precision(mode)
g = torch.Generator().manual_seed(0)
text = torch.randn((2 * B, 1, 768), generator=g).cuda()
noise = torch.randn((B, 5, 256), generator=g).cuda()
lengths = [196] * B
eng = model._bind()
for _ in range(3):
    model._diffusion_reverse(text, lengths, latents=noise)
torch.cuda.synchronize()
eng.trace_read()          # clear
model._diffusion_reverse(text, lengths, latents=noise)
tr = [t for t in eng.trace_read(extra=True) if t[1] < (1 << 63)]
tr.sort(key=lambda t: t[1])
# steady-state: launches of the middle step
per_step = [t for t in tr if "M%d " % (2 * B * 5) in t[0]]
n = len(per_step) // n_steps
mid = per_step[(n_steps // 2) * n:(n_steps // 2 + 1) * n]
print(f"{len(tr)} traced launches, {n} per step; step {n_steps // 2}:")
print(f"{'kernel':28s} {'start':>8s} {'dep-wait':>9s} {'accum':>8s} {'done':>8s} | {'dur':>6s} {'gap-to-next-start':>8s} {'prev-done->wait':>8s}")
t0 = mid[0][1]
prev_done = None
for i, (nm, s, w, a, d, *xs) in enumerate(mid):
    nxt = mid[i + 1][1] if i + 1 < len(mid) else None
    print(f"{nm:28s} {(s - t0) / 1e3:8.2f} {(w - t0) / 1e3:9.2f} {(a - t0) / 1e3:8.2f} {(d - t0) / 1e3:8.2f} | {(d - s) / 1e3:6.2f} "
          f"{((nxt - d) / 1e3 if nxt else 0):8.2f} {((w - prev_done) / 1e3 if prev_done else 0):8.2f}")
    if any(0 < x < (1 << 63) for x in xs):
        print("    kernel stamps 4..7 (us after dep-wait): " + " ".join(f"{(x - w) / 1e3:7.2f}" if 0 < x < (1 << 63) else "      -" for x in xs))
    prev_done = d
print(f"step span {(mid[-1][4] - mid[0][1]) / 1e3:.1f} us")
