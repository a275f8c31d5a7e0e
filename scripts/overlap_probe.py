"""Does decoding batch i on a low-priority side stream while batch i+1 runs its reverse loop raise throughput?
python scripts/overlap_probe.py [mode] [B] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ladiff_b200 as L
from ladiff_b200.data import SyntheticDataModule
from ladiff_b200.modeltype import LADIFF

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
torch.set_grad_enabled(False)
cfg = L.default_config("humanml3d", num_inference_timesteps=50)
torch.manual_seed(1234)
model = LADIFF(cfg, SyntheticDataModule(263, 22))
model.text_encoder = None
model = model.cuda().eval()
model.set_precision(mode)
g = torch.Generator().manual_seed(0)
text = torch.randn((2 * B, 1, 768), generator=g).cuda()
noise = torch.randn((B, 5, 256), generator=g).cuda()
lengths = [196] * B
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
main = torch.cuda.current_stream()

def sequential():
    outs = []
    for _ in range(K):
        z = model._diffusion_reverse(text, lengths, latents=noise)
        outs.append(model.vae.decode(z, lengths))
    return outs

def overlapped(side, rev_stream):
    outs = []
    evs = []
    with torch.cuda.stream(rev_stream):
        pass
    for _ in range(K):
        with torch.cuda.stream(rev_stream):
            z = model._diffusion_reverse(text, lengths, latents=noise)
            ev = torch.cuda.Event(); ev.record(rev_stream)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            z.record_stream(side)
            outs.append(model.vae.decode(z, lengths))
    main.wait_stream(side); main.wait_stream(rev_stream)
    return outs

def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); outs = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, outs

t_seq, o_seq = timeit(sequential)
print(f"{mode} B={B}: sequential {t_seq:.3f} ms/step -> {B / t_seq * 1e3:.0f} seq/s")
for name, (ps, pr) in {"side low / reverse high": (0, -1), "both default": (0, 0)}.items():
    side = torch.cuda.Stream(priority=ps)
    rev = torch.cuda.Stream(priority=pr)
    rev.wait_stream(main); side.wait_stream(main)
    t_ov, o_ov = timeit(lambda: overlapped(side, rev))
    same = all(torch.equal(a, b) for a, b in zip(o_seq, o_ov))
    print(f"{mode} B={B}: overlapped ({name}) {t_ov:.3f} ms/step -> {B / t_ov * 1e3:.0f} seq/s; outputs identical to sequential: {same}")
