"""Two reverse loops in flight (two engines = two workspaces) + low-priority decodes: ms per batch of 128.
python scripts/overlap_probe2.py [mode] [B] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ladiff_b200 as L
from ladiff_b200.data import SyntheticDataModule
from ladiff_b200.modeltype import LADIFF

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 12
torch.set_grad_enabled(False)
cfg = L.default_config("humanml3d", num_inference_timesteps=50)
models = []
for _ in range(2):
    torch.manual_seed(1234)
    m = LADIFF(cfg, SyntheticDataModule(263, 22))
    m.text_encoder = None
    m = m.cuda().eval()
    m.set_precision(mode)
    models.append(m)
g = torch.Generator().manual_seed(0)
text = torch.randn((2 * B, 1, 768), generator=g).cuda()
noise = torch.randn((B, 5, 256), generator=g).cuda()
lengths = [196] * B
main = torch.cuda.current_stream()

def sequential():
    return [models[0].vae.decode(models[0]._diffusion_reverse(text, lengths, latents=noise), lengths) for _ in range(K)]

def pipelined():
    return list(models[0].sample_stream((text, lengths, noise) for _ in range(K)))

def two_lanes(dec_low=True):
    revs = [torch.cuda.Stream(priority=-1) for _ in range(2)]
    decs = [torch.cuda.Stream(priority=0 if dec_low else -1) for _ in range(2)]
    for s in revs + decs:
        s.wait_stream(main)
    outs = []
    for i in range(K):
        lane = i & 1
        m = models[lane]
        with torch.cuda.stream(revs[lane]):
            z = m._diffusion_reverse(text, lengths, latents=noise)
            ev = torch.cuda.Event(); ev.record(revs[lane])
        decs[lane].wait_event(ev)
        with torch.cuda.stream(decs[lane]):
            z.record_stream(decs[lane])
            outs.append(m.vae.decode(z, lengths))
    for s in revs + decs:
        main.wait_stream(s)
    return outs

def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); outs = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K, outs

t, o0 = timeit(sequential); print(f"{mode} B={B}: sequential          {t:.3f} ms/batch -> {B / t * 1e3:.0f} seq/s")
t, o1 = timeit(pipelined); print(f"{mode} B={B}: sample_stream       {t:.3f} ms/batch -> {B / t * 1e3:.0f} seq/s  identical {all(torch.equal(a, b) for a, b in zip(o0, o1))}")
t, o2 = timeit(two_lanes); print(f"{mode} B={B}: two reverse lanes   {t:.3f} ms/batch -> {B / t * 1e3:.0f} seq/s  identical {all(torch.equal(a, b) for a, b in zip(o0, o2))}")
