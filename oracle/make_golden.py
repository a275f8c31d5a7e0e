"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules
(imported from /root/reference/src, authoring container only) on seeded
synthetic weights + inputs.  TEST INFRASTRUCTURE.

    python -m oracle.make_golden

The weights are regenerated anywhere from ``make_state_dict(seed)`` (torch CPU
generator, deterministic), so only inputs' seeds and the reference OUTPUTS are
stored.  The loop glue (LADIFF._diffusion_reverse LAD branch + DDIM) comes from
the restatement in ladiff_oracle.py because LADIFF / diffusers cannot be
imported (see ref_import.py); all learned arithmetic is the reference's.
"""
import os

import numpy as np
import torch

from . import ladiff_oracle as O
from . import ref_import as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    assert R.available(), "needs /root/reference"
    torch.set_grad_enabled(False)
    os.makedirs(OUT, exist_ok=True)
    seed = 1234
    sd = O.make_state_dict(seed, 263, perturb=True)
    den, vae = R.build_reference(263)
    R.load_synthetic(den, vae, sd)

    def den_fn(x, t, ehs, mie):
        return den(sample=x, timestep=t, encoder_hidden_states=ehs, lengths=None, max_iter_elements=mie)[0]

    # ---- 1. one denoiser call with per-layer traces (forward hooks on the reference blocks)
    lengths = [196, 40, 100]
    B = len(lengths)
    g = torch.Generator().manual_seed(7)
    text = torch.randn((2 * B, 1, 768), generator=g)
    x = O.initial_latents(torch.randn((B, 5, 256), generator=g), lengths)
    x2 = torch.cat([x] * 2)
    mie2 = torch.cat([O.max_iter_elements_of(lengths)] * 2)
    traces, hooks = {}, []
    enc = den.encoder
    blocks = ([(f"input_blocks.{i}", m) for i, m in enumerate(enc.input_blocks)] + [("middle_block", enc.middle_block)]
              + [(f"output_blocks.{i}", m) for i, m in enumerate(enc.output_blocks)])
    for name, mod in blocks:
        hooks.append(mod.register_forward_hook(
            lambda m, i, o, name=name: traces.__setitem__(name, o.permute(1, 0, 2).contiguous().numpy().copy())))
    out = den_fn(x2, torch.tensor(981), text, mie2)
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(OUT, "denoiser_step.npz"), weight_seed=seed, input_seed=7, lengths=lengths,
                        timestep=981, out=out.numpy(), **{"layer." + k: v for k, v in traces.items()})

    # ---- 2. full sampling: 50 steps + decode, B=4 ragged
    lengths = [196, 40, 100, 148]
    text, noise, _ = O.synthetic_inputs(len(lengths), seed=11)
    rec = {"steps": (1, 10, 50)}
    z = O.diffusion_reverse(sd, text, lengths, noise, 50, 7.5, record=rec, denoiser_fn=den_fn)
    feats = vae.decode(z, lengths)
    z20 = O.diffusion_reverse(sd, text, lengths, noise, 20, 7.5, denoiser_fn=den_fn)
    np.savez_compressed(os.path.join(OUT, "sampling.npz"), weight_seed=seed, input_seed=11, lengths=lengths,
                        lat1=rec["latents_after_1"].numpy(), lat10=rec["latents_after_10"].numpy(),
                        lat50=rec["latents_after_50"].numpy(), z=z.numpy(), z20=z20.numpy(),
                        feats=feats.numpy().astype(np.float32))

    # ---- 3. decode only, ragged incl. a length that is not a multiple of 4 and the minimum m=1
    lengths = [196, 44, 96, 145, 57]
    g = torch.Generator().manual_seed(13)
    zin = O.initial_latents(torch.randn((len(lengths), 5, 256), generator=g), lengths).permute(1, 0, 2).contiguous()
    feats = vae.decode(zin, lengths)
    np.savez_compressed(os.path.join(OUT, "decode.npz"), weight_seed=seed, input_seed=13, lengths=lengths,
                        feats=feats.numpy())

    # ---- 4. KIT-ML (nfeats 251): only final_layer / skel_embedding change
    sdk = O.make_state_dict(seed, 251, perturb=True)
    denk, vaek = R.build_reference(251)
    R.load_synthetic(denk, vaek, sdk)
    lengths = [120, 196]
    g = torch.Generator().manual_seed(17)
    zin = O.initial_latents(torch.randn((2, 5, 256), generator=g), lengths).permute(1, 0, 2).contiguous()
    feats = vaek.decode(zin, lengths)
    np.savez_compressed(os.path.join(OUT, "decode_kit.npz"), weight_seed=seed, input_seed=17, lengths=lengths,
                        feats=feats.numpy())
    # ---- 5. LADiffVae.encode: the deterministic part of the reference output (dist.loc / dist.scale) on ragged motions
    lengths = [196, 44, 96, 145, 57]
    g = torch.Generator().manual_seed(19)
    motion = 0.5 * torch.randn((len(lengths), max(lengths), 263), generator=g)
    for i, L in enumerate(lengths):
        motion[i, L:] = 0
    torch.manual_seed(0)
    lat, dist, mie = vae.encode(motion, lengths)
    np.savez_compressed(os.path.join(OUT, "encode.npz"), weight_seed=seed, input_seed=19, lengths=lengths,
                        mu=dist.loc.numpy(), std=dist.scale.numpy(), mie=mie.numpy())

    # ---- 6. one denoiser call on the ARDIFF branch (enclat conditioning, no key-padding mask: ladiff_denoiser.py:218-255)
    g = torch.Generator().manual_seed(23)
    Bq = 3
    text = torch.randn((2 * Bq, 1, 768), generator=g)
    xs = torch.randn((2 * Bq, 1, 256), generator=g)
    enclat = torch.randn((2 * Bq, 2, 256), generator=g)
    out = den(sample=xs, timestep=torch.tensor(441), encoder_hidden_states=text, enclat=enclat, lengths=None)[0]
    np.savez_compressed(os.path.join(OUT, "denoiser_step_ardiff.npz"), weight_seed=seed, input_seed=23, timestep=441,
                        out=out.numpy())

    # ---- 7. recover_from_ric (data/humanml/scripts/motion_process.py:415-430), the reference's own function
    import sys
    sys.path.insert(0, R.REF_SRC)
    from ladiff.data.humanml.scripts.motion_process import recover_from_ric
    g = torch.Generator().manual_seed(29)
    data = 0.3 * torch.randn((3, 40, 263), generator=g)
    np.savez_compressed(os.path.join(OUT, "recover_from_ric.npz"), input_seed=29,
                        joints22=recover_from_ric(data, 22).numpy(),
                        joints21=recover_from_ric(data[..., :251], 21).numpy())

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
