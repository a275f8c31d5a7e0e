"""Torch-CPU emulation of the *hoisted, ragged* dataflow the CUDA path implements
(DESIGN.md section 3): one python function per kernel / plan stage, same buffers, same
row packing.  TEST INFRASTRUCTURE: proves on CPU (no GPU needed) that the
algebraic hoists of SURVEY.md section 8a are exact w.r.t. the un-hoisted oracle, and
documents what each kernel must compute.  Never imported by ``ladiff_b200``.
"""
import math

import torch
import torch.nn.functional as F

from oracle import ladiff_oracle as O

D, H, HD = 256, 4, 64


def blocks():
    return O._block_names()


def ln(x, g, b):
    return F.layer_norm(x, (x.shape[-1],), g, b, 1e-5)


class DenoiserPlan:
    """Everything that does not depend on the latents (computed once per call)."""

    def __init__(self, sd, text768, lengths, n_steps, linear=F.linear):
        self.sd, self.lin = sd, linear
        P = lambda k: sd["denoiser." + k]
        B2 = text768.shape[0]
        self.B = B2 // 2
        self.m = O.max_iter_elements_of(lengths).tolist()
        m2 = self.m * 2
        # row packing: sequence s (0..2B-1; uncond first) owns rows off[s] .. off[s]+m[s]
        self.off = [0]
        for x in m2:
            self.off.append(self.off[-1] + x)
        self.R = self.off[-1]
        self.row_seq = torch.tensor([s for s, x in enumerate(m2) for _ in range(x)])
        self.row_t = torch.tensor([t for s, x in enumerate(m2) for t in range(x)])
        self.ts = O.ddim_timesteps(n_steps)
        acp = O.ddim_alphas_cumprod()
        self.c1, self.c2 = [], []
        for t in self.ts:
            p = int(t) - 1000 // n_steps
            a_t = acp[int(t)].item()
            a_p = acp[p].item() if p >= 0 else acp[0].item()
            self.c1.append(math.sqrt(a_p / a_t))
            self.c2.append(math.sqrt(1 - a_p) - math.sqrt(a_p) * math.sqrt(1 - a_t) / math.sqrt(a_t))
        # ---- time side: tables over steps
        temb = O.timestep_embedding(torch.tensor(self.ts).float())
        t1 = F.silu(linear(temb, P("time_embedding.linear_1.weight"), P("time_embedding.linear_1.bias")))
        self.temb = linear(t1, P("time_embedding.linear_2.weight"), P("time_embedding.linear_2.bias"))  # [N,256]
        st = F.silu(self.temb)
        # ---- text side
        self.text = linear(F.relu(text768[:, 0]), P("emb_proj.1.weight"), P("emb_proj.1.bias"))           # [2B,256]
        self.timekv, self.textkv, self.mod_ca, self.mod_ffn, self.lny, self.delta = {}, {}, {}, {}, {}, {}
        for b in blocks():
            p = f"encoder.{b}."
            w, bi = P(p + "sa_block.self_attn.in_proj_weight"), P(p + "sa_block.self_attn.in_proj_bias")
            self.timekv[b] = linear(self.temb, w[D:], bi[D:])     # [N,512]  (k | v)
            self.textkv[b] = linear(self.text, w[D:], bi[D:])     # [2B,512]
            self.mod_ca[b] = linear(st, P(p + "ca_block.proj_out.emb_layers.1.weight"), P(p + "ca_block.proj_out.emb_layers.1.bias"))
            self.mod_ffn[b] = linear(st, P(p + "ffn.proj_out.emb_layers.1.weight"), P(p + "ffn.proj_out.emb_layers.1.bias"))
            tn = ln(self.text, P(p + "ca_block.text_norm.weight"), P(p + "ca_block.text_norm.bias"))
            y = linear(tn, P(p + "ca_block.value.weight"), P(p + "ca_block.value.bias"))       # N==1: softmax over text tokens == 1
            self.lny[b] = ln(y, P(p + "ca_block.proj_out.norm.weight"), P(p + "ca_block.proj_out.norm.bias"))
            # ca delta for every (step, seq): [N, 2B, 256]
            sc, sh = self.mod_ca[b][:, None, :D], self.mod_ca[b][:, None, D:]
            a = F.silu(self.lny[b][None] * (1 + sc) + sh)
            self.delta[b] = linear(a, P(p + "ca_block.proj_out.out_layers.2.weight"), P(p + "ca_block.proj_out.out_layers.2.bias"))


def k_attn_small(plan, qkv, textkv, timekv_step):
    """denoiser self-attention: per (seq, head), queries = m rows, keys = m rows + text + time."""
    out = torch.zeros(plan.R, D)
    scale = 1.0 / math.sqrt(HD)
    for s in range(2 * plan.B):
        r0, r1 = plan.off[s], plan.off[s + 1]
        q = qkv[r0:r1, :D]
        k = torch.cat([qkv[r0:r1, D:2 * D], textkv[s:s + 1, :D], timekv_step[None, :D]], 0)
        v = torch.cat([qkv[r0:r1, 2 * D:], textkv[s:s + 1, D:], timekv_step[None, D:]], 0)
        for h in range(H):
            sl = slice(h * HD, (h + 1) * HD)
            att = torch.softmax((q[:, sl] * scale) @ k[:, sl].T, dim=-1)
            out[r0:r1, sl] = att @ v[:, sl]
    return out


def denoiser_layer(plan, b, step, x):
    sd, lin = plan.sd, plan.lin
    P = lambda k: sd[f"denoiser.encoder.{b}." + k]
    qkv = lin(x, P("sa_block.self_attn.in_proj_weight"), P("sa_block.self_attn.in_proj_bias"))
    a = k_attn_small(plan, qkv, plan.textkv[b], plan.timekv[b][step])
    x1 = ln(x + lin(a, P("sa_block.self_attn.out_proj.weight"), P("sa_block.self_attn.out_proj.bias")),
            P("sa_block.norm1.weight"), P("sa_block.norm1.bias"))
    h = F.relu(lin(x1, P("sa_block.linear1.weight"), P("sa_block.linear1.bias")))
    x2 = ln(x1 + lin(h, P("sa_block.linear2.weight"), P("sa_block.linear2.bias")),
            P("sa_block.norm2.weight"), P("sa_block.norm2.bias"))
    x3 = x2 + plan.delta[b][step][plan.row_seq]                       # ca_block, hoisted
    h = F.gelu(lin(x3, P("ffn.linear1.weight"), P("ffn.linear1.bias")))
    y = lin(h, P("ffn.linear2.weight"), P("ffn.linear2.bias"))
    sc, sh = plan.mod_ffn[b][step][:D], plan.mod_ffn[b][step][D:]
    s = F.silu(ln(y, P("ffn.proj_out.norm.weight"), P("ffn.proj_out.norm.bias")) * (1 + sc) + sh)
    return x3 + lin(s, P("ffn.proj_out.out_layers.2.weight"), P("ffn.proj_out.out_layers.2.bias"))


def denoiser_tokens(plan, step, x):
    """x [R,256] = latents + pe for both CFG halves -> tokens before the final LayerNorm."""
    sd, lin = plan.sd, plan.lin
    xs = []
    for i in range(4):
        x = denoiser_layer(plan, f"input_blocks.{i}", step, x)
        xs.append(x)
    x = denoiser_layer(plan, "middle_block", step, x)
    for i in range(4):
        w, bi = sd[f"denoiser.encoder.linear_blocks.{i}.weight"], sd[f"denoiser.encoder.linear_blocks.{i}.bias"]
        x = lin(x, w[:, :D]) + lin(xs.pop(), w[:, D:]) + bi          # two-source GEMM, no concat
        x = denoiser_layer(plan, f"output_blocks.{i}", step, x)
    return x


def diffusion_reverse(sd, text768, lengths, noise, n_steps=50, guidance=7.5, linear=F.linear):
    plan = DenoiserPlan(sd, text768, lengths, n_steps, linear)
    B = plan.B
    pe = sd["denoiser.query_pos.pe"][:, 0]
    g, bb = sd["denoiser.encoder.norm.weight"], sd["denoiser.encoder.norm.bias"]
    lat = O.initial_latents(noise, lengths)                          # [B,5,256]
    half = plan.R // 2
    seq, tt = plan.row_seq[:half], plan.row_t[:half]
    for step in range(n_steps):
        xh = lat[seq, tt] + pe[tt]
        tok = denoiser_tokens(plan, step, torch.cat([xh, xh], 0))
        eps_u, eps_c = ln(tok[:half], g, bb), ln(tok[half:], g, bb)
        eps = eps_u + guidance * (eps_c - eps_u)
        lat[seq, tt] = plan.c1[step] * lat[seq, tt] + plan.c2[step] * eps     # k_cfg_ddim
    return lat.permute(1, 0, 2).contiguous()                         # masked rows were never touched: exact zeros


# ---- decoder -----------------------------------------------------------------
def vae_decode(sd, z, lengths, linear=F.linear):
    lin = linear
    B = len(lengths)
    m = O.max_iter_elements_of(lengths).tolist()
    off = [0]
    for L in lengths:
        off.append(off[-1] + L)
    row_t = torch.tensor([t for L in lengths for t in range(L)])
    mem = [z[:m[b], b] for b in range(B)]                            # valid latent rows only
    x = sd["vae.query_pos_decoder.pe"][:, 0][row_t]
    scale = 1.0 / math.sqrt(HD)

    def mha(xq, kv_of, w_o, b_o):
        out = torch.zeros_like(xq)
        for b in range(B):
            r0, r1 = off[b], off[b + 1]
            k, v = kv_of(b)
            for h in range(H):
                sl = slice(h * HD, (h + 1) * HD)
                att = torch.softmax((xq[r0:r1, sl] * scale) @ k[:, sl].T, dim=-1)
                out[r0:r1, sl] = att @ v[:, sl]
        return lin(out, w_o, b_o)

    def layer(bn, x):
        P = lambda k: sd[f"vae.decoder.{bn}." + k]
        qkv = lin(x, P("self_attn.in_proj_weight"), P("self_attn.in_proj_bias"))
        a = mha(qkv[:, :D], lambda b: (qkv[off[b]:off[b + 1], D:2 * D], qkv[off[b]:off[b + 1], 2 * D:]),
                P("self_attn.out_proj.weight"), P("self_attn.out_proj.bias"))
        x1 = ln(x + a, P("norm1.weight"), P("norm1.bias"))
        w, bi = P("multihead_attn.in_proj_weight"), P("multihead_attn.in_proj_bias")
        q2 = lin(x1, w[:D], bi[:D])
        a = mha(q2, lambda b: (lin(mem[b], w[D:2 * D], bi[D:2 * D]), lin(mem[b], w[2 * D:], bi[2 * D:])),
                P("multihead_attn.out_proj.weight"), P("multihead_attn.out_proj.bias"))
        x2 = ln(x1 + a, P("norm2.weight"), P("norm2.bias"))
        h = F.gelu(lin(x2, P("linear1.weight"), P("linear1.bias")))
        return ln(x2 + lin(h, P("linear2.weight"), P("linear2.bias")), P("norm3.weight"), P("norm3.bias"))

    xs = []
    for i in range(4):
        x = layer(f"input_blocks.{i}", x)
        xs.append(x)
    x = layer("middle_block", x)
    for i in range(4):
        w, bi = sd[f"vae.decoder.linear_blocks.{i}.weight"], sd[f"vae.decoder.linear_blocks.{i}.bias"]
        x = lin(x, w[:, :D]) + lin(xs.pop(), w[:, D:]) + bi
        x = layer(f"output_blocks.{i}", x)
    x = ln(x, sd["vae.decoder.norm.weight"], sd["vae.decoder.norm.bias"])
    y = lin(x, sd["vae.final_layer.weight"], sd["vae.final_layer.bias"])
    out = torch.zeros(B, max(lengths), y.shape[-1])
    for b in range(B):
        out[b, :lengths[b]] = y[off[b]:off[b + 1]]
    return out
