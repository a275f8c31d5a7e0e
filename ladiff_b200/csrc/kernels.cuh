// Non-GEMM kernels of the sampling path: row bookkeeping, embeddings, LayerNorm, the two tiny-key attentions,
// the ragged (varlen) decoder self-attention, the fused final-LN + CFG + DDIM step, weight packing, feats2joints.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// row bookkeeping: cnt[S] (rows per sequence) -> off[S+1] (exclusive prefix), total, row -> (seq, t) maps.
// Sequence i takes cnt[i % mod] rows (mod = S: plain; mod = B: the CFG-doubled batch shares one m[B] vector).
__global__ void __launch_bounds__(1024) k_scan_counts(const int* __restrict__ cnt, int S, int mod, int* __restrict__ off,
                                                      int* __restrict__ total) {
  __shared__ int part[1024];
  const int chunk = (S + 1023) / 1024;
  const int b = threadIdx.x * chunk, e = min(S, b + chunk);
  int s = 0;
  for (int i = b; i < e; ++i) s += cnt[i % mod];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 1024; ++i) {
      const int t = part[i];
      part[i] = run;
      run += t;
    }
    off[S] = run;
    *total = run;
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int i = b; i < e; ++i) {
    off[i] = run;
    run += cnt[i % mod];
  }
}

// one warp per sequence; dst_stride > 0 additionally writes row_dst[row] = seq * dst_stride + t (decoder scatter map)
__global__ void k_fill_rows(const int* __restrict__ off, int S, int* __restrict__ row_seq, int* __restrict__ row_t,
                            int* __restrict__ row_dst, int dst_stride) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= S) return;
  const int r0 = off[s], n = off[s + 1] - r0;
  for (int t = lane; t < n; t += 32) {
    row_seq[r0 + t] = s;
    row_t[r0 + t] = t;
    if (row_dst) row_dst[r0 + t] = s * dst_stride + t;
  }
}

// ------------------------------------------------------------------------------------------------
// generic fp32 -> Act conversions
enum Unary : int { U_COPY = 0, U_RELU = 1, U_SILU = 2 };

__global__ void k_unary(const float* __restrict__ in, int ld_in, int rows, int cols, int op, Act out, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(rows) * cols) return;
  const int r = i / cols, c = i % cols;
  float v = in[static_cast<long>(r) * ld_in + c];
  if (op == U_RELU) v = fmaxf(v, 0.f);
  else if (op == U_SILU) v = silu(v);
  act_store(out, planes, r, c, v);
}

// emb_proj's ReLU (architectures/ladiff_denoiser.py:72-73) on the text rows of one chain of prompts [b0, b0+Bc):
// chain row r < Bc is the uncond row b0+r of src [2*Btot,768], row Bc+r the cond row Btot+b0+r (ladiff.py:258-264).
__global__ void k_text_relu(const float* __restrict__ src, int Btot, int b0, int Bc, Act out, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 2L * Bc * 768) return;
  const int r = i / 768, c = i % 768;
  const long sr = r < Bc ? b0 + r : static_cast<long>(Btot) + b0 + (r - Bc);
  act_store(out, planes, r, c, fmaxf(src[sr * 768 + c], 0.f));
}

// LayerNorm over 256 columns, one warp per row (lane = 8 consecutive columns: two 16-byte loads, 16- / 8-byte stores; `in` rows must
// be 16-byte aligned: ld_in % 4 == 0).  rows = min(rows_max, *rows_dev).
__global__ void k_layernorm256(const float* __restrict__ in, int ld_in, int rows_max, const int* __restrict__ rows_dev,
                               const float* __restrict__ g, const float* __restrict__ b, Act out, int planes) {
  pdl_prologue();
  const int rows = rows_dev ? min(rows_max, *rows_dev) : rows_max;
  const long row = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, c0 = lane * 8;
  if (row >= rows) return;
  const float4 x0 = *reinterpret_cast<const float4*>(in + row * ld_in + c0), x1 = *reinterpret_cast<const float4*>(in + row * ld_in + c0 + 4);
  float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float d = v[j] - mean;
    q += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.f / 256.f) + LD_EPS);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c0)), g1 = __ldg(reinterpret_cast<const float4*>(g + c0 + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + 4));
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float y[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = (v[j] - mean) * rstd * gg[j] + bb[j];
  const long o = row * out.ld + c0;
  if (out.f32) {
    *reinterpret_cast<float4*>(out.f32 + o) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(out.f32 + o + 4) = make_float4(y[4], y[5], y[6], y[7]);
  }
  if (out.pl && planes > 0) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split2_op(y[2 * k], y[2 * k + 1], planes, hi[k], lo[k]);
    *reinterpret_cast<uint4*>(out.pl + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes > 1) *reinterpret_cast<uint4*>(out.pl + static_cast<long>(out.rows_alloc) * out.ld + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Sinusoidal timestep embedding, flip_sin_to_cos=True, freq_shift=0, dim 768 -> [cos | sin]
// (architectures/tools/embeddings.py:245-285).  Frequencies are formed in double and rounded to fp32 so the
// fp32 product t*f matches torch's to the last bit in almost every entry.
__global__ void k_sinus_embed(const int* __restrict__ timesteps, int n, Act out, int planes) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 768) return;
  const int r = i / 768, c = i % 768;
  const int j = c % 384;
  const float f = static_cast<float>(exp(-9.210340371976184 * static_cast<double>(static_cast<float>(j)) / 384.0));
  // torch: exponent = (-log(10000) * arange(fp32)) / 384 in fp32, then exp in fp32
  const float ex = (static_cast<float>(-9.210340371976184) * static_cast<float>(j)) / 384.0f;
  const float f32 = static_cast<float>(exp(static_cast<double>(ex)));
  (void)f;
  const float arg = static_cast<float>(timesteps[r]) * f32;
  const float v = (c < 384) ? static_cast<float>(cos(static_cast<double>(arg))) : static_cast<float>(sin(static_cast<double>(arg)));
  act_store(out, planes, r, c, v);
}

// ca_block hoist: a[(n, s), :] = SiLU(lny[s, :] * (1 + scale[n, :]) + shift[n, :]),  mod row n = [scale(256) | shift(256)]
// (architectures/mdiff_transformer.py:158-162 with h = LN(y) hoisted out of the step loop).
__global__ void k_ca_prologue(const float* __restrict__ lny, int ld_lny, const float* __restrict__ mod, int ld_mod,
                              int n_steps, int S, Act out, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(n_steps) * S * 256) return;
  const int c = i & 255;
  const long rs = i >> 8;
  const int s = rs % S, n = rs / S;
  const float v = lny[static_cast<long>(s) * ld_lny + c] * (1.f + mod[static_cast<long>(n) * ld_mod + c]) +
                  mod[static_cast<long>(n) * ld_mod + 256 + c];
  act_store(out, planes, rs, c, silu(v));
}

// x[row, :] = src[(seq % src_mod) * T + t, :] + pe[t, :]   (query_pos: architectures/ladiff_denoiser.py:251)
__global__ void k_pack_x(const float* __restrict__ src, int src_mod, int T, const float* __restrict__ pe,
                         const int* __restrict__ row_seq, const int* __restrict__ row_t, const int* __restrict__ R_dev,
                         Act x, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long row = i >> 8;
  const int c = i & 255;
  if (row >= *R_dev) return;
  const int s = row_seq[row], t = row_t[row];
  act_store(x, planes, row, c, src[(static_cast<long>(s % src_mod) * T + t) * 256 + c] + pe[t * 256 + c]);
}

// ------------------------------------------------------------------------------------------------
// denoiser self-attention (sa_block, architectures/mdiff_transformer.py:307-313): per (sequence, head) the m valid
// latent rows attend to [m latent rows ; text token ; time token].  One warp per (seq, head); lane owns 2 of 64 dims.
// The conditioning tokens only act as keys/values (their own outputs are discarded at :313).
template <int MAXT>
__global__ void k_attn_small(const float* __restrict__ qkv, const int* __restrict__ off, int S,
                             const float* __restrict__ textkv, int ld_textkv, const float* __restrict__ timekv,
                             Act out, int planes) {
  pdl_prologue();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int s = gw >> 2, h = gw & 3;
  if (s >= S) return;
  const int r0 = off[s], m = min(off[s + 1] - r0, MAXT);
  const int d = h * 64 + 2 * lane;
  float2 k[MAXT + 2], v[MAXT + 2];
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    if (j < m) {
      k[j] = *reinterpret_cast<const float2*>(qkv + static_cast<long>(r0 + j) * 768 + 256 + d);
      v[j] = *reinterpret_cast<const float2*>(qkv + static_cast<long>(r0 + j) * 768 + 512 + d);
    } else {
      k[j] = make_float2(0.f, 0.f);
      v[j] = make_float2(0.f, 0.f);
    }
  }
  k[MAXT] = *reinterpret_cast<const float2*>(textkv + static_cast<long>(s) * ld_textkv + d);
  v[MAXT] = *reinterpret_cast<const float2*>(textkv + static_cast<long>(s) * ld_textkv + 256 + d);
  k[MAXT + 1] = *reinterpret_cast<const float2*>(timekv + d);
  v[MAXT + 1] = *reinterpret_cast<const float2*>(timekv + 256 + d);
  for (int i = 0; i < m; ++i) {
    float2 q = *reinterpret_cast<const float2*>(qkv + static_cast<long>(r0 + i) * 768 + d);
    q.x *= 0.125f;  // 1/sqrt(64), applied to q before q.k^T like nn.MultiheadAttention
    q.y *= 0.125f;
    float sc[MAXT + 2], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXT + 2; ++j) {
      sc[j] = warp_sum(q.x * k[j].x + q.y * k[j].y);
      if (j >= m && j < MAXT) sc[j] = -INFINITY;  // masked latent slots (key_padding_mask)
      mx = fmaxf(mx, sc[j]);
    }
    float den = 0.f, ox = 0.f, oy = 0.f;
#pragma unroll
    for (int j = 0; j < MAXT + 2; ++j) {
      const float p = expf(sc[j] - mx);
      den += p;
      ox += p * v[j].x;
      oy += p * v[j].y;
    }
    const float inv = 1.0f / den;
    act_store(out, planes, r0 + i, d, ox * inv);
    act_store(out, planes, r0 + i, d + 1, oy * inv);
  }
}

__device__ __forceinline__ unsigned long long ktimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// sa_block attention with the out-projection folded into the values, + residual + LayerNorm (norm1) in one kernel.
// The in-projection is extended at load time to  [ q(256) | k(256) | v'_0 .. v'_3 (4 x 256) | X(256) ]  with
// v'_h = W_o[:, head h] (W_v[head h] x + b_v[head h])  (nn.MultiheadAttention out_proj is linear in the per-head values) and
// X = the layer's input tokens themselves (folded residual / skip merge, see fold_denoiser_transitions), so
//   x1[row] = LN( X[row] + b_o + sum_h sum_j softmax_j(q_h . k_hj / 8) v'_hj )        (mdiff_transformer.py:54-62,307-313)
// and neither the out-proj GEMM nor the residual GEMM is a separate link of the step's dependency chain.
// CTA = sequence (128 threads).  Every global load (q, k rows -> smem; v' and residual -> registers) is issued up front, so
// the kernel pays one memory round trip; scores: one (row, head, key) dot product per thread; output: 2 columns per thread.
#define DQ_LD 1536    // q | k | 4 x v'
#define DQX_LD 1792   // ... | X : row pitch of the in-projection buffer
#define DC_LD 1280    // conditioning-token table row per layer: k | 4 x v'
// SEQS sequences per CTA (256 threads each, independent named barriers): SEQS = 2 halves the grid to <= 148 CTAs at B = 128
template <int MAXT, int SEQS>
// min 2 CTAs per SM = at most 128 registers: what matters is the CAP it implies -- never 3 CTAs on one SM.  The kernel starts under
// the tail of the in-projection (PDL), whose CTAs still hold most SMs; with 3 allowed the few free SMs take 3 CTAs each and become
// the critical path of this latency-bound link (same-box A/B, profiles/r02_attn_ln_occupancy.txt: reverse loop 20.98 -> 20.52 ms).
__global__ void __launch_bounds__(256 * SEQS, SEQS == 1 ? 2 : 1) k_attn_ln(const float* __restrict__ qkvx, const int* __restrict__ off, int S,
                                                 const float* __restrict__ textkv, int ld_textkv,
                                                 const float* __restrict__ timekv, const float* __restrict__ res, int ld_res,
                                                 const float* __restrict__ bo, const float* __restrict__ g,
                                                 const float* __restrict__ b, Act out, Act xcopy, int planes,
                                                 unsigned long long* trace) {
  if (trace && threadIdx.x == 0) atomicMin(trace, ktimer());
  constexpr int NK = MAXT + 2;
  constexpr int G = NK <= 8 ? 8 : 16;   // lanes per (row, head) group of the score phase = padded key count of a probability row
  __shared__ __align__(16) float Qs_[SEQS][MAXT][256];
  __shared__ __align__(16) float Ks_[SEQS][NK][256];
  __shared__ __align__(16) float Ps_[SEQS][MAXT][4][G];
  __shared__ __align__(16) float red_[SEQS][2][MAXT][8];   // [sum | M2][row][warp]
  const int half = SEQS > 1 ? threadIdx.x >> 8 : 0, tid = threadIdx.x & 255;
  float (*Qs)[256] = Qs_[half];
  float (*Ks)[256] = Ks_[half];
  float (*Ps)[4][G] = Ps_[half];
  float (*red)[MAXT][8] = red_[half];
  const int s = blockIdx.x * SEQS + half;
  // the row offsets and the per-column vectors are written once per call, long before the previous grid: load them BEFORE the
  // dependency wait (a dependent global round trip costs ~1.8 us on the critical path of every layer otherwise)
  const int r0 = s < S ? __ldg(off + s) : 0, m = s < S ? min(__ldg(off + s + 1) - r0, MAXT) : 0;
  const float boc = __ldg(bo + tid), gc = __ldg(g + tid), bc = __ldg(b + tid);
  pdl_prologue();
  if (trace && tid == 0) atomicMin(trace + 1, ~ktimer());
  if (s >= S) return;
  if (m <= 0) return;
  const int warp = tid >> 5, lane = tid & 31;
  const int c = tid;  // this thread's output column
  const float* tk = textkv + static_cast<long>(s) * ld_textkv;
  // ---- all global loads up front (one memory round trip): k rows -> smem, v' and the residual -> registers, q -> smem
  float v[4][NK], xr[MAXT];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const bool on = j < m || j >= MAXT;
    // row base such that k sits at +256 and v'_h at +512 + 256 h for latent rows and conditioning-table rows alike
    const float* base = (j < MAXT) ? qkvx + static_cast<long>(r0 + j) * DQX_LD : (j == MAXT ? tk - 256 : timekv - 256);
    Ks[j][c] = on ? base[256 + c] : 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) v[h][j] = on ? base[512 + h * 256 + c] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    float q = 0.f;
    xr[i] = 0.f;
    if (i < m) {
      q = qkvx[static_cast<long>(r0 + i) * DQX_LD + c];
      xr[i] = res[static_cast<long>(r0 + i) * ld_res + c];
    }
    Qs[i][c] = q * 0.125f;  // 1/sqrt(64) on q, like nn.MultiheadAttention
  }
  asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
  if (trace && tid == 0) atomicMin(trace + 2, ~ktimer());
  // ---- scores + softmax in one phase.  The compute part of this kernel is bound by shared-memory INSTRUCTIONS, not by math (a
  // scalar broadcast load per FMA): so the 64-long dot products read q and k as float4 (16-byte chunks rotated by the lane: the 8
  // lanes of a quarter-warp phase hit 32 distinct banks), the G lanes of a group hold the keys of one (row, head) and do the
  // softmax with shuffles (no shared-memory round trip, no single-warp serial pass), and the probabilities are stored as padded
  // rows of G floats that the output phase reads back as float4.
  for (int t = tid; t < MAXT * 4 * G; t += 256) {      // warp-uniform trip count (multiples of 32)
    const int j = t % G, ih = t / G, h = ih & 3, i = ih >> 2;
    const bool on = i < m && j < NK && (j < m || j >= MAXT);
    float sc = -INFINITY;
    if (on) {
      const float4* qp = reinterpret_cast<const float4*>(&Qs[i][h * 64]);
      const float4* kp = reinterpret_cast<const float4*>(&Ks[j][h * 64]);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const int dd = (d + lane) & 15;
        const float4 q4 = qp[dd], k4 = kp[dd];
        a0 = fmaf(q4.x, k4.x, a0);
        a1 = fmaf(q4.y, k4.y, a1);
        a2 = fmaf(q4.z, k4.z, a2);
        a3 = fmaf(q4.w, k4.w, a3);
      }
      sc = (a0 + a1) + (a2 + a3);
    }
    float mx = sc;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float pe = on ? expf(sc - mx) : 0.f;       // rows >= m: every lane off -> probabilities 0
    float den = pe;
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    Ps[i][h][j] = den > 0.f ? pe / den : 0.f;
  }
  asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
  if (trace && tid == 0) atomicMin(trace + 5, ~ktimer());
  // ---- out[i, c] = sum_h sum_j P[i][h][j] v'[h][j][c]  + out_proj bias + residual
  float acc[MAXT];
#pragma unroll
  for (int i = 0; i < MAXT; ++i) acc[i] = boc + xr[i];
#pragma unroll
  for (int i = 0; i < MAXT; ++i)
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float pr[G];
#pragma unroll
      for (int q = 0; q < G / 4; ++q) {
        const float4 p4 = *reinterpret_cast<const float4*>(&Ps[i][h][4 * q]);   // broadcast float4: G / 4 loads per (row, head)
        pr[4 * q] = p4.x; pr[4 * q + 1] = p4.y; pr[4 * q + 2] = p4.z; pr[4 * q + 3] = p4.w;
      }
#pragma unroll
      for (int j = 0; j < NK; ++j) acc[i] = fmaf(pr[j], v[h][j], acc[i]);        // rows >= m: P == 0
    }
  if (trace && tid == 0) atomicMin(trace + 6, ~ktimer());
  // ---- LayerNorm over the 256 columns (one per thread), exact two-pass variance with ONE CTA barrier: every warp reduces its
  // 32 columns about its own mean, the 8 (sum, M2) pairs are combined with the parallel-variance formula
  //   M2 = sum_w M2_w + 32 sum_w (mean_w - mean)^2
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    const float s1 = warp_sum(i < m ? acc[i] : 0.f);
    const float d = acc[i] - s1 * (1.f / 32.f);
    const float s2 = warp_sum(i < m ? d * d : 0.f);
    if (lane == 0) {
      red[0][i][warp] = s1;
      red[1][i][warp] = s2;
    }
  }
  asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
  if (trace && tid == 0) atomicMin(trace + 7, ~ktimer());
  // normalised rows (and the layer input, when it is kept) go through the dead q / k tiles so that the global stores below are
  // 16 bytes (fp32) / 8 bytes (a plane) per thread instead of one 4- / 2-byte element: 4x fewer store instructions
  const bool keep_x = xcopy.f32 || xcopy.pl;
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    if (i < m) {
      const float4 sa = *reinterpret_cast<const float4*>(&red[0][i][0]), sb = *reinterpret_cast<const float4*>(&red[0][i][4]);
      const float4 qa = *reinterpret_cast<const float4*>(&red[1][i][0]), qb = *reinterpret_cast<const float4*>(&red[1][i][4]);
      const float mean = ((sa.x + sa.y) + (sa.z + sa.w) + (sb.x + sb.y) + (sb.z + sb.w)) * (1.f / 256.f);
      const float sw[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
      float m2 = (qa.x + qa.y) + (qa.z + qa.w) + (qb.x + qb.y) + (qb.z + qb.w);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float dm = sw[w] * (1.f / 32.f) - mean;
        m2 = fmaf(32.f * dm, dm, m2);
      }
      const float rstd = 1.0f / sqrtf(m2 * (1.f / 256.f) + LD_EPS);
      Qs[i][c] = (acc[i] - mean) * rstd * gc + bc;
      if (keep_x) Ks[i][c] = xr[i];
    }
  }
  asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
  for (int idx = tid; idx < m * 64; idx += 256) {
    const int i = idx >> 6, c4 = (idx & 63) * 4;
    const long row = r0 + i;
    const float4 y = *reinterpret_cast<const float4*>(&Qs[i][c4]);
    if (out.f32) *reinterpret_cast<float4*>(out.f32 + row * out.ld + c4) = y;
    if (out.pl && planes > 0) {
      uint32_t h0, l0, h1, l1;
      split2_op(y.x, y.y, planes, h0, l0);
      split2_op(y.z, y.w, planes, h1, l1);
      *reinterpret_cast<uint2*>(out.pl + row * out.ld + c4) = make_uint2(h0, h1);
      if (planes > 1) *reinterpret_cast<uint2*>(out.pl + (static_cast<long>(out.rows_alloc) + row) * out.ld + c4) = make_uint2(l0, l1);
    }
    // the layer input itself, kept for the U-Net skip connections (third K source of the mirrored layer's in-projection)
    if (keep_x) {
      const float4 x = *reinterpret_cast<const float4*>(&Ks[i][c4]);
      if (xcopy.f32) *reinterpret_cast<float4*>(xcopy.f32 + row * xcopy.ld + c4) = x;
      if (xcopy.pl && planes > 0) {
        uint32_t h0, l0, h1, l1;
        split2_op(x.x, x.y, planes, h0, l0);
        split2_op(x.z, x.w, planes, h1, l1);
        *reinterpret_cast<uint2*>(xcopy.pl + row * xcopy.ld + c4) = make_uint2(h0, h1);
        if (planes > 1) *reinterpret_cast<uint2*>(xcopy.pl + (static_cast<long>(xcopy.rows_alloc) + row) * xcopy.ld + c4) = make_uint2(l0, l1);
      }
    }
  }
  if (trace && tid == 0) atomicMin(trace + 3, ~ktimer());
}

// The same block with 128 threads per sequence (two adjacent output columns per thread) and at most 120 registers: a CTA then holds
// 15 360 registers and ~13 KB of shared memory, so TWO of them fit next to the in-projection CTA that still occupies the SM
// (33 280 registers, 197 KB) and all 256 sequences of a B = 128 step are resident -- row offsets and per-column vectors loaded --
// when the in-projection ends, instead of ~100 CTAs being launched behind its tail.  __maxnreg__(120) also keeps the per-SM
// cap at 4 CTAs = 16 warps, the same as the 2 x 256-thread cap of k_attn_ln (more resident warps measured slower,
// profiles/r02_attn_ln_occupancy.txt).
template <int MAXT>
__global__ void __maxnreg__(120) k_attn_ln2(const float* __restrict__ qkvx, const int* __restrict__ off, int S,
                                                     const float* __restrict__ textkv, int ld_textkv,
                                                     const float* __restrict__ timekv, const float* __restrict__ res, int ld_res,
                                                     const float* __restrict__ bo, const float* __restrict__ g,
                                                     const float* __restrict__ b, Act out, Act xcopy, int planes,
                                                     unsigned long long* trace) {
  if (trace && threadIdx.x == 0) atomicMin(trace, ktimer());
  constexpr int NK = MAXT + 2;
  __shared__ __align__(16) float Qs[MAXT][256];
  __shared__ __align__(16) float Ks[NK][256];
  __shared__ float Ps[MAXT][4][NK];
  __shared__ float red[2][4][MAXT];
  const int tid = threadIdx.x, s = blockIdx.x;
  const int c = 2 * tid;   // this thread's two output columns
  const int r0 = s < S ? __ldg(off + s) : 0, m = s < S ? min(__ldg(off + s + 1) - r0, MAXT) : 0;
  const float2 boc = __ldg(reinterpret_cast<const float2*>(bo + c)), gc = __ldg(reinterpret_cast<const float2*>(g + c)),
               bc = __ldg(reinterpret_cast<const float2*>(b + c));
  pdl_prologue();
  if (trace && tid == 0) atomicMin(trace + 1, ~ktimer());
  if (s >= S || m <= 0) return;
  const int warp = tid >> 5, lane = tid & 31;
  const float* tk = textkv + static_cast<long>(s) * ld_textkv;
  // ---- all global loads up front (one memory round trip)
  float2 v[4][NK], xr[MAXT];
#pragma unroll
  for (int j = 0; j < NK; ++j) {
    const bool on = j < m || j >= MAXT;
    const float* base = (j < MAXT) ? qkvx + static_cast<long>(r0 + j) * DQX_LD : (j == MAXT ? tk - 256 : timekv - 256);
    const float2 kk = on ? *reinterpret_cast<const float2*>(base + 256 + c) : make_float2(0.f, 0.f);
    *reinterpret_cast<float2*>(&Ks[j][c]) = kk;
#pragma unroll
    for (int h = 0; h < 4; ++h) v[h][j] = on ? *reinterpret_cast<const float2*>(base + 512 + h * 256 + c) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    float2 q = make_float2(0.f, 0.f);
    xr[i] = make_float2(0.f, 0.f);
    if (i < m) {
      q = *reinterpret_cast<const float2*>(qkvx + static_cast<long>(r0 + i) * DQX_LD + c);
      xr[i] = *reinterpret_cast<const float2*>(res + static_cast<long>(r0 + i) * ld_res + c);
    }
    *reinterpret_cast<float2*>(&Qs[i][c]) = make_float2(q.x * 0.125f, q.y * 0.125f);
  }
  __syncthreads();
  if (trace && tid == 0) atomicMin(trace + 2, ~ktimer());
  // ---- scores: element e = (i, h, j), one 64-long dot product each
  for (int e = tid; e < m * 4 * NK; e += 128) {
    const int j = e % NK, h = (e / NK) & 3, i = e / (4 * NK);
    float sc = -INFINITY;
    if (j < m || j >= MAXT) {
      const float* qp = &Qs[i][h * 64];
      const float* kp = &Ks[j][h * 64];
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int d = 0; d < 64; d += 4) {
        const int d0 = (d + lane) & 63, d1 = (d + 1 + lane) & 63, d2 = (d + 2 + lane) & 63, d3 = (d + 3 + lane) & 63;
        a0 = fmaf(qp[d0], kp[d0], a0);
        a1 = fmaf(qp[d1], kp[d1], a1);
        a2 = fmaf(qp[d2], kp[d2], a2);
        a3 = fmaf(qp[d3], kp[d3], a3);
      }
      sc = (a0 + a1) + (a2 + a3);
    }
    Ps[i][h][j] = sc;
  }
  __syncthreads();
  if (tid < m * 4) {
    const int i = tid >> 2, h = tid & 3;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NK; ++j) mx = fmaxf(mx, Ps[i][h][j]);
    float pj[NK], den = 0.f;
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      pj[j] = expf(Ps[i][h][j] - mx);
      den += pj[j];
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < NK; ++j) Ps[i][h][j] = pj[j] * inv;
  }
  __syncthreads();
  // ---- out[i, c..c+1] = sum_h sum_j P[i][h][j] v'[h][j] + out_proj bias + residual
  float2 acc[MAXT];
#pragma unroll
  for (int i = 0; i < MAXT; ++i) acc[i] = make_float2(boc.x + xr[i].x, boc.y + xr[i].y);
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int j = 0; j < NK; ++j)
#pragma unroll
      for (int i = 0; i < MAXT; ++i) {
        const float w = Ps[i][h][j];
        acc[i].x = fmaf(w, v[h][j].x, acc[i].x);
        acc[i].y = fmaf(w, v[h][j].y, acc[i].y);
      }
  // ---- LayerNorm over the 256 columns (two per thread, 4 warps)
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    const float w = warp_sum(i < m ? acc[i].x + acc[i].y : 0.f);
    if (lane == 0) red[0][warp][i] = w;
  }
  __syncthreads();
  float mean[MAXT];
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    const float t = (red[0][0][i] + red[0][1][i]) + (red[0][2][i] + red[0][3][i]);
    mean[i] = t * (1.f / 256.f);
    const float dx = acc[i].x - mean[i], dy = acc[i].y - mean[i];
    const float w = warp_sum(i < m ? dx * dx + dy * dy : 0.f);
    if (lane == 0) red[1][warp][i] = w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MAXT; ++i) {
    if (i < m) {
      const float t = (red[1][0][i] + red[1][1][i]) + (red[1][2][i] + red[1][3][i]);
      const float rstd = 1.0f / sqrtf(t * (1.f / 256.f) + LD_EPS);
      const long row = r0 + i;
      const float y0 = (acc[i].x - mean[i]) * rstd * gc.x + bc.x, y1 = (acc[i].y - mean[i]) * rstd * gc.y + bc.y;
      if (out.f32) *reinterpret_cast<float2*>(out.f32 + row * out.ld + c) = make_float2(y0, y1);
      if (out.pl && planes > 0) {
        uint32_t hi, lo;
        split2_op(y0, y1, planes, hi, lo);
        *reinterpret_cast<uint32_t*>(out.pl + row * out.ld + c) = hi;
        if (planes > 1) *reinterpret_cast<uint32_t*>(out.pl + (static_cast<long>(out.rows_alloc) + row) * out.ld + c) = lo;
      }
      // the layer input itself, kept for the U-Net skip connections (third K source of the mirrored layer's in-projection)
      if (xcopy.f32) *reinterpret_cast<float2*>(xcopy.f32 + row * xcopy.ld + c) = xr[i];
      if (xcopy.pl && planes > 0) {
        uint32_t hi, lo;
        split2_op(xr[i].x, xr[i].y, planes, hi, lo);
        *reinterpret_cast<uint32_t*>(xcopy.pl + row * xcopy.ld + c) = hi;
        if (planes > 1) *reinterpret_cast<uint32_t*>(xcopy.pl + (static_cast<long>(xcopy.rows_alloc) + row) * xcopy.ld + c) = lo;
      }
    }
  }
  if (trace && tid == 0) atomicMin(trace + 3, ~ktimer());
}

// Final LayerNorm (encoder.norm) + CFG combine + DDIM step + next-step input, one warp per (prompt, latent row):
//   eps = LN(tok_u) + g (LN(tok_c) - LN(tok_u));  lat' = c1 lat + c2 eps   (models/modeltype/ladiff.py:487-492)
//   x_next[row_u] = x_next[row_c] = lat' + pe[t]                           (ladiff.py:472-474 + ladiff_denoiser.py:251)
// Philox4x32-10 counter-based generator + Box-Muller: the variance noise of the DDPM scheduler step (diffusers draws
// torch.randn inside DDPMScheduler.step; here the stream is a pure function of (seed, step, element) so a captured graph
// replays with fresh noise by bumping the seed in device memory).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

// One scheduler step of the reverse loop, fused with what surrounds it (ladiff.py:487-492 + encoder.norm of
// ladiff_denoiser.py + query_pos of the next step): final LayerNorm of both CFG halves, eps = u + g (c - u),
// x' = c1 x + c2 eps + c3 noise, next input x' + pe (fp32 + bf16 planes).  coef = {c1, c2, c3, guidance} of this step (device
// table).  c3 != 0 (DDPM variance term): noise = (*noise_pp)[step] when a tensor was injected, else Philox(seed, step, element).
// flags & 1 (ARDIFF, ladiff.py:419-467): only latent slot 0 is being denoised, slots >= 1 are fixed context latents.
// elem0 / n_elem: offset of this chain's latents inside the call's [Btot, T, 256] tensor and its size (noise addressing).
__global__ void k_cfg_ddim(const float* __restrict__ tok, const int* __restrict__ off, int B, int T,
                           const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ coef,
                           float* __restrict__ lat, const float* __restrict__ pe, Act x, int planes, int flags, int step,
                           const float* const* __restrict__ noise_pp, const unsigned long long* __restrict__ seed_p,
                           long elem0, long n_elem) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int bi = gw / T, t = gw % T;
  // row offsets: written once per call, long before the previous grid -> loaded before the dependency wait
  int o0 = 0, o1 = 0, oc = 0;
  if (bi < B) {
    o0 = __ldg(off + bi);
    o1 = __ldg(off + bi + 1);
    oc = __ldg(off + bi + B);
  }
  pdl_prologue();
  if (bi >= B) return;
  const int m = o1 - o0;
  if (t >= m) return;
  const long ru = o0 + t, rc = oc + t;
  float u[8], c[8], su = 0.f, sc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    u[j] = tok[ru * 256 + lane + 32 * j];
    c[j] = tok[rc * 256 + lane + 32 * j];
    su += u[j];
    sc += c[j];
  }
  const float mu = warp_sum(su) * (1.f / 256.f), mc = warp_sum(sc) * (1.f / 256.f);
  float qu = 0.f, qc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    qu += (u[j] - mu) * (u[j] - mu);
    qc += (c[j] - mc) * (c[j] - mc);
  }
  const float ru_ = 1.0f / sqrtf(warp_sum(qu) * (1.f / 256.f) + LD_EPS);
  const float rc_ = 1.0f / sqrtf(warp_sum(qc) * (1.f / 256.f) + LD_EPS);
  const float c1 = coef[0], c2 = coef[1], c3 = coef[2], guidance = coef[3];
  const bool frozen = (flags & 1) && t > 0;
  const float* nz = (c3 != 0.f && noise_pp) ? *noise_pp : nullptr;
  const unsigned long long seed = (c3 != 0.f && !nz && seed_p) ? *seed_p : 0ull;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = lane + 32 * j;
    const float eu = (u[j] - mu) * ru_ * g[col] + b[col];
    const float ec = (c[j] - mc) * rc_ * g[col] + b[col];
    const float eps = eu + guidance * (ec - eu);
    const long li = (static_cast<long>(bi) * T + t) * 256 + col;
    float nl = lat[li];
    if (!frozen) {
      nl = c1 * nl + c2 * eps;
      if (c3 != 0.f) {
        float zn;
        if (nz) {
          zn = nz[static_cast<long>(step) * n_elem + elem0 + li];
        } else {   // one Philox block per element pair; element e uses half (e & 1) of block e >> 1
          const long e = elem0 + li;
          const uint4 r4 = philox4x32_10(make_uint4(static_cast<uint32_t>(e >> 1), static_cast<uint32_t>(e >> 33), static_cast<uint32_t>(step), 0x4C414466u),
                                         make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
          const float2 n2 = box_muller(r4.x, r4.y);
          zn = (e & 1) ? n2.y : n2.x;
        }
        nl = fmaf(c3, zn, nl);
      }
      lat[li] = nl;
    }
    const float xn = nl + pe[t * 256 + col];
    act_store(x, planes, ru, col, xn);
    act_store(x, planes, rc, col, xn);
  }
}

// standalone CFG + DDIM on dense [B,T,256] tensors (ladiff_cfg_ddim_step)
__global__ void k_cfg_ddim_dense(const float* __restrict__ pred, float* __restrict__ lat, long n_half, float guidance,
                                 float c1, float c2) {
  const long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n_half) return;
  const float4 u = *reinterpret_cast<const float4*>(pred + i);
  const float4 c = *reinterpret_cast<const float4*>(pred + n_half + i);
  float4 l = *reinterpret_cast<float4*>(lat + i);
  l.x = c1 * l.x + c2 * (u.x + guidance * (c.x - u.x));
  l.y = c1 * l.y + c2 * (u.y + guidance * (c.y - u.y));
  l.z = c1 * l.z + c2 * (u.z + guidance * (c.z - u.z));
  l.w = c1 * l.w + c2 * (u.w + guidance * (c.w - u.w));
  *reinterpret_cast<float4*>(lat + i) = l;
}

// out[s, t, :] = t < m[s] ? LN(tok[row]) : 0      (standalone denoiser_forward: encoder.norm + un-pack)
__global__ void k_final_ln_out(const float* __restrict__ tok, const int* __restrict__ off, int S, int T,
                               const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ out) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int s = gw / T, t = gw % T;
  if (s >= S) return;
  float* o = out + (static_cast<long>(s) * T + t) * 256;
  const int m = off[s + 1] - off[s];
  if (t >= m) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o[lane + 32 * j] = 0.f;
    return;
  }
  const long row = off[s] + t;
  float v[8], sm = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = tok[row * 256 + lane + 32 * j];
    sm += v[j];
  }
  const float mean = warp_sum(sm) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) q += (v[j] - mean) * (v[j] - mean);
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.f / 256.f) + LD_EPS);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = lane + 32 * j;
    o[col] = (v[j] - mean) * rstd * g[col] + b[col];
  }
}

// z[t, b, :] = t < m[b] ? lat[b, t, :] : 0      (ladiff.py:500 permute + :562-566 re-zeroing)
__global__ void k_z_out(const float* __restrict__ lat, const int* __restrict__ cnt, int B, int T, float* __restrict__ z) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(B) * T * 256) return;
  const int c = i & 255;
  const long bt = i >> 8;
  const int t = bt % T, b = bt / T;
  const int m = cnt[b];
  z[(static_cast<long>(t) * B + b) * 256 + c] = (t < m) ? lat[i] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// LA-VAE encoder (LADiffVae.encode, architectures/ladiff_vae.py:162-286): ragged token sequence per motion =
// [ m mu tokens | m logvar tokens | L frames ]  (the masked distribution tokens / padded frames of the reference never
// influence a valid row, so they are simply not computed; positions keep the reference's padded layout: mu i at i,
// logvar i at T + i, frame f at 2 T + f).

// features [B, max_len, nfeats] -> packed frame rows [sum L, Kp] (zero-padded columns), fp32 + bf16 planes
__global__ void k_enc_pack_feats(const float* __restrict__ feats, int max_len, int nfeats, int Kp, const int* __restrict__ frow_seq,
                                 const int* __restrict__ frow_t, const int* __restrict__ R_dev, Act out, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long row = i / Kp;
  const int col = static_cast<int>(i % Kp);
  if (row >= *R_dev) return;
  const float v = col < nfeats ? feats[(static_cast<long>(frow_seq[row]) * max_len + frow_t[row]) * nfeats + col] : 0.f;
  act_store(out, planes, row, col, v);
}

// token rows of the encoder input: global_motion_token / embedded frames + learned positions (ladiff_vae.py:189,213,220)
__global__ void k_enc_init(const float* __restrict__ gmt, const float* __restrict__ emb, const float* __restrict__ pe, int T,
                           const int* __restrict__ row_seq, const int* __restrict__ row_t, const int* __restrict__ mcnt,
                           const int* __restrict__ foff, const int* __restrict__ R_dev, Act x, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long row = i >> 8;
  const int col = static_cast<int>(i & 255);
  if (row >= *R_dev) return;
  const int s = row_seq[row], j = row_t[row], m = mcnt[s];
  float v;
  if (j < m) v = gmt[j * 256 + col] + pe[j * 256 + col];
  else if (j < 2 * m) v = gmt[(T + j - m) * 256 + col] + pe[(T + j - m) * 256 + col];
  else v = emb[(static_cast<long>(foff[s]) + (j - 2 * m)) * 256 + col] + pe[(2 * T + j - 2 * m) * 256 + col];
  act_store(x, planes, row, col, v);
}

// mu / std / latent [T, B, 256] from the final tokens (ladiff_vae.py:258-268): std = exp(logvar)^0.5, latent = mu + std eps,
// rows t >= m: latent exactly 0 (and mu = 0, std = 1: the reference leaves values of masked tokens there that nothing reads)
__global__ void k_enc_out(const float* __restrict__ tok, const int* __restrict__ off, const int* __restrict__ mcnt, int B, int T,
                          const float* __restrict__ eps, float* __restrict__ latent, float* __restrict__ mu_out,
                          float* __restrict__ std_out) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(T) * B * 256) return;
  const int col = static_cast<int>(i & 255);
  const long tb = i >> 8;
  const int b = static_cast<int>(tb % B), t = static_cast<int>(tb / B);
  const int m = mcnt[b];
  float mu = 0.f, sd = 1.f, z = 0.f;
  if (t < m) {
    const long r = off[b] + t;
    mu = tok[r * 256 + col];
    sd = sqrtf(expf(tok[(r + m) * 256 + col]));
    z = mu + sd * (eps ? eps[i] : 0.f);
  }
  if (latent) latent[i] = z;
  if (mu_out) mu_out[i] = mu;
  if (std_out) std_out[i] = sd;
}

// ------------------------------------------------------------------------------------------------
// decoder
// queries = 0 + pe[:L]  (architectures/ladiff_vae.py:299,334)
__global__ void k_dec_init(const float* __restrict__ pe, const int* __restrict__ row_t, const int* __restrict__ R_dev,
                           Act x, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;   // one thread per 8 columns
  const long row = i >> 5;
  const int c0 = (i & 31) * 8;
  if (row >= *R_dev) return;
  const float* src = pe + static_cast<long>(row_t[row]) * 256 + c0;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
  const long o = row * x.ld + c0;
  if (x.f32) {
    *reinterpret_cast<float4*>(x.f32 + o) = a;
    *reinterpret_cast<float4*>(x.f32 + o + 4) = b;
  }
  if (x.pl && planes > 0) {
    uint32_t hi[4], lo[4];
    split2_op(a.x, a.y, planes, hi[0], lo[0]);
    split2_op(a.z, a.w, planes, hi[1], lo[1]);
    split2_op(b.x, b.y, planes, hi[2], lo[2]);
    split2_op(b.z, b.w, planes, hi[3], lo[3]);
    *reinterpret_cast<uint4*>(x.pl + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (planes > 1) *reinterpret_cast<uint4*>(x.pl + static_cast<long>(x.rows_alloc) * x.ld + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// zrows[moff[b] + t, :] = z[t, b, :] for t < m[b]  (valid memory rows only)
__global__ void k_gather_z(const float* __restrict__ z, const int* __restrict__ moff, int B, int T, Act out, int planes) {
  pdl_prologue();
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(B) * T * 256) return;
  const int c = i & 255;
  const long bt = i >> 8;
  const int t = bt % T, b = bt / T;
  const int m = moff[b + 1] - moff[b];
  if (t < m) act_store(out, planes, moff[b] + t, c, z[(static_cast<long>(t) * B + b) * 256 + c]);
}

// Decoder cross-attention block (operator/cross_attention.py:373-377, 408-409: multihead_attn(q = tgt, k = v = memory, key-padding =
// the m_i <= 5 valid latents) -> + residual -> norm2) as ONE kernel.  With so few keys the two projections around the attention are
// folded into the per-sequence memory table at weight-load / table-build time (exact algebra, products formed in fp64):
//     score_hj = (x . kq_hj + cq_hj) / 8      kq_hj = W_q[h]^T k_hj  (256-vector),  cq_hj = b_q[h] . k_hj
//     out      = b_o + sum_h sum_j softmax_j(score_h.) v'_hj              v'_hj = W_o[:, h] v_hj  (256-vector)
// so neither the query GEMM nor the out-projection GEMM exists any more: per frame 2 x 20 dot / axpy operations of length 256
// in fp32 plus the LayerNorm.  Table row of a memory latent, per layer (CX_LD floats): kq[4][256] | v'[4][256] | cq[4] | pad.
// CTA = (32 frames, sequence): the sequence's table rows (<= 42 KB) are staged in shared memory once, a warp owns 4 consecutive
// frames so that every table value it reads is used four times; lane <-> 8 consecutive columns; heads are processed one at a time
// (20 live scores instead of 80) so that two CTAs fit per SM.
#define CX_LD 2112
template <int MAXT>
__global__ void __launch_bounds__(256, 2) k_cross_ln(const float* __restrict__ x, const float* __restrict__ tab, int ld_tab, int tab_off,
                                                     const int* __restrict__ foff, const int* __restrict__ moff,
                                                     const float* __restrict__ bo, const float* __restrict__ g,
                                                     const float* __restrict__ bvec, Act out, int planes) {
  constexpr int F = 4;
  extern __shared__ float cx_tab[];                      // [m][CX_LD]: this sequence's table rows of this layer
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f0 = blockIdx.x * 32 + warp * F;
  const int c0 = lane * 8;
  float4 bo4[2], g4[2], b4[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    bo4[q] = __ldg(reinterpret_cast<const float4*>(bo + c0) + q);
    g4[q] = __ldg(reinterpret_cast<const float4*>(g + c0) + q);
    b4[q] = __ldg(reinterpret_cast<const float4*>(bvec + c0) + q);
  }
  pdl_prologue();
  const int r0 = foff[b], L = foff[b + 1] - r0;
  if (static_cast<int>(blockIdx.x) * 32 >= L) return;    // CTA-uniform
  const int m0 = moff[b], m = min(moff[b + 1] - m0, MAXT);
  // ---- the table (kq | v' | cq of the m latents) and this warp's four frames: one global round trip for the whole CTA
  float xv[F][8];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    const bool on = f0 + f < L;
    const float4* xp = reinterpret_cast<const float4*>(x + static_cast<long>(r0 + f0 + f) * 256 + c0);
    const float4 a = on ? xp[0] : make_float4(0.f, 0.f, 0.f, 0.f), c = on ? xp[1] : make_float4(0.f, 0.f, 0.f, 0.f);
    xv[f][0] = a.x; xv[f][1] = a.y; xv[f][2] = a.z; xv[f][3] = a.w; xv[f][4] = c.x; xv[f][5] = c.y; xv[f][6] = c.z; xv[f][7] = c.w;
  }
  for (int i = threadIdx.x; i < m * (CX_LD / 4); i += 256) {
    const int j = i / (CX_LD / 4), c4 = i % (CX_LD / 4);
    reinterpret_cast<float4*>(cx_tab)[i] = __ldg(reinterpret_cast<const float4*>(tab + static_cast<long>(m0 + j) * ld_tab + tab_off) + c4);
  }
  __syncthreads();
  if (f0 >= L) return;                                   // warp-uniform (after the only CTA barrier)
  // ---- y = x + b_o + sum_h sum_j softmax_j((x . kq_hj + cq_hj) / 8) v'_hj, one head at a time
  float y[F][8];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    y[f][0] = xv[f][0] + bo4[0].x; y[f][1] = xv[f][1] + bo4[0].y; y[f][2] = xv[f][2] + bo4[0].z; y[f][3] = xv[f][3] + bo4[0].w;
    y[f][4] = xv[f][4] + bo4[1].x; y[f][5] = xv[f][5] + bo4[1].y; y[f][6] = xv[f][6] + bo4[1].z; y[f][7] = xv[f][7] + bo4[1].w;
  }
  // The 4 frames x MAXT keys partial dot products of a head are reduced over the 32 lanes with a MULTI-VALUE butterfly instead of
  // one 5-step butterfly per value: the xor-16 and xor-8 steps halve the number of values a lane carries (lanes with bits (4, 3)
  // = f keep frame f), only the last three steps run on all MAXT values -- 2 MAXT + MAXT + 3 MAXT = 6 MAXT shuffles instead of
  // 20 MAXT.  Every lane then owns the complete scores of ONE frame, does that frame's softmax once (not all four), and the
  // probabilities come back to all lanes with one indexed shuffle each.
  static_assert(F == 4, "the butterfly below is written for four frames per warp");
  const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
#pragma unroll 1
  for (int h = 0; h < 4; ++h) {
    float pp[F][MAXT];
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      if (j < m) {                                       // warp-uniform
        const float* row = cx_tab + j * CX_LD + h * 256 + c0;
        const float4 a = *reinterpret_cast<const float4*>(row), c = *reinterpret_cast<const float4*>(row + 4);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float p = xv[f][0] * a.x;
          p = fmaf(xv[f][1], a.y, p); p = fmaf(xv[f][2], a.z, p); p = fmaf(xv[f][3], a.w, p);
          p = fmaf(xv[f][4], c.x, p); p = fmaf(xv[f][5], c.y, p); p = fmaf(xv[f][6], c.z, p); p = fmaf(xv[f][7], c.w, p);
          pp[f][j] = p;
        }
      } else {
#pragma unroll
        for (int f = 0; f < F; ++f) pp[f][j] = 0.f;
      }
    }
    float r2[MAXT];
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      // xor 16: frames {0, 1} stay in the lower half-warp, {2, 3} in the upper one
      const float k0 = up16 ? pp[2][j] : pp[0][j], s0 = up16 ? pp[0][j] : pp[2][j];
      const float k1 = up16 ? pp[3][j] : pp[1][j], s1 = up16 ? pp[1][j] : pp[3][j];
      const float q0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
      const float q1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
      // xor 8: the even frame of the pair stays where bit 3 is clear
      const float k2 = up8 ? q1 : q0, s2 = up8 ? q0 : q1;
      r2[j] = k2 + __shfl_xor_sync(0xffffffffu, s2, 8);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
      for (int j = 0; j < MAXT; ++j) r2[j] += __shfl_xor_sync(0xffffffffu, r2[j], o);
    // this lane: scores of frame (lane >> 3) & 3 against the m keys -> softmax
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      r2[j] = (j < m) ? (r2[j] + cx_tab[j * CX_LD + 2048 + h]) * 0.125f : -INFINITY;
      mx = fmaxf(mx, r2[j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      r2[j] = (j < m) ? __expf(r2[j] - mx) : 0.f;
      den += r2[j];
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      if (j < m) {
        const float pj = r2[j] * inv;
        const float* row = cx_tab + j * CX_LD + 1024 + h * 256 + c0;
        const float4 a = *reinterpret_cast<const float4*>(row), c = *reinterpret_cast<const float4*>(row + 4);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const float p = __shfl_sync(0xffffffffu, pj, f * 8);   // lane 8 f holds frame f
          y[f][0] = fmaf(p, a.x, y[f][0]); y[f][1] = fmaf(p, a.y, y[f][1]); y[f][2] = fmaf(p, a.z, y[f][2]); y[f][3] = fmaf(p, a.w, y[f][3]);
          y[f][4] = fmaf(p, c.x, y[f][4]); y[f][5] = fmaf(p, c.y, y[f][5]); y[f][6] = fmaf(p, c.z, y[f][6]); y[f][7] = fmaf(p, c.w, y[f][7]);
        }
      }
    }
  }
  // ---- LayerNorm (exact two-pass) and store: fp32 master + operand planes
#pragma unroll
  for (int f = 0; f < F; ++f) {
    if (f0 + f >= L) break;                              // warp-uniform
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += y[f][k];
    const float mean = warp_sum(s) * (1.f / 256.f);
    float q2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = y[f][k] - mean;
      q2 += d * d;
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.f / 256.f) + LD_EPS);
    const float gg[8] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w};
    const float bb[8] = {b4[0].x, b4[0].y, b4[0].z, b4[0].w, b4[1].x, b4[1].y, b4[1].z, b4[1].w};
    float z[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) z[k] = (y[f][k] - mean) * rstd * gg[k] + bb[k];
    const long o = static_cast<long>(r0 + f0 + f) * out.ld + c0;
    if (out.f32) {
      *reinterpret_cast<float4*>(out.f32 + o) = make_float4(z[0], z[1], z[2], z[3]);
      *reinterpret_cast<float4*>(out.f32 + o + 4) = make_float4(z[4], z[5], z[6], z[7]);
    }
    if (out.pl && planes > 0) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split2_op(z[2 * k], z[2 * k + 1], planes, hi[k], lo[k]);
      *reinterpret_cast<uint4*>(out.pl + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (planes > 1) *reinterpret_cast<uint4*>(out.pl + static_cast<long>(out.rows_alloc) * out.ld + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// Ragged self-attention over the L valid frames of one sequence (operator/cross_attention.py:367-369), fp32.
// CTA = (query block of 64, head, sequence), 8 warps x 8 queries (4 at a time).  K^T and V of the (sequence, head)
// are staged in shared memory once per CTA; nothing is computed or stored for padded frames.
#define SA_MAXL 256
#define SA_QB 64
struct SelfAttnSmem {
  float Kt[64][SA_MAXL + 1];
  float Vs[SA_MAXL][64];
  float Ps[8][SA_MAXL][4];
  float Qs[8][4][64];
};

__global__ void __launch_bounds__(256) k_attn_self(const float* __restrict__ qkv, const int* __restrict__ foff, Act out,
                                                   int planes) {
  pdl_prologue();
  extern __shared__ uint8_t sa_raw[];
  SelfAttnSmem& sm = *reinterpret_cast<SelfAttnSmem*>(sa_raw);
  const int b = blockIdx.z, h = blockIdx.y, qb = blockIdx.x;
  const int r0 = foff[b], L = min(foff[b + 1] - r0, SA_MAXL);
  if (qb * SA_QB >= L) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < L * 64; i += 256) {
    const int key = i >> 6, d = i & 63;
    const float* base = qkv + static_cast<long>(r0 + key) * 768 + h * 64 + d;
    sm.Kt[d][key] = base[256];
    sm.Vs[key][d] = base[512];
  }
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    const int q0 = qb * SA_QB + warp * 8 + pass * 4;  // 4 queries q0..q0+3 (warp-uniform)
    if (q0 >= L) break;
    for (int i = lane; i < 4 * 64; i += 32) {
      const int qi = i >> 6, d = i & 63;
      const int qr = q0 + qi;
      sm.Qs[warp][qi][d] = (qr < L) ? qkv[static_cast<long>(r0 + qr) * 768 + h * 64 + d] * 0.125f : 0.f;
    }
    __syncwarp();
    float sc[4][SA_MAXL / 32];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi)
#pragma unroll
      for (int i = 0; i < SA_MAXL / 32; ++i) sc[qi][i] = 0.f;
    for (int d = 0; d < 64; ++d) {
      float qv[4];
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) qv[qi] = sm.Qs[warp][qi][d];
#pragma unroll
      for (int i = 0; i < SA_MAXL / 32; ++i) {
        const int key = lane + 32 * i;
        if (key < L) {
          const float kv = sm.Kt[d][key];
#pragma unroll
          for (int qi = 0; qi < 4; ++qi) sc[qi][i] = fmaf(qv[qi], kv, sc[qi][i]);
        }
      }
    }
    float inv[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < SA_MAXL / 32; ++i)
        if (lane + 32 * i < L) mx = fmaxf(mx, sc[qi][i]);
      mx = warp_max(mx);
      float den = 0.f;
#pragma unroll
      for (int i = 0; i < SA_MAXL / 32; ++i) {
        const int key = lane + 32 * i;
        if (key < L) {
          const float p = expf(sc[qi][i] - mx);
          den += p;
          sm.Ps[warp][key][qi] = p;
        }
      }
      inv[qi] = 1.0f / warp_sum(den);
    }
    __syncwarp();
    float o[4][2];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) o[qi][0] = o[qi][1] = 0.f;
    for (int key = 0; key < L; ++key) {
      const float4 p = *reinterpret_cast<const float4*>(&sm.Ps[warp][key][0]);
      const float v0 = sm.Vs[key][lane], v1 = sm.Vs[key][lane + 32];
      o[0][0] = fmaf(p.x, v0, o[0][0]); o[0][1] = fmaf(p.x, v1, o[0][1]);
      o[1][0] = fmaf(p.y, v0, o[1][0]); o[1][1] = fmaf(p.y, v1, o[1][1]);
      o[2][0] = fmaf(p.z, v0, o[2][0]); o[2][1] = fmaf(p.z, v1, o[2][1]);
      o[3][0] = fmaf(p.w, v0, o[3][0]); o[3][1] = fmaf(p.w, v1, o[3][1]);
    }
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const int qr = q0 + qi;
      if (qr < L) {
        act_store(out, planes, r0 + qr, h * 64 + lane, o[qi][0] * inv[qi]);
        act_store(out, planes, r0 + qr, h * 64 + lane + 32, o[qi][1] * inv[qi]);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Ragged decoder self-attention on the tensor pipe (tensor-core modes).  CTA = (head, sequence), 8 warps; K and V^T of the
// (sequence, head) are staged ONCE in shared memory as bf16 hi/lo planes, each warp owns 16-query tiles and keeps the whole
// score row block S[16, L<=NT*8] in registers (no online softmax needed at L <= 256): S = Q K^T by mma.sync.m16n8k16
// (bf16 x bf16 -> fp32), fp32 softmax in registers, P re-used directly as the A fragments of P V.  NSPLIT = 2 forms every
// product from hi/lo split operands (lo*hi + hi*lo + hi*hi, fp32 accumulate) like the bf16x3 GEMMs, so the attention keeps
// fp32-grade accuracy; NSPLIT = 1 is the plain bf16 mode.  Nothing is computed or stored for padded frames.
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pack2(float x, float y, uint32_t& hi, uint32_t& lo) { split2_bf16(x, y, hi, lo); }

#define SAT_KLD 72  // bf16 per K row (64 + 8 pad): B-fragment reads are bank-conflict free

template <int NT>
constexpr int sat_smem_bytes(int nsplit) { return nsplit * (NT * 8 * SAT_KLD + 64 * (NT * 8 + 8)) * 2; }

template <int NT, int NSPLIT>
__global__ void __launch_bounds__(256, 1) k_attn_self_tc(const float* __restrict__ qkv, const int* __restrict__ foff,
                                                         Act out, int planes) {
  constexpr int KP = NT * 8;        // padded key capacity
  constexpr int VLD = KP + 8;       // bf16 per V^T row
  pdl_prologue();
  extern __shared__ uint8_t sat_raw[];
  uint32_t* Kh = reinterpret_cast<uint32_t*>(sat_raw);            // [KP][SAT_KLD/2] words
  uint32_t* Vh = Kh + KP * SAT_KLD / 2;                           // [64][VLD/2] words
  uint32_t* Kl = Vh + 64 * VLD / 2;
  uint32_t* Vl = Kl + KP * SAT_KLD / 2;
  const int h = blockIdx.x, b = blockIdx.y;
  const int r0 = foff[b], L = min(foff[b + 1] - r0, KP);
  if (L <= 0) return;
  const int Lp = (L + 15) & ~15;    // keys processed (multiple of 16); keys in [L, Lp) are zero-filled and masked
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // ---- stage K (row-major, d contiguous) and V^T (key contiguous) as bf16 hi/lo
  for (int i = threadIdx.x; i < Lp * 32; i += 256) {
    const int key = i >> 5, dp = i & 31;
    float2 k = make_float2(0.f, 0.f);
    if (key < L) k = *reinterpret_cast<const float2*>(qkv + static_cast<long>(r0 + key) * 768 + 256 + h * 64 + 2 * dp);
    uint32_t hi, lo;
    split_pack2(k.x, k.y, hi, lo);
    Kh[key * (SAT_KLD / 2) + dp] = hi;
    if (NSPLIT > 1) Kl[key * (SAT_KLD / 2) + dp] = lo;
  }
  for (int i = threadIdx.x; i < (Lp / 2) * 64; i += 256) {
    const int d = i & 63, jp = i >> 6;
    const float* vp = qkv + static_cast<long>(r0 + 2 * jp) * 768 + 512 + h * 64 + d;
    const float v0 = (2 * jp < L) ? vp[0] : 0.f;
    const float v1 = (2 * jp + 1 < L) ? vp[768] : 0.f;
    uint32_t hi, lo;
    split_pack2(v0, v1, hi, lo);
    Vh[d * (VLD / 2) + jp] = hi;
    if (NSPLIT > 1) Vl[d * (VLD / 2) + jp] = lo;
  }
  __syncthreads();

  for (int qt = warp; qt * 16 < L; qt += 8) {
    const int q0 = qt * 16;
    // ---- Q fragments (scaled by 1/sqrt(64) = 2^-3: exact, commutes with the split)
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {      // rows g / g+8
        const int qr = q0 + g + 8 * half;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {          // k = 2t / 2t+8
          float2 q = make_float2(0.f, 0.f);
          if (qr < L) q = *reinterpret_cast<const float2*>(qkv + static_cast<long>(r0 + qr) * 768 + h * 64 + kk * 16 + kh * 8 + 2 * t);
          split_pack2(q.x * 0.125f, q.y * 0.125f, qh[kk][half + 2 * kh], ql[kk][half + 2 * kh]);
        }
      }
    }
    // ---- S = Q K^T
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt * 8 < Lp) {
        const uint32_t* kh_ = Kh + (nt * 8 + g) * (SAT_KLD / 2) + t;
        const uint32_t* kl_ = Kl + (nt * 8 + g) * (SAT_KLD / 2) + t;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t bh0 = kh_[kk * 8], bh1 = kh_[kk * 8 + 4];
          if (NSPLIT > 1) {
            const uint32_t bl0 = kl_[kk * 8], bl1 = kl_[kk * 8 + 4];
            mma16816(s[nt], ql[kk], bh0, bh1);
            mma16816(s[nt], qh[kk], bl0, bl1);
          }
          mma16816(s[nt], qh[kk], bh0, bh1);
        }
      }
    }
    // ---- softmax over the L valid keys (rows g and g+8 of the tile)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < Lp) {
        const int key = nt * 8 + 2 * t;
        if (key >= L) s[nt][0] = s[nt][2] = -INFINITY;
        if (key + 1 >= L) s[nt][1] = s[nt][3] = -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < Lp) {
        s[nt][0] = expf(s[nt][0] - mx0);
        s[nt][1] = expf(s[nt][1] - mx0);
        s[nt][2] = expf(s[nt][2] - mx1);
        s[nt][3] = expf(s[nt][3] - mx1);
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // ---- O = P V  (P fragments straight from the S accumulators)
    float o[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < NT / 2; ++ks) {
      if (ks * 16 < Lp) {
        uint32_t ph[4], pl[4];
        split_pack2(s[2 * ks][0], s[2 * ks][1], ph[0], pl[0]);
        split_pack2(s[2 * ks][2], s[2 * ks][3], ph[1], pl[1]);
        split_pack2(s[2 * ks + 1][0], s[2 * ks + 1][1], ph[2], pl[2]);
        split_pack2(s[2 * ks + 1][2], s[2 * ks + 1][3], ph[3], pl[3]);
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
          const uint32_t* vh_ = Vh + (nd * 8 + g) * (VLD / 2) + ks * 8 + t;
          const uint32_t bh0 = vh_[0], bh1 = vh_[4];
          if (NSPLIT > 1) {
            const uint32_t* vl_ = Vl + (nd * 8 + g) * (VLD / 2) + ks * 8 + t;
            const uint32_t bl0 = vl_[0], bl1 = vl_[4];
            mma16816(o[nd], pl, bh0, bh1);
            mma16816(o[nd], ph, bl0, bl1);
          }
          mma16816(o[nd], ph, bh0, bh1);
        }
      }
    }
    // ---- normalise and store (fp32 master and/or bf16 planes)
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int qr = q0 + g + 8 * half;
      if (qr >= L) continue;
      const float inv = half ? inv1 : inv0;
      const long base = static_cast<long>(r0 + qr) * out.ld + h * 64 + 2 * t;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        const float x = o[nd][2 * half] * inv, y = o[nd][2 * half + 1] * inv;
        if (out.f32) *reinterpret_cast<float2*>(out.f32 + base + nd * 8) = make_float2(x, y);
        if (out.pl && planes > 0) {
          uint32_t hi, lo;
          split2_op(x, y, planes, hi, lo);     // the OUTPUT planes follow the mode's operand format (common.cuh op16)
          *reinterpret_cast<uint32_t*>(out.pl + base + nd * 8) = hi;
          if (planes > 1) *reinterpret_cast<uint32_t*>(out.pl + static_cast<long>(out.rows_alloc) * out.ld + base + nd * 8) = lo;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: W [N, K] (row stride ldw) fp32 -> Wt [K][N] fp32 and three 16-bit planes [3][n_pad][K] (zero padded rows):
// fp16 hi | fp16 lo (x3 mode) | bf16 (bf16 mode)
__global__ void k_pack_weight(const float* __restrict__ W, int ldw, int N, int K, int n_pad, float* __restrict__ Wt,
                              op16* __restrict__ pl) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(n_pad) * K) return;
  const int n = i / K, k = i % K;
  float v = 0.f;
  if (n < N) {
    v = W[static_cast<long>(n) * ldw + k];
    Wt[static_cast<long>(k) * N + n] = v;
  }
  op16 hi, lo, b, unused;
  split_op(v, 2, hi, lo);
  split_op(v, 1, b, unused);
  pl[i] = hi;
  pl[static_cast<long>(n_pad) * K + i] = lo;
  pl[2 * static_cast<long>(n_pad) * K + i] = b;
}

// Streaming image of a 256 -> 1024 -> 256 feed-forward weight for k_ffn_swap: the weight ring of that kernel is filled by ONE contiguous
// cp.async.bulk per stage instead of one 2-D tensor box per plane (per-SM ingest from L2: 2-D boxes 41 - 48 B/clk, contiguous 32 KB
// copies 64 B/clk and more: scripts/micro/tma_ingest_probe.cu), so every stage is stored exactly as it has to land in shared memory:
// 16 KB blocks of [128 weight rows x 64 k] 16-bit elements, K-major, 16-byte chunks XOR-swizzled by (row & 7) (what TMA's 128-byte
// swizzle would have written), in the order the kernel consumes them:
//   which = 0 (W1 [1024, 256], phase A): block (rank, g, kb)   = rows rank 256 + g 128 .., k-block kb;         index (rank, g 4 + kb)
//   which = 1 (W2 [256, 1024], phase B): block (rank, e)  with g = (e >> 1) & 1, kb = ((e >> 2) & 1) 2 + (e & 1)
//                                         = output rows g 128 .., k-block rank 4 + kb   (ordered by the availability of the hidden tiles)
// Layout: x3 image [rank][idx][hi | lo] (32 KB stages), then the bf16 image [rank][idx] (16 KB stages).  `pl` = the three planes
// [3][n_pad][K] of k_pack_weight (fp16 hi | fp16 lo | bf16).
__global__ void k_pack_swap_image(const op16* __restrict__ pl, int n_pad, int K, int which, op16* __restrict__ img) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;   // one thread per image element: [3 planes][32 blocks][128][64]
  if (i >= 3L * 32 * 128 * 64) return;
  const int c = i & 63, r = (i >> 6) & 127, blk = (i >> 13) & 31, plane = static_cast<int>(i >> 18);
  const int rank = blk >> 3, e = blk & 7;
  int n, k;
  if (which == 0) {
    n = rank * 256 + (e >> 2) * 128 + r;
    k = (e & 3) * 64 + c;
  } else {
    n = ((e >> 1) & 1) * 128 + r;
    k = rank * 256 + (((e >> 2) & 1) * 2 + (e & 1)) * 64 + c;
  }
  const op16 v = pl[(static_cast<long>(plane) * n_pad + n) * K + k];
  const long in_block = r * 64 + ((((c >> 3) ^ (r & 7)) << 3) | (c & 7));   // element offset inside the 16 KB block
  const long dst = plane < 2 ? (static_cast<long>(blk) * 2 + plane) * 8192 + in_block : 64L * 8192 + static_cast<long>(blk) * 8192 + in_block;
  img[dst] = v;
}

// Weight folding at load time (finalize): C[n, k2] = sum_k1 A[n, k1] * B[k1, k2]  (fp32 in, fp64 accumulate, fp32 out).
// Used to pre-multiply consecutive nn.Linear weights ([out, in] row-major: y = x W^T, so (W_b W_a) applies a then b).
__global__ void k_fold_matmul(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int N, int K1,
                              int K2, float* __restrict__ C, int ldc) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(N) * K2) return;
  const int n = i / K2, k2 = i % K2;
  double acc = 0.0;
  for (int k = 0; k < K1; ++k) acc += static_cast<double>(A[static_cast<long>(n) * lda + k]) * static_cast<double>(B[static_cast<long>(k) * ldb + k2]);
  C[static_cast<long>(n) * ldc + k2] = static_cast<float>(acc);
}
// out[n] = sum_k A[n, k] * v[k] + (b ? b[n] : 0)
__global__ void k_fold_matvec(const float* __restrict__ A, int lda, const float* __restrict__ v, const float* __restrict__ b,
                              int N, int K, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  double acc = b ? static_cast<double>(b[n]) : 0.0;
  for (int k = 0; k < K; ++k) acc += static_cast<double>(A[static_cast<long>(n) * lda + k]) * static_cast<double>(v[k]);
  out[n] = static_cast<float>(acc);
}
// C[n, k2] = sum_k1 A[k1, n] * B[k1, k2]   (A^T B; fp64 accumulate) -- folds a query projection into the keys it meets
__global__ void k_fold_tn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int N, int K1, int K2,
                          float* __restrict__ C, int ldc) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(N) * K2) return;
  const int n = i / K2, k2 = i % K2;
  double acc = 0.0;
  for (int k = 0; k < K1; ++k) acc += static_cast<double>(A[static_cast<long>(k) * lda + n]) * static_cast<double>(B[static_cast<long>(k) * ldb + k2]);
  C[static_cast<long>(n) * ldc + k2] = static_cast<float>(acc);
}
// dst[r, c] = (r == c) : rows x rows identity block written into a wider matrix
__global__ void k_set_identity(float* __restrict__ dst, int ldd, int n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(n) * n) return;
  const int r = i / n, c = i % n;
  dst[static_cast<long>(r) * ldd + c] = (r == c) ? 1.f : 0.f;
}
// strided 2-D copy of fp32 blocks (assembling folded weight matrices)
__global__ void k_copy2d(const float* __restrict__ src, int lds, int rows, int cols, float* __restrict__ dst, int ldd) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long>(rows) * cols) return;
  const int r = i / cols, c = i % cols;
  dst[static_cast<long>(r) * ldd + c] = src[static_cast<long>(r) * lds + c];
}

// deterministic pseudo-random fill in [-scale, scale] (benchmark operands)
__global__ void k_fill_pseudo(float* __restrict__ p, long n, float scale, unsigned seed) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned x = static_cast<unsigned>(i) * 2654435761u + seed * 40503u;
  x ^= x >> 16; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
  p[i] = scale * (static_cast<float>(x & 0xFFFFFF) / 8388608.0f - 1.0f);
}

// ------------------------------------------------------------------------------------------------
// feats2joints: de-normalise + recover_from_ric (data/HumanML3D.py:44-48;
// data/humanml/scripts/motion_process.py:355-381,415-430; quaternion.py:16-20,54-73).  One CTA per sequence:
// the two cumulative sums over frames (root yaw, root XZ) are block scans, then every (frame, joint) is independent.
__global__ void __launch_bounds__(256) k_feats2joints(const float* __restrict__ feats, const float* __restrict__ mean,
                                                      const float* __restrict__ stdv, int L, int nfeats, int njoints,
                                                      float* __restrict__ joints) {
  __shared__ float ang[SA_MAXL], px[SA_MAXL], pz[SA_MAXL], py[SA_MAXL];
  const int b = blockIdx.x;
  const float* f = feats + static_cast<long>(b) * L * nfeats;
  auto feat = [&](int t, int c) { return f[static_cast<long>(t) * nfeats + c] * stdv[c] + mean[c]; };
  // r_rot_ang[t] = sum_{u<t} rot_vel[u]  (sequential in fp32 to follow torch.cumsum's left-to-right order)
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int t = 0; t < L; ++t) {
      if (t > 0) a += feat(t - 1, 0);
      ang[t] = a;
    }
  }
  __syncthreads();
  // rotated root velocity, then cumsum
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    float vx = 0.f, vz = 0.f;
    if (t > 0) {
      vx = feat(t - 1, 1);
      vz = feat(t - 1, 2);
    }
    // qrot with q_inv = (cos a, 0, -sin a, 0), v = (vx, 0, vz)
    const float w = cosf(ang[t]), qy = -sinf(ang[t]);
    const float uvx = qy * vz, uvz = -qy * vx;          // cross((0,qy,0), v)
    const float uuvx = qy * uvz, uuvz = -qy * uvx;      // cross((0,qy,0), uv)
    px[t] = vx + 2.f * (w * uvx + uuvx);
    pz[t] = vz + 2.f * (w * uvz + uuvz);
    py[t] = feat(t, 3);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sx = 0.f, sz = 0.f;
    for (int t = 0; t < L; ++t) {
      sx += px[t];
      sz += pz[t];
      px[t] = sx;
      pz[t] = sz;
    }
  }
  __syncthreads();
  float* out = joints + static_cast<long>(b) * L * njoints * 3;
  for (int i = threadIdx.x; i < L * njoints; i += blockDim.x) {
    const int t = i / njoints, j = i % njoints;
    float x, y, z;
    if (j == 0) {
      x = px[t]; y = py[t]; z = pz[t];
    } else {
      const float vx = feat(t, 4 + (j - 1) * 3), vy = feat(t, 5 + (j - 1) * 3), vz = feat(t, 6 + (j - 1) * 3);
      const float w = cosf(ang[t]), qy = -sinf(ang[t]);
      const float uvx = qy * vz, uvz = -qy * vx;
      const float uuvx = qy * uvz, uuvz = -qy * uvx;
      x = vx + 2.f * (w * uvx + uuvx) + px[t];
      y = vy;
      z = vz + 2.f * (w * uvz + uuvz) + pz[t];
    }
    out[static_cast<long>(i) * 3 + 0] = x;
    out[static_cast<long>(i) * 3 + 1] = y;
    out[static_cast<long>(i) * 3 + 2] = z;
  }
}
