// Fused feed-forward pairs of the denoiser layer, "swapped" formulation for the latency-bound batch sizes (R <= 1776 rows;
// above that the four separate fused linears are faster -- measured, profiles/r02d_large_batch.txt -- and are what the plans use):
// the WEIGHT rows sit on the 128 lanes of the tcgen05 M axis and a small group of RT <= 48 tokens on the N axis
//       h^T  [256-slice x RT]  = W1[slice, :]   . X^T          (phase A, two 128-row M tiles)
//       out^T[256       x RT]  = W2[:, slice]   . h[:, slice]^T (phase B, K-split partial)
// so a cluster of only CL = 4 CTAs owns RT tokens (not a 128-row GEMM tile) and 1280 rows spread over 27 clusters /
// 108 SMs (an earlier variant with 128-row tiles on clusters of 8 CTAs took 39 us per layer against 24 us).  What bounds the fused FFN on B200 is the distributed-shared-memory exchange (~16 B/clk per SM measured,
// scripts/micro/dsmem_probe.cu): its volume is  tokens x 1 KB  per CTA whatever the cluster size, so the token group is made
// small (48 KB per exchange instead of 128 KB) and paid for with weight re-streaming from L2 (512 KB per pair and CTA through a
// 4-stage TMA ring that runs ahead across the exchanges), which the 126 MB L2 serves at > 80 B/clk per SM.
//
//   pair 0 :  x3 = LN( x1 + W2 relu(W1 x1 + b1) + b2 ) g + b + delta[add_idx[row]]      (mdiff_transformer.py:60-62 + hoisted ca_block)
//   pair 1 :  s  = SiLU( LN( W2' gelu(W1' x3 + b1') + b2' ) (1 + scale) + shift )       (mdiff_transformer.py:137-162,248-262)
//
// Accumulators are transposed (lane = feature, column = token): the hidden activation is written as the next B operand
// (token-major, K-major 128B-swizzled bf16 hi/lo planes) straight from the TMEM lanes; the K-split partials are reduce-scattered
// by tokens (CTA r owns RT/4 tokens; 32 lanes store 128 contiguous bytes per remote instruction; fixed summation order ->
// deterministic); owners do bias / residual / LayerNorm / modulation with one warp per token and broadcast x3 as the next X operand.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"
#include "linear.cuh"
#include "tc_ptx.cuh"

struct FfnArgs {
  int M_max;
  const int* M_dev;
  int npairs;            // 1 or 2
  int act[2];            // EPI_RELU / EPI_GELU of the hidden layer
  int kind[2];           // EPI_LN (+res, +addv) or EPI_LN_MOD_SILU
  const float* b1[2];    // [1024]
  const float* b2[2];    // [256]
  const float* ln_g[2];
  const float* ln_b[2];
  const float* res;      // pair 0 residual (fp32 rows of X), ld 256; may be null
  const float* addv;     // EPI_LN: optional broadcast add
  const int* add_idx;
  int ld_add;
  const float* mod[2];   // EPI_LN_MOD_SILU: [scale(256) | shift(256)]
  Act out[2];
  int out_planes;
  int x_plane_rows;      // row offset of the lo plane in the X tensor map
  int rt;                // k_ffn_swap: tokens per cluster (16 / 32 / 48)
  // k_ffn_swap with the sa_block attention fused in front (rt == 48 only): instead of loading X by TMA, every CTA computes
  //   x1[row] = LN( Xin[row] + b_o + sum_h sum_j softmax_j(q_h . k_hj / 8) v'_hj )   (see k_attn_ln in kernels.cuh)
  // for its 12 owned rows from the extended in-projection buffer and broadcasts it as the X operand of the cluster
  int att;                         // 1: fused attention prologue
  const float* att_qkvx;           // [rows, 1792]: q | k | 4 x v' | X
  const int* att_off;              // [S + 1] row offsets of the sequences
  const int* att_row_seq;          // [rows] sequence of a row
  const float* att_textkv;         // per-sequence conditioning row: k | 4 x v'
  int att_ld_textkv;
  const float* att_timekv;         // this (step, layer)'s time-token row: k | 4 x v'
  const float* att_res;            // layer input rows (residual), ld att_ld_res
  int att_ld_res;
  const float *att_bo, *att_g, *att_b;   // out_proj bias, norm1 weight / bias
  Act att_x1;                      // fp32 master of x1 (residual of pair 0)
  Act att_xcopy;                   // optional copy of the layer input (U-Net skip source), fp32 and/or planes
  int w1_plane_rows[2], w2_plane_rows[2];
  const uint8_t* w1_img[2];   // streaming images of the pair's weights in this mode (k_pack_swap_image): [rank][8 stages][STG bytes]
  const uint8_t* w2_img[2];
  unsigned long long* trace;
  long long* dbg;        // optional per-CTA clock64 stamps [ncta][128] (ladiff_ffn_test with LADIFF_DBG_STAMPS=1)
};

template <int NSPLIT>
struct SwapCfg {
  static constexpr int CL = 4, D = 256, FF = 1024, HS = FF / CL, BK = 64;
  static constexpr int U = 128 * BK * 2;       // one [128 x 64] bf16 weight block (16 KB)
  static constexpr int STG = NSPLIT * U;       // ring stage: all planes of one weight block
  static constexpr int RT_MAX = 48;
  static constexpr int EPI_WARPS = 8, THREADS = 64 + EPI_WARPS * 32;
  // x3 mode: the B operand of one MMA is [x_hi ; x_lo] (the two planes are adjacent in shared memory, same 128-byte row pitch), so
  // an accumulator tile is AW = 2 RT columns wide: [0, RT) = w_hi.x_hi + w_lo.x_hi, [RT, 2 RT) = w_hi.x_lo; summed by its reader.
  // Two MMAs per k-step instead of three, and the 4 KB weight tile is fetched from shared memory twice instead of three times
  // (the N = 48 MMAs are bound by that fetch, not by the tensor pipe: scripts/micro/mma_small_n_probe.cu).
  static constexpr int TMEM_COLS = NSPLIT == 2 ? 512 : 256;   // acc A: [0, 2 AW)   acc B: [2 AW, 4 AW)
  static constexpr int SMEM_LIMIT = 227 * 1024;
  static constexpr int TAIL = 512 /*barriers*/ + 1024 /*alignment slack*/;
  __host__ __device__ static constexpr int op_bytes(int rt) { return 4 * NSPLIT * rt * 128; }   // [kb][plane][rt x 128 B]
  __host__ __device__ static constexpr int hr_bytes(int rt) { return rt * D * 4 > op_bytes(rt) ? rt * D * 4 : op_bytes(rt); }
  __host__ __device__ static constexpr int stages(int rt) {
    const int n = (SMEM_LIMIT - TAIL - op_bytes(rt) - hr_bytes(rt)) / STG;
    return n > 4 ? 4 : n;
  }
  __host__ __device__ static constexpr int smem_bytes(int rt) { return op_bytes(rt) + hr_bytes(rt) + stages(rt) * STG + TAIL; }
  static_assert(stages(RT_MAX) == 4 && stages(16) == 4, "the ring depth is a compile-time 4 (slot = stage & 3)");
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// cluster barrier without release / acquire fences: pure control dependency (what it orders was observed through an mbarrier)
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v2u(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void st_shared_v2u(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// generic-proxy writes (st.shared / st.shared::cluster) -> visible to the async proxy (tcgen05.mma operand reads), all state spaces
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_v4u(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// sa_block attention + out-proj (folded into the values) + residual + LayerNorm for the 12 rows this CTA owns, executed by the
// 256 epilogue threads (mdiff_transformer.py:54-62,296-313; same arithmetic as k_attn_ln, kernels.cuh).  Thread (cp, hs):
// output columns 2 cp, 2 cp + 1 of the owned rows 6 hs .. 6 hs + 5.  `scratch` is the (still unused) h / receive region:
//   Ks[20][256] latent key rows of every touched sequence | Ts[12][256] text-token keys | Tm[256] time-token key |
//   Qs[12][256] (pre-scaled by 1/8) | Ps[12][4][8] | red[2][8][6]
// x1 goes (i) as fp32 to p.att_x1 (the residual the owners of pair 0 read back), (ii) as bf16 hi/lo planes into the X operand of
// all four CTAs of the cluster (token-major, K-major, 128-byte swizzle).
template <int NSPLIT>
__device__ __forceinline__ void swap_attention_prologue(const FfnArgs& p, uint8_t* scratch, uint32_t xop_local, int XT, int row0,
                                                        int rank, int tpc, int M, int et) {
  constexpr int MAXT = 5, NK = MAXT + 2, OWN = 12, HALF = 6;
  float* Ks = reinterpret_cast<float*>(scratch);   // [20][256]
  float* Ts = Ks + 20 * 256;                       // [12][256]
  float* Tm = Ts + 12 * 256;                       // [256]
  float* Qs = Tm + 256;                            // [12][256]
  float* Ps = Qs + 12 * 256;                       // [12][4][8]
  float* red = Ps + 12 * 4 * 8;                    // [2][8][6]
  const int R0 = row0 + rank * tpc;                // first owned row
  const int nown = max(0, min(OWN, M - R0));       // valid owned rows (tpc == 12)
  const int lane = et & 31, wrp = et >> 5;
  const int cp = et & 127, hs = et >> 7, c = 2 * cp;
  if (nown > 0) {
    const int s0 = __ldg(p.att_row_seq + R0), s1 = __ldg(p.att_row_seq + R0 + nown - 1);
    const int Rk0 = __ldg(p.att_off + s0), Rk1 = __ldg(p.att_off + s1 + 1);   // latent key rows [Rk0, Rk1): at most 20
    // ---- stage keys / queries: thread = column
    for (int j = 0; j < Rk1 - Rk0; ++j) Ks[j * 256 + et] = p.att_qkvx[static_cast<long>(Rk0 + j) * DQX_LD + 256 + et];
    for (int s = s0; s <= s1; ++s) Ts[(s - s0) * 256 + et] = p.att_textkv[static_cast<long>(s) * p.att_ld_textkv + et];
    Tm[et] = p.att_timekv[et];
    for (int i = 0; i < nown; ++i) Qs[i * 256 + et] = 0.125f * p.att_qkvx[static_cast<long>(R0 + i) * DQX_LD + et];
  }
  // per owned row of this thread's half: sequence, first row of the sequence, number of latent keys
  int rs_[HALF], rr0[HALF], rm[HALF];
  float2 xr[HALF];
#pragma unroll
  for (int u = 0; u < HALF; ++u) {
    const int i = hs * HALF + u;
    rs_[u] = -1; rr0[u] = 0; rm[u] = 0;
    xr[u] = make_float2(0.f, 0.f);
    if (i < nown) {
      const int s = __ldg(p.att_row_seq + R0 + i);
      rs_[u] = s;
      rr0[u] = __ldg(p.att_off + s);
      rm[u] = min(__ldg(p.att_off + s + 1) - rr0[u], MAXT);
      xr[u] = *reinterpret_cast<const float2*>(p.att_res + static_cast<long>(R0 + i) * p.att_ld_res + c);
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (et == 0) trace_mark(p.trace, 4);
  if (nown > 0) {
    const int s0 = __ldg(p.att_row_seq + R0);
    const int Rk0 = __ldg(p.att_off + s0);
    // ---- scores: element e = (i, h, j), one 64-long dot product each (d rotated by the lane: no bank conflicts)
    for (int e = et; e < nown * 4 * 8; e += 256) {
      const int j = e & 7, h = (e >> 3) & 3, i = e >> 5;
      const int s = __ldg(p.att_row_seq + R0 + i);
      const int r0 = __ldg(p.att_off + s), m = min(__ldg(p.att_off + s + 1) - r0, MAXT);
      float sc = -INFINITY;
      if (j < NK && (j < m || j >= MAXT)) {
        const float* kp = (j < MAXT ? Ks + (r0 - Rk0 + j) * 256 : (j == MAXT ? Ts + (s - s0) * 256 : Tm)) + h * 64;
        const float* qp = Qs + i * 256 + h * 64;
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
      Ps[e] = sc;
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (et < nown * 4) {   // softmax over the 7 keys of (row, head)
    float* pr_ = Ps + et * 8;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NK; ++j) mx = fmaxf(mx, pr_[j]);
    float pj[NK], den = 0.f;
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      pj[j] = expf(pr_[j] - mx);
      den += pj[j];
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < NK; ++j) pr_[j] = pj[j] * inv;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (et == 0) trace_mark(p.trace, 5);
  // ---- out[i, c..c+1] = Xin + b_o + sum_h sum_j P[i][h][j] v'[h][j][c..c+1]; v' of a sequence is loaded once per run of rows
  float2 acc[HALF];
  {
    const float2 bo = *reinterpret_cast<const float2*>(p.att_bo + c);
    float2 v[4][NK];
    int cur = -2;
#pragma unroll
    for (int u = 0; u < HALF; ++u) {
      acc[u] = make_float2(bo.x + xr[u].x, bo.y + xr[u].y);
      if (rs_[u] >= 0) {
        if (rs_[u] != cur) {
          cur = rs_[u];
          const float* tk = p.att_textkv + static_cast<long>(cur) * p.att_ld_textkv;
#pragma unroll
          for (int j = 0; j < NK; ++j) {
            const bool on = j < rm[u] || j >= MAXT;
            const float* base = j < MAXT ? p.att_qkvx + static_cast<long>(rr0[u] + j) * DQX_LD + 512 : (j == MAXT ? tk + 256 : p.att_timekv + 256);
#pragma unroll
            for (int h = 0; h < 4; ++h) v[h][j] = on ? *reinterpret_cast<const float2*>(base + h * 256 + c) : make_float2(0.f, 0.f);
          }
        }
        const float* pp = Ps + (hs * HALF + u) * 32;
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
          for (int j = 0; j < NK; ++j) {
            const float w = pp[h * 8 + j];   // masked keys: P == 0 (score -inf) and v == 0
            acc[u].x = fmaf(w, v[h][j].x, acc[u].x);
            acc[u].y = fmaf(w, v[h][j].y, acc[u].y);
          }
      }
    }
  }
  if (et == 0) trace_mark(p.trace, 6);
  // ---- LayerNorm over the 256 columns of each row: 128 threads (4 warps) of this half hold them
  float mean[HALF];
#pragma unroll
  for (int u = 0; u < HALF; ++u) {
    const float w = warp_sum(acc[u].x + acc[u].y);
    if (lane == 0) red[(0 * 8 + wrp) * HALF + u] = w;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
  for (int u = 0; u < HALF; ++u) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) t += red[(0 * 8 + hs * 4 + w) * HALF + u];
    mean[u] = t * (1.f / 256.f);
    const float dx = acc[u].x - mean[u], dy = acc[u].y - mean[u];
    const float w = warp_sum(dx * dx + dy * dy);
    if (lane == 0) red[(1 * 8 + wrp) * HALF + u] = w;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float2 gc = *reinterpret_cast<const float2*>(p.att_g + c), bc = *reinterpret_cast<const float2*>(p.att_b + c);
#pragma unroll
  for (int u = 0; u < HALF; ++u) {
    const int i = hs * HALF + u;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) t += red[(1 * 8 + hs * 4 + w) * HALF + u];
    const float rstd = 1.0f / sqrtf(t * (1.f / 256.f) + LD_EPS);
    const bool valid = i < nown;
    const float y0 = valid ? (acc[u].x - mean[u]) * rstd * gc.x + bc.x : 0.f;
    const float y1 = valid ? (acc[u].y - mean[u]) * rstd * gc.y + bc.y : 0.f;
    uint32_t hi, lo;
    split2_op<NSPLIT>(y0, y1, hi, lo);
    if (valid) {
      const long row = R0 + i;
      *reinterpret_cast<float2*>(p.att_x1.f32 + row * p.att_x1.ld + c) = make_float2(y0, y1);
      if (p.att_xcopy.f32) *reinterpret_cast<float2*>(p.att_xcopy.f32 + row * p.att_xcopy.ld + c) = xr[u];
      if (p.att_xcopy.pl && p.out_planes > 0) {
        uint32_t xh, xl;
        split2_op(xr[u].x, xr[u].y, p.out_planes, xh, xl);
        op16* dh = p.att_xcopy.pl + row * p.att_xcopy.ld + c;
        *reinterpret_cast<uint32_t*>(dh) = xh;
        if (p.out_planes > 1) *reinterpret_cast<uint32_t*>(dh + static_cast<long>(p.att_xcopy.rows_alloc) * p.att_xcopy.ld) = xl;
      }
    }
    // X operand of every CTA of the cluster: token tcl, k-block c / 64, 4 bytes inside swizzle chunk (c & 63) / 8
    const int tcl = rank * tpc + i;
    const uint32_t off = (c >> 6) * NSPLIT * XT + (tcl >> 3) * 1024 + (tcl & 7) * 128 + (((((c & 63) >> 3) ^ (tcl & 7)) << 4) | ((c & 7) * 2));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t base = tc::mapa(xop_local, k) + off;
      asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(base), "r"(hi) : "memory");
      if (NSPLIT == 2) asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(base + XT), "r"(lo) : "memory");
    }
  }
}

// ATT: the opt-in attention prologue is a template parameter so that the default instantiation carries none of its code
// (as a run-time branch it cost the FFN kernel 1.8 us per launch through register allocation alone).
template <int NSPLIT, bool ATT>
__global__ void __cluster_dims__(1, 4, 1) __launch_bounds__(SwapCfg<NSPLIT>::THREADS, 1)
k_ffn_swap(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1a,
           const __grid_constant__ CUtensorMap tmW2a, const __grid_constant__ CUtensorMap tmW1b,
           const __grid_constant__ CUtensorMap tmW2b, const FfnArgs p) {
  using C = SwapCfg<NSPLIT>;
  constexpr int U = C::U, STG = C::STG;
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int rt = p.rt, tpc = rt >> 2;       // tokens per cluster / owned per CTA
  const int aw = NSPLIT == 2 ? 2 * rt : rt; // accumulator tile width (TMEM columns)
  const int row0 = blockIdx.x * rt;
  tc::pdl_launch_dependents();
  if (row0 >= M) return;  // cluster-uniform
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) trace_mark(p.trace, 0);
  long long* dbg = p.dbg ? p.dbg + (blockIdx.x * C::CL + rank) * 128 : nullptr;
#define SSTAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) SSTAMP(0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int XT = rt * 128;                  // one operand tile: rt tokens x 64 bf16
  constexpr int nst = 4;
  uint8_t* xop = smem;                      // X operand  [kb][plane][rt x 128 B]
  uint8_t* hr = smem + C::op_bytes(rt);     // h operand (same layout over the hidden slice); later the receive buffer
  uint8_t* ring = hr + C::hr_bytes(rt);     // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + nst * STG);
  uint64_t* full = bars;                    // [4]
  uint64_t* empty = bars + 4;               // [4]
  uint64_t* x_full = bars + 8;              // [4]
  uint64_t* acca = bars + 12;               // [2]  hidden M tile g accumulated
  uint64_t* hfull = bars + 14;              // [2]  256 arrivals: h operand k-blocks 2g, 2g+1 written
  uint64_t* accb = bars + 16;               // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  // warp index through a shuffle: provably warp-uniform, so the MMA / TMA operands of the role branches stay in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int total = p.npairs * 16;          // weight stages: per pair 8 of W1 (g, kb) then 8 of W2 (g, kb)

  auto issue_load = [&](int j) {
    // stage r of a pair: r < 8 phase A (M tile g = r >> 2, k-block r & 3); r >= 8 phase B ordered by the availability of h:
    // first every (out tile, k-block) that reads hidden M tile 0 (k-blocks 0, 1), then those reading tile 1 (k-blocks 2, 3).
    // The streaming image holds the stages of a (weight, rank) in exactly that order: one contiguous bulk copy per stage.
    const int slot = j & 3, pr = j >> 4, r = j & 15;
    const uint8_t* src = (r < 8 ? p.w1_img[pr] : p.w2_img[pr]) + (static_cast<size_t>(rank) * 8 + (r & 7)) * STG;
    tc::mbar_expect_tx(&full[slot], STG);
    tc::bulk_load_1d(ring + slot * STG, src, STG, &full[slot]);
  };

  if (warp == 0) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmX);
      for (int i = 0; i < 4; ++i) {
        tc::mbar_init(&full[i], 1);
        tc::mbar_init(&empty[i], 1);
        tc::mbar_init(&x_full[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&acca[i], 1);
        tc::mbar_init(&hfull[i], C::EPI_WARPS * 32);
      }
      tc::mbar_init(accb, 1);
      tc::fence_barrier_init();
      tc::fence_proxy_async();
    }
    __syncwarp();
    // weights never depend on the previous grid: fill the ring before the dependency wait
    if (tc::elect_one())
      for (int j = 0; j < nst && j < total; ++j) issue_load(j);
    __syncwarp();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: X once, then the weight stages; runs ahead of the exchanges by the depth of the ring =====
    int next_load = nst < total ? nst : total;
    tc::pdl_wait();
    if (lane == 0) {
      trace_mark(p.trace, 1);
      SSTAMP(1);
    }
    if (!ATT && tc::elect_one()) {
      for (int kb = 0; kb < 4; ++kb) {
        tc::mbar_expect_tx(&x_full[kb], NSPLIT * XT);
        for (int pl = 0; pl < NSPLIT; ++pl)   // tmX has a box of rt rows
          tc::tma_load_2d(xop + (kb * NSPLIT + pl) * XT, &tmX, &x_full[kb], kb * C::BK, pl * p.x_plane_rows + row0);
      }
    }
    __syncwarp();
    if (ATT) tc::cluster_sync();  // #0: x1 has been broadcast by the attention prologue of every CTA
    for (int pr = 0; pr < p.npairs; ++pr) {
      // every load whose slot is freed by MMAs of pairs <= pr can be issued before this pair's cluster barriers
      const int lim = min(total, 16 * (pr + 1) + nst);
      for (; next_load < lim; ++next_load) {
        const int prev = next_load - nst;
        tc::mbar_wait(&empty[prev & 3], (prev >> 2) & 1);
        if (dbg && next_load < 24 && lane == 0) dbg[80 + next_load] = clock64();
        if (tc::elect_one()) issue_load(next_load);
        __syncwarp();
      }
      cluster_sync_relaxed();  // #1
      tc::cluster_sync();  // #2
      if (pr + 1 < p.npairs) tc::cluster_sync();  // #3
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =====
    const uint32_t idesc = tc::idesc_op<NSPLIT>(128, rt), idesc_cat = tc::idesc_op<NSPLIT>(128, 2 * rt);
    const uint32_t ring_u = tc::smem_u32(ring), xop_u = tc::smem_u32(xop), hr_u = tc::smem_u32(hr);
    const uint32_t tmem_u = tmem_base;
    if (ATT) {
      tc::cluster_sync();  // #0
      fence_proxy_async_all();
    }
    for (int pr = 0; pr < p.npairs; ++pr) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int slot = r & 3, ph = r >> 3;   // 16 stages per pair: slot = (16 pr + r) & 3; stage order: see issue_load
        const int g = ph ? (r >> 1) & 1 : (r >> 2) & 1, kb = ph ? ((r >> 2) & 1) * 2 + (r & 1) : r & 3;
        if (ph == 0) {
          if (pr == 0 && !ATT) tc::mbar_wait(&x_full[kb], 0);
        } else if ((r & 3) == 0) {
          tc::mbar_wait(&hfull[kb >> 1], pr & 1);
        }
        if (dbg && pr == 0 && lane == 0) dbg[48 + 2 * r] = clock64();
        tc::mbar_wait(&full[slot], (r >> 2) & 1);   // fill number 4 pr + (r >> 2) of this slot
        tc::tc_fence_after();
        if (dbg && pr == 0 && lane == 0) dbg[48 + 2 * r + 1] = clock64();
        if (lane == 0) {
          if (r == 0) SSTAMP(2 + 20 * pr);
          if (r == 8) SSTAMP(3 + 20 * pr);
        }
        const uint32_t d = tmem_u + (ph ? 2 * aw : 0) + g * aw;
        const uint32_t sw = ring_u + slot * STG;
        const uint32_t sx = (ph ? hr_u : xop_u) + kb * NSPLIT * XT;
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t koff = kk * 32;
            const uint64_t w_hi = tc::smem_desc_sw128(sw + koff), x_hi = tc::smem_desc_sw128(sx + koff);
            const uint32_t acc0 = (kb | kk) != 0;
            if (NSPLIT == 1) {
              tc::mma_bf16_ss(d, w_hi, x_hi, idesc, acc0);
            } else {
              const uint64_t w_lo = tc::smem_desc_sw128(sw + U + koff);
              tc::mma_bf16_ss(d, w_hi, x_hi, idesc_cat, acc0);   // N = 2 rt: x_hi rows, then the x_lo plane right behind them
              tc::mma_bf16_ss(d, w_lo, x_hi, idesc, 1u);
            }
          }
          tc::mma_commit(&empty[slot]);
          if (ph == 0 && kb == 3) tc::mma_commit(&acca[g]);
          if (r == 15) tc::mma_commit(accb);
        }
        __syncwarp();
      }
      if (lane == 0) SSTAMP(4 + 20 * pr);
      cluster_sync_relaxed();  // #1
      tc::cluster_sync();  // #2
      if (pr + 1 < p.npairs) {
        tc::cluster_sync();  // #3: the next X operand has been written by the owners (generic proxy, remote CTAs)
        fence_proxy_async_all();
      }
    }
  } else {
    // ===== epilogue warps =====
    const int e = warp - 2;                // 0..7
    const int g = e >> 2;                  // M tile handled by this warp
    const int q = warp & 3;                // TMEM lane quarter of this warp
    const uint32_t tlane = static_cast<uint32_t>(q * 32) << 16;
    const int f = g * 128 + q * 32 + lane; // feature (hidden-slice feature in phase A, output feature in phase B)
    const uint32_t recv_local = tc::smem_u32(hr), xop_local = tc::smem_u32(xop);
    // owner phase: one half-warp per owned token (tpc <= 12 tokens, 16 half-warps); lane l of the half-warp owns the features
    // 4 l + 64 j .. + 3 (j = 0..3: one float4 per k-block -> conflict-free shared loads, 128-byte coalesced global stores)
    const int otl = e * 2 + (lane >> 4), c0 = (lane & 15) * 4;
    const bool own = otl < tpc;
    const int otcl = static_cast<int>(rank) * tpc + otl;      // token inside the cluster's group
    const long orow = static_cast<long>(row0) + otcl;
    const bool ovalid = own && orow < M;
    tc::pdl_wait();
    if (ATT) {
      swap_attention_prologue<NSPLIT>(p, hr, xop_local, XT, row0, static_cast<int>(rank), tpc, M, threadIdx.x - 64);
      if (threadIdx.x == 64) trace_mark(p.trace, 7);
      fence_proxy_async_all();
      tc::cluster_sync();  // #0: every CTA's x1 rows have landed in every X operand; x1 fp32 (global) is visible to the owners
    }
    // owner-side operands straight into registers while the mainloop runs: va = residual (pair 0) / 1 + scale (pair 1),
    // vb = hoisted ca_block delta (pair 0) / shift (pair 1)
    float va[16], vb[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) va[k] = vb[k] = 0.f;
    if (ovalid && p.kind[0] == EPI_LN) {
      if (p.res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4* rp = reinterpret_cast<const float4*>(p.res + orow * C::D + c0 + 64 * j);
          const float4 a = ATT ? __ldcg(rp) : *rp;   // ATT: written by this CTA a moment ago -> read through L2
          va[4 * j] = a.x; va[4 * j + 1] = a.y; va[4 * j + 2] = a.z; va[4 * j + 3] = a.w;
        }
      }
      if (p.addv) {
        const float* ap = p.addv + static_cast<long>(__ldg(p.add_idx + orow)) * p.ld_add + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = *reinterpret_cast<const float4*>(ap + 64 * j);
          vb[4 * j] = a.x; vb[4 * j + 1] = a.y; vb[4 * j + 2] = a.z; vb[4 * j + 3] = a.w;
        }
      }
    }

    for (int pr = 0; pr < p.npairs; ++pr) {
      const uint32_t par = pr & 1;
      const bool ln_res = p.kind[pr] == EPI_LN;
      const bool last = pr + 1 >= p.npairs;
      // (1) hidden activation: accA^T -> + b1 -> act -> bf16 hi/lo planes of the phase-B operand (token-major, K-major swizzled).
      // All 8 warps work on M tile 0 (while the MMAs of tile 1 run), then on tile 1: warp (q, hsel) owns TMEM lanes 32 q .. and
      // the token half hsel, so the part of this epilogue that is exposed after the last phase-A MMA is half a tile.
      {
        const int hsel = e >> 2, tok0 = hsel * (rt >> 1), nch = rt >> 4;   // rt / 2 tokens in chunks of 8
        const bool relu = p.act[pr] == EPI_RELU;
#pragma unroll 1
        for (int gt = 0; gt < 2; ++gt) {
          const int fh = gt * 128 + q * 32 + lane;     // hidden feature inside this CTA's slice
          const float b1 = __ldg(p.b1[pr] + rank * C::HS + fh);
          tc::mbar_wait(&acca[gt], par);
          tc::tc_fence_after();
          if (pr == 0 && gt == 0 && e == 0 && lane == 0) trace_mark(p.trace, 2);
          if (gt == 0 && e == 0 && lane == 0) SSTAMP(8 + 20 * pr);
          const int kbh = fh >> 6, within = fh & 63, chunk = within >> 3;
          uint8_t* hb = hr + (kbh * NSPLIT) * XT + (within & 7) * 2;
          for (int c = 0; c < nch; ++c) {
            float v[8];
            const int t0 = tok0 + c * 8;               // multiple of 8: token t0 + k sits in row k of 8-row group t0 / 8
            tmem_ld8(tmem_base + tlane + gt * aw + t0, v);
            if (NSPLIT == 2) {
              float v2[8];
              tmem_ld8(tmem_base + tlane + gt * aw + rt + t0, v2);
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] += v2[k];
            }
            uint8_t* hrow = hb + (t0 >> 3) * 1024;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float x = v[k] + b1;
              x = relu ? fmaxf(x, 0.f) : gelu_erf_fast(x);
              op16 hi, lo;
              split_op(x, NSPLIT, hi, lo);
              const int off = k * 128 + ((chunk ^ k) << 4);
              *reinterpret_cast<op16*>(hrow + off) = hi;
              if (NSPLIT == 2) *reinterpret_cast<op16*>(hrow + XT + off) = lo;
            }
          }
          tc::fence_proxy_async();
          tc::mbar_arrive(&hfull[gt]);
        }
      }
      if (e == 0 && lane == 0) SSTAMP(9 + 20 * pr);
      // per-feature vectors of the owner phase (weights: L2 hits), requested while phase B runs
      float4 b2v[4], gv[4], bv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        b2v[j] = __ldg(reinterpret_cast<const float4*>(p.b2[pr] + c0 + 64 * j));
        gv[j] = __ldg(reinterpret_cast<const float4*>(p.ln_g[pr] + c0 + 64 * j));
        bv[j] = __ldg(reinterpret_cast<const float4*>(p.ln_b[pr] + c0 + 64 * j));
      }
      if (!ln_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 s4 = *reinterpret_cast<const float4*>(p.mod[pr] + c0 + 64 * j);
          const float4 h4 = *reinterpret_cast<const float4*>(p.mod[pr] + C::D + c0 + 64 * j);
          va[4 * j] = 1.f + s4.x; va[4 * j + 1] = 1.f + s4.y; va[4 * j + 2] = 1.f + s4.z; va[4 * j + 3] = 1.f + s4.w;
          vb[4 * j] = h4.x; vb[4 * j + 1] = h4.y; vb[4 * j + 2] = h4.z; vb[4 * j + 3] = h4.w;
        }
      }
      // (2) partial outputs -> owners (tokens tpc*k .. of the group belong to CTA k); 32 lanes = 128 contiguous bytes
      tc::mbar_wait(accb, par);
      tc::tc_fence_after();
      if (e == 0 && lane == 0) SSTAMP(10 + 20 * pr);
      cluster_sync_relaxed();  // #1: every CTA has retired its phase-B MMAs -> the h region is free to receive
      if (e == 0 && lane == 0) SSTAMP(11 + 20 * pr);
      {
        const uint32_t mine = (rank * tpc * C::D + f) * 4;
        // this CTA's own tokens take the plain shared-memory path: a st.shared::cluster to the own rank still travels through the
        // SM-to-SM network, whose ~16 B/clk per SM is what bounds this exchange
        int dst = 0, tl = 0;
        bool self = rank == 0;
        uint32_t base = (self ? recv_local : tc::mapa(recv_local, 0)) + mine;
        for (int c = 0; c < rt / 16; ++c) {
          float v[16];
          tmem_ld16(tmem_base + tlane + 2 * aw + g * aw + c * 16, v);
          if (NSPLIT == 2) {
            float v2[16];
            tmem_ld16(tmem_base + tlane + 2 * aw + g * aw + rt + c * 16, v2);
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] += v2[k];
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (self) st_shared_f32(base + tl * (C::D * 4), v[k]);
            else tc::st_cluster_f32(base + tl * (C::D * 4), v[k]);
            if (++tl == tpc) {
              tl = 0;
              ++dst;
              self = static_cast<uint32_t>(dst & 3) == rank;
              base = (self ? recv_local : tc::mapa(recv_local, dst & 3)) + mine;
            }
          }
        }
      }
      tc::tc_fence_before();
      if (e == 0 && lane == 0) SSTAMP(12 + 20 * pr);
      tc::cluster_sync();  // #2: all partials have landed
      if (e == 0 && lane == 0) SSTAMP(13 + 20 * pr);
      // (3) owner: fixed-order sum of the 4 partials + bias (+ residual), LayerNorm over the token (16 lanes), epilogue math
      {
        const Act& o = p.out[pr];
        float y[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = 0.f;
        if (own) {
#pragma unroll
          for (int s4 = 0; s4 < C::CL; ++s4) {
            const float* rp = reinterpret_cast<const float*>(hr) + (s4 * tpc + otl) * C::D + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 a = *reinterpret_cast<const float4*>(rp + 64 * j);
              y[4 * j] += a.x; y[4 * j + 1] += a.y; y[4 * j + 2] += a.z; y[4 * j + 3] += a.w;
            }
          }
        }
        if (e == 0 && lane == 0) SSTAMP(14 + 20 * pr);
        float sm = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[4 * j] += b2v[j].x; y[4 * j + 1] += b2v[j].y; y[4 * j + 2] += b2v[j].z; y[4 * j + 3] += b2v[j].w;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (ln_res) y[k] += va[k];
          sm += y[k];
        }
#pragma unroll
        for (int ofs = 8; ofs > 0; ofs >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, ofs);
        const float mean = sm * (1.f / 256.f);
        float q2 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float dd = y[k] - mean;
          q2 += dd * dd;
        }
#pragma unroll
        for (int ofs = 8; ofs > 0; ofs >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, ofs);
        const float rstd = 1.0f / sqrtf(q2 * (1.f / 256.f) + LD_EPS);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          y[4 * j] = (y[4 * j] - mean) * rstd * gv[j].x + bv[j].x;
          y[4 * j + 1] = (y[4 * j + 1] - mean) * rstd * gv[j].y + bv[j].y;
          y[4 * j + 2] = (y[4 * j + 2] - mean) * rstd * gv[j].z + bv[j].z;
          y[4 * j + 3] = (y[4 * j + 3] - mean) * rstd * gv[j].w + bv[j].w;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (ln_res) {
            y[k] += vb[k];
          } else {  // SiLU with the fast exponential / reciprocal (<= 2 ulp each; the bf16x3 products are ~1e-5 relative)
            const float z = y[k] * va[k] + vb[k];
            y[k] = __fdividef(z, 1.0f + __expf(-z));
          }
        }
        if (e == 0 && lane == 0) SSTAMP(15 + 20 * pr);
        uint2 hi2[4], lo2[4];   // per k-block j: features c0 + 64 j .. + 3
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          split2_op<NSPLIT>(y[4 * j], y[4 * j + 1], hi2[j].x, lo2[j].x);
          split2_op<NSPLIT>(y[4 * j + 2], y[4 * j + 3], hi2[j].y, lo2[j].y);
        }
        // the exchange first: it is on the critical path of the next pair, the global stores are not
        if (!last && own) {
          // next X operand into every CTA of the cluster: k-block j, 8 bytes inside swizzle chunk (c0 >> 3)
          const uint32_t rowoff = (otcl >> 3) * 1024 + (otcl & 7) * 128 + ((((c0 >> 3) ^ (otcl & 7)) << 4) | ((c0 & 7) * 2));
#pragma unroll
          for (int k = 1; k < C::CL; ++k) {   // the three peers first (the slow path), then the own copy through plain st.shared
            const uint32_t base = tc::mapa(xop_local, (rank + k) & 3) + rowoff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              st_cluster_v2u(base + (j * NSPLIT) * XT, hi2[j].x, hi2[j].y);
              if (NSPLIT == 2) st_cluster_v2u(base + (j * NSPLIT + 1) * XT, lo2[j].x, lo2[j].y);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            st_shared_v2u(xop_local + rowoff + (j * NSPLIT) * XT, hi2[j].x, hi2[j].y);
            if (NSPLIT == 2) st_shared_v2u(xop_local + rowoff + (j * NSPLIT + 1) * XT, lo2[j].x, lo2[j].y);
          }
        }
        if (e == 0 && lane == 0) SSTAMP(16 + 20 * pr);
        if (!last) {
          fence_proxy_async_all();
          tc::cluster_sync();  // #3
        }
        if (ovalid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (o.f32) *reinterpret_cast<float4*>(o.f32 + orow * o.ld + c0 + 64 * j) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
            if (o.pl && p.out_planes > 0) {
              op16* dh = o.pl + orow * o.ld + c0 + 64 * j;
              *reinterpret_cast<uint2*>(dh) = hi2[j];
              if (p.out_planes > 1) *reinterpret_cast<uint2*>(dh + static_cast<long>(o.rows_alloc) * o.ld) = lo2[j];
            }
          }
        }
        if (e == 0 && lane == 0) SSTAMP(18 + 20 * pr);
      }
      if (e == 0 && lane == 0) SSTAMP(17 + 20 * pr);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p.trace, 3);
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
#undef SSTAMP
}
