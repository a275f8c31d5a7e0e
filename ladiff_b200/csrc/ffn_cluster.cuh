// Fused feed-forward pairs of the denoiser layer on a thread-block cluster:
//
//   pair 0 (sa_block FFN, mdiff_transformer.py:60-62 / cross_attention.py forward_post):
//       x3 = LN( x1 + W2 relu(W1 x1 + b1) + b2 ) * g + b  + delta[add_idx[row]]         (delta = hoisted ca_block output)
//   pair 1 (FFN + StylizationBlock prologue, mdiff_transformer.py:137-162,248-262):
//       s  = SiLU( LN( W2' gelu(W1' x3 + b1') + b2' ) * g' + b' ) * (1 + scale) + shift )
//
// One cluster of CL = 8 CTAs owns a 128-row tile.  CTA `rank` computes the 128-column slice h[:, 128 rank ..] of the hidden
// activation (phase A: X[128,256] . W1[slice]^T, accumulator in TMEM), converts it in place to bf16 hi/lo operand planes in
// shared memory (never to HBM) and multiplies it with the matching K-slice of W2 (phase B: partial out[128,256] in TMEM).
// The 8 partial outputs are reduce-scattered by rows through distributed shared memory (CTA r receives rows 16 r .. 16 r + 15
// from everyone, fixed summation order -> deterministic), so bias / residual / LayerNorm / modulation are local to the owner.
// Between the pairs the owners broadcast x3 as the next X operand (bf16 planes, UMMA swizzled layout) into every CTA of the
// cluster; weights never depend on activations, so the W1 slices of the next phase are prefetched while the reduction runs.
//
// Per-SM inbound TMA traffic is what bounds these links: 128 KB X + 128 KB W1 slice + 128 KB W2 slice
// per pair and CTA, instead of 512 KB of A re-read per CTA of the unfused K = 1024 LayerNorm GEMM.
#pragma once
#include <cuda.h>

#include "common.cuh"
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
  unsigned long long* trace;
  long long* dbg;        // optional per-CTA clock64 stamps [ncta][48] (ladiff_ffn_test with LADIFF_DBG_STAMPS=1)
};

template <int NSPLIT>
struct FfnCfg {
  static constexpr int CL = 8, BM = 128, BK = 64, D = 256, FF = 1024;
  static constexpr int HS = FF / CL;                // hidden columns per CTA (128)
  static constexpr int ROWS = BM / CL;              // rows owned per CTA after the reduce-scatter (16)
  static constexpr int U = BM * BK * 2;             // one [128 x 64] bf16 operand block: 16 KB
  static constexpr int X_BYTES = 4 * NSPLIT * U;    // X[kb][plane]; later W2[j][plane] (256 x 64 blocks), later the receive buffer
  static constexpr int RECV_BYTES = CL * ROWS * D * 4;  // [src][row16][256] fp32 = 128 KB
  static constexpr int XR_BYTES = X_BYTES > RECV_BYTES ? X_BYTES : RECV_BYTES;
  static constexpr int R_STAGE = NSPLIT * U;        // one W1 k-block (128 rows), all planes
  static constexpr int R_OFF = XR_BYTES;
  static constexpr int R_BYTES = 2 * R_STAGE;       // 2-stage W1 ring; later h[j][plane]
  static constexpr int VEC_OFF = R_OFF + R_BYTES;
  static constexpr int VEC_FLOATS = 2 * (HS + 3 * D + 2 * D);   // per pair: b1 slice | b2 | g | b | 1+scale | shift
  static constexpr int BAR_OFF = VEC_OFF + VEC_FLOATS * 4;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024 /*align slack*/;
  static constexpr int EPI_WARPS = 8, EPI_THREADS = 256, THREADS = 64 + EPI_THREADS;
  static constexpr int TMEM_COLS = 512;             // accA: [0,128)  accB: [128,384)
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_v2u(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int NSPLIT>
__global__ void __cluster_dims__(1, 8, 1) __launch_bounds__(FfnCfg<NSPLIT>::THREADS, 1)
k_ffn_cluster(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1a,
              const __grid_constant__ CUtensorMap tmW2a, const __grid_constant__ CUtensorMap tmW1b,
              const __grid_constant__ CUtensorMap tmW2b, const FfnArgs p) {
  using C = FfnCfg<NSPLIT>;
  constexpr int U = C::U;
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int tile_m = blockIdx.x;
  tc::pdl_launch_dependents();
  if (tile_m * C::BM >= M) return;  // cluster-uniform
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) trace_mark(p.trace, 0);
  long long* dbg = p.dbg ? p.dbg + (blockIdx.x * 8 + rank) * 48 : nullptr;
#define FSTAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) FSTAMP(0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xr = smem;                 // X operand / W2 blocks / receive buffer
  uint8_t* rr = smem + C::R_OFF;      // W1 ring / h operand
  float* vec = reinterpret_cast<float*>(smem + C::VEC_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* x_full = bars;            // [4]
  uint64_t* w1_full = bars + 4;       // [2]
  uint64_t* kb_done = bars + 6;       // [4]  phase-A MMAs of k-block kb retired (kb_done[3] == accumulator A ready)
  uint64_t* h_full = bars + 10;       // [2]  256 arrivals
  uint64_t* w2_full = bars + 12;      // [2]
  uint64_t* accb_full = bars + 14;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // provably warp-uniform
  constexpr int PV = C::HS + 5 * C::D;  // floats per pair in `vec`

  auto load_w1 = [&](int pr, int kb) {
    const int s = kb & 1;
    const CUtensorMap* m = pr == 0 ? &tmW1a : &tmW1b;
    tc::mbar_expect_tx(&w1_full[s], C::R_STAGE);
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl)
      tc::tma_load_2d(rr + s * C::R_STAGE + pl * U, m, &w1_full[s], kb * C::BK, pl * p.w1_plane_rows[pr] + static_cast<int>(rank) * C::HS);
  };
  // W2 block (j, pl): all 256 output rows x 64 hidden columns [rank*128 + 64 j, +64) -> X units [(j*NSPLIT+pl)*2, +2)
  auto load_w2 = [&](int pr, int j, int pl) {
    const CUtensorMap* m = pr == 0 ? &tmW2a : &tmW2b;
    if (pl == 0) tc::mbar_expect_tx(&w2_full[j], NSPLIT * 2 * U);
    tc::tma_load_2d(xr + (j * NSPLIT + pl) * 2 * U, m, &w2_full[j], static_cast<int>(rank) * C::HS + j * C::BK, pl * p.w2_plane_rows[pr]);
  };

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(&x_full[i], 1);
      tc::mbar_init(&kb_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&w1_full[i], 1);
      tc::mbar_init(&h_full[i], C::EPI_THREADS);
      tc::mbar_init(&w2_full[i], 1);
    }
    tc::mbar_init(accb_full, 1);
    tc::fence_barrier_init();
    tc::fence_proxy_async();
  }
  if (warp == 0) {
    __syncwarp();
    // weights never depend on the previous grid: the first two W1 k-blocks are requested before the dependency wait
    if (tc::elect_one()) {
      load_w1(0, 0);
      load_w1(0, 1);
    }
    __syncwarp();
    tc::pdl_wait();
    if (lane == 0) {
      trace_mark(p.trace, 1);
      FSTAMP(1);
    }
    if (tc::elect_one()) {
      for (int kb = 0; kb < 4; ++kb) {
        tc::mbar_expect_tx(&x_full[kb], NSPLIT * U);
#pragma unroll
        for (int pl = 0; pl < NSPLIT; ++pl)
          tc::tma_load_2d(xr + (kb * NSPLIT + pl) * U, &tmX, &x_full[kb], kb * C::BK, pl * p.x_plane_rows + tile_m * C::BM);
      }
    }
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
  const uint32_t tmem_a = tmem_base, tmem_b = tmem_base + C::HS;
  // every CTA of the cluster must have initialised its barriers / be resident before any DSMEM traffic: the first
  // cluster barrier below (#1 of pair 0) is reached only after phase B, long after this point for every CTA.

  if (warp == 0) {
    // ===== TMA producer =====
    for (int pr = 0; pr < p.npairs; ++pr) {
      {
        const uint32_t par = pr & 1;
        for (int kb = 0; kb < 4; ++kb) {
          tc::mbar_wait(&kb_done[kb], par);
          if (tc::elect_one()) {
            if (kb < 2) load_w1(pr, kb + 2);
            // W2 blocks whose destination (X k-blocks) is dead once k-block kb has been consumed
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
              for (int pl = 0; pl < NSPLIT; ++pl)
                if (((j * NSPLIT + pl) * 2 + 1) / NSPLIT == kb) load_w2(pr, j, pl);
          }
          __syncwarp();
        }
        if (pr + 1 < p.npairs) {
          tc::mbar_wait(accb_full, par);  // ring (W1 k-blocks 2,3 / h) is dead: prefetch the next pair's first W1 k-blocks
          if (tc::elect_one()) {
            load_w1(pr + 1, 0);
            load_w1(pr + 1, 1);
          }
        }
      }
      __syncwarp();
      tc::cluster_sync();  // #1
      tc::cluster_sync();  // #2
      if (pr + 1 < p.npairs) {
        tc::cluster_sync();  // #3
        tc::cluster_sync();  // #4
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc_a = tc::idesc_bf16_f32(C::BM, C::HS);
    constexpr uint32_t idesc_b = tc::idesc_bf16_f32(C::BM, C::D);
    const uint32_t xr_u = tc::smem_u32(xr), rr_u = tc::smem_u32(rr);
    for (int pr = 0; pr < p.npairs; ++pr) {
      {
        const uint32_t par = pr & 1;
        // phase A: accA[128 x 128] = X . W1[slice]^T
        for (int kb = 0; kb < 4; ++kb) {
          if (pr == 0) tc::mbar_wait(&x_full[kb], 0);
          tc::mbar_wait(&w1_full[kb & 1], (kb >> 1) & 1);
          tc::tc_fence_after();
          if (lane == 0) {
            if (kb == 0) FSTAMP(2 + 20 * pr);
            if (kb == 3) FSTAMP(3 + 20 * pr);
          }
          const uint32_t sa = xr_u + kb * NSPLIT * U;
          const uint32_t sw = rr_u + (kb & 1) * C::R_STAGE;
          if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t koff = kk * 32;
            const uint64_t a_hi = tc::smem_desc_sw128(sa + koff), w_hi = tc::smem_desc_sw128(sw + koff);
            const uint32_t acc0 = (kb | kk) != 0;
            if (NSPLIT == 1) {
              tc::mma_bf16_ss(tmem_a, a_hi, w_hi, idesc_a, acc0);
            } else {
              const uint64_t a_lo = tc::smem_desc_sw128(sa + U + koff), w_lo = tc::smem_desc_sw128(sw + U + koff);
              tc::mma_bf16_ss(tmem_a, a_lo, w_hi, idesc_a, acc0);
              tc::mma_bf16_ss(tmem_a, a_hi, w_lo, idesc_a, 1u);
              tc::mma_bf16_ss(tmem_a, a_hi, w_hi, idesc_a, 1u);
            }
          }
          tc::mma_commit(&kb_done[kb]);
          }
          __syncwarp();
        }
        // phase B: accB[128 x 256] = h[:, slice] . W2[:, slice]^T
        for (int j = 0; j < 2; ++j) {
          tc::mbar_wait(&w2_full[j], par);
          if (j == 0 && lane == 0) FSTAMP(4 + 20 * pr);
          tc::mbar_wait(&h_full[j], par);
          tc::tc_fence_after();
          if (lane == 0) FSTAMP(5 + j + 20 * pr);
          const uint32_t sa = rr_u + j * NSPLIT * U;
          const uint32_t sw = xr_u + j * NSPLIT * 2 * U;
          if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t koff = kk * 32;
            const uint64_t a_hi = tc::smem_desc_sw128(sa + koff), w_hi = tc::smem_desc_sw128(sw + koff);
            const uint32_t acc0 = (j | kk) != 0;
            if (NSPLIT == 1) {
              tc::mma_bf16_ss(tmem_b, a_hi, w_hi, idesc_b, acc0);
            } else {
              const uint64_t a_lo = tc::smem_desc_sw128(sa + U + koff), w_lo = tc::smem_desc_sw128(sw + 2 * U + koff);
              tc::mma_bf16_ss(tmem_b, a_lo, w_hi, idesc_b, acc0);
              tc::mma_bf16_ss(tmem_b, a_hi, w_lo, idesc_b, 1u);
              tc::mma_bf16_ss(tmem_b, a_hi, w_hi, idesc_b, 1u);
            }
          }
          if (j == 1) tc::mma_commit(accb_full);
          }
          __syncwarp();
        }
      }
      __syncwarp();
      tc::cluster_sync();  // #1
      tc::cluster_sync();  // #2
      if (pr + 1 < p.npairs) {
        tc::cluster_sync();  // #3
        tc::cluster_sync();  // #4: the next X operand has been written by the owners (generic proxy, remote CTAs)
        fence_proxy_async_all();
      }
    }
  } else {
    // ===== epilogue warps =====
    const int et = threadIdx.x - 64;       // 0..255
    const int wq = warp & 3;               // TMEM lane quarter of this warp
    const int ch = (warp - 2) >> 2;        // column half
    const int r = wq * 32 + lane;          // accumulator row of this thread
    const uint32_t trow = static_cast<uint32_t>(wq * 32) << 16;
    // reduce-scatter ownership: thread (row16, cg) owns columns (g*16 + cg)*4 .. +3, g = 0..3 of row 16*rank + row16
    const int row16 = et >> 4, cg = et & 15;
    const long orow = static_cast<long>(tile_m) * C::BM + rank * C::ROWS + row16;
    const bool ovalid = orow < M;
    // per-column vectors of both pairs (weights: before the dependency wait)
    for (int pr = 0; pr < p.npairs; ++pr) {
      float* v = vec + pr * PV;
      for (int i = et; i < C::HS; i += C::EPI_THREADS) v[i] = __ldg(p.b1[pr] + rank * C::HS + i);
      for (int i = et; i < C::D; i += C::EPI_THREADS) {
        v[C::HS + i] = __ldg(p.b2[pr] + i);
        v[C::HS + C::D + i] = __ldg(p.ln_g[pr] + i);
        v[C::HS + 2 * C::D + i] = __ldg(p.ln_b[pr] + i);
      }
    }
    tc::pdl_wait();
    for (int pr = 0; pr < p.npairs; ++pr) {
      if (p.kind[pr] == EPI_LN_MOD_SILU) {
        float* v = vec + pr * PV;
        for (int i = et; i < C::D; i += C::EPI_THREADS) {
          v[C::HS + 3 * C::D + i] = 1.f + p.mod[pr][i];
          v[C::HS + 4 * C::D + i] = p.mod[pr][C::D + i];
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");

    const uint32_t recv_local = tc::smem_u32(xr);
    const uint32_t dst_rank = static_cast<uint32_t>(r >> 4);
    const uint32_t recv_dst = tc::mapa(recv_local, dst_rank) + ((rank * C::ROWS + (r & 15)) * C::D) * 4;

    for (int pr = 0; pr < p.npairs; ++pr) {
      const uint32_t par = pr & 1;
      const float* pv = vec + pr * PV;
      const bool ln_res = p.kind[pr] == EPI_LN;
      // (0) owner-side operands straight into registers while the mainloop runs
      float4 rs[4], ad[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        rs[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        ad[g] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (ln_res && ovalid) {
        if (pr == 0 && p.res) {
#pragma unroll
          for (int g = 0; g < 4; ++g) rs[g] = *reinterpret_cast<const float4*>(p.res + orow * C::D + (g * 16 + cg) * 4);
        }
        if (p.addv) {
          const long sr = static_cast<long>(__ldg(p.add_idx + orow));
#pragma unroll
          for (int g = 0; g < 4; ++g) ad[g] = *reinterpret_cast<const float4*>(p.addv + sr * p.ld_add + (g * 16 + cg) * 4);
        }
      }
      // (1) hidden activation: accA -> +b1 -> act -> bf16 hi/lo planes in the UMMA K-major 128B-swizzled layout
      tc::mbar_wait(&kb_done[3], par);
      tc::tc_fence_after();
      if (pr == 0 && et == 0) trace_mark(p.trace, 2);
      if (et == 0) FSTAMP(8 + 20 * pr);
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        const int c = ch + 2 * i;  // 32-column chunk of the 128-column slice
        float v[32];
        tc::tmem_ld32(tmem_a + trow + c * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(pv + c * 32 + 4 * q);
          v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
        }
        if (p.act[pr] == EPI_RELU) {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = gelu_erf_fast(v[k]);
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) split2_bf16(v[2 * k], v[2 * k + 1], hi[k], lo[k]);
        const int j = c >> 1;
        uint8_t* hb = rr + j * NSPLIT * U + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c16 = (((c & 1) * 4 + q) ^ (r & 7)) * 16;
          *reinterpret_cast<uint4*>(hb + c16) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
          if (NSPLIT == 2) *reinterpret_cast<uint4*>(hb + U + c16) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&h_full[j]);
      }
      // (2) partial outputs -> owners (rows 16 k .. of the tile belong to CTA k)
      if (et == 0) FSTAMP(9 + 20 * pr);
      tc::mbar_wait(accb_full, par);
      tc::tc_fence_after();
      if (et == 0) FSTAMP(10 + 20 * pr);
      tc::cluster_sync();  // #1: every CTA has retired its phase-B MMAs -> X region is free to receive
      if (et == 0) FSTAMP(11 + 20 * pr);
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        const int c = ch + 2 * i;
        float v[32];
        tc::tmem_ld32(tmem_b + trow + c * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) st_cluster_v4(recv_dst + (c * 32 + q * 4) * 4, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      tc::tc_fence_before();
      if (et == 0) FSTAMP(12 + 20 * pr);
      tc::cluster_sync();  // #2: all partials have landed
      if (et == 0) FSTAMP(13 + 20 * pr);
      // (3) owner: fixed-order sum of the 8 partials + bias (+ residual), LayerNorm over the row (16 lanes), epilogue math
      float y[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 b4 = *reinterpret_cast<const float4*>(pv + C::HS + (g * 16 + cg) * 4);
        float4 a = *reinterpret_cast<const float4*>(xr + ((0 * C::ROWS + row16) * C::D + (g * 16 + cg) * 4) * 4);
#pragma unroll
        for (int s = 1; s < C::CL; ++s) {
          const float4 t = *reinterpret_cast<const float4*>(xr + ((s * C::ROWS + row16) * C::D + (g * 16 + cg) * 4) * 4);
          a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        y[g][0] = a.x + b4.x + rs[g].x; y[g][1] = a.y + b4.y + rs[g].y; y[g][2] = a.z + b4.z + rs[g].z; y[g][3] = a.w + b4.w + rs[g].w;
      }
      if (et == 0) FSTAMP(14 + 20 * pr);
      if (pr + 1 < p.npairs) tc::cluster_sync();  // #3: receive buffers consumed everywhere -> X region may take the next operand
      if (et == 0) FSTAMP(15 + 20 * pr);
      float sm = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) sm += (y[g][0] + y[g][1]) + (y[g][2] + y[g][3]);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
      const float mean = sm * (1.f / 256.f);
      float q2 = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = y[g][e] - mean;
          q2 += d * d;
        }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
      const float rstd = 1.0f / sqrtf(q2 * (1.f / 256.f) + LD_EPS);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c0 = (g * 16 + cg) * 4;
        const float4 g4 = *reinterpret_cast<const float4*>(pv + C::HS + C::D + c0);
        const float4 b4 = *reinterpret_cast<const float4*>(pv + C::HS + 2 * C::D + c0);
        y[g][0] = (y[g][0] - mean) * rstd * g4.x + b4.x;
        y[g][1] = (y[g][1] - mean) * rstd * g4.y + b4.y;
        y[g][2] = (y[g][2] - mean) * rstd * g4.z + b4.z;
        y[g][3] = (y[g][3] - mean) * rstd * g4.w + b4.w;
        if (!ln_res) {
          const float4 s4 = *reinterpret_cast<const float4*>(pv + C::HS + 3 * C::D + c0);
          const float4 h4 = *reinterpret_cast<const float4*>(pv + C::HS + 4 * C::D + c0);
          y[g][0] = silu(y[g][0] * s4.x + h4.x);
          y[g][1] = silu(y[g][1] * s4.y + h4.y);
          y[g][2] = silu(y[g][2] * s4.z + h4.z);
          y[g][3] = silu(y[g][3] * s4.w + h4.w);
        } else {
          y[g][0] += ad[g].x; y[g][1] += ad[g].y; y[g][2] += ad[g].z; y[g][3] += ad[g].w;
        }
      }
      // (4) stores: fp32 master and bf16 planes of this pair's output; next X operand into every CTA of the cluster
      const Act& o = p.out[pr];
      const int trow_l = static_cast<int>(rank) * C::ROWS + row16;  // row inside the tile
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int c0 = (g * 16 + cg) * 4;
        uint32_t h01, l01, h23, l23;
        split2_bf16(y[g][0], y[g][1], h01, l01);
        split2_bf16(y[g][2], y[g][3], h23, l23);
        if (ovalid) {
          if (o.f32) *reinterpret_cast<float4*>(o.f32 + orow * o.ld + c0) = make_float4(y[g][0], y[g][1], y[g][2], y[g][3]);
          if (o.pl && p.out_planes > 0) {
            __nv_bfloat16* dh = o.pl + orow * o.ld + c0;
            *reinterpret_cast<uint2*>(dh) = make_uint2(h01, h23);
            if (p.out_planes > 1) *reinterpret_cast<uint2*>(dh + static_cast<long>(o.rows_alloc) * o.ld) = make_uint2(l01, l23);
          }
        }
        if (pr + 1 < p.npairs) {
          const int kb = c0 >> 6, within = c0 & 63;
          const uint32_t off = (kb * NSPLIT) * U + (trow_l >> 3) * 1024 + (trow_l & 7) * 128 + (((within >> 3) ^ (trow_l & 7)) * 16) +
                               (within & 7) * 2;
#pragma unroll
          for (int k = 0; k < C::CL; ++k) {
            const uint32_t base = tc::mapa(recv_local, k) + off;
            st_cluster_v2u(base, h01, h23);
            if (NSPLIT == 2) st_cluster_v2u(base + U, l01, l23);
          }
        }
      }
      if (et == 0) FSTAMP(16 + 20 * pr);
      if (pr + 1 < p.npairs) {
        fence_proxy_async_all();
        tc::cluster_sync();  // #4
      }
      if (et == 0) FSTAMP(17 + 20 * pr);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p.trace, 3);
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
#undef FSTAMP
}
