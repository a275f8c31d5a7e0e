// Fused feed-forward pairs of the denoiser layer, "swapped" formulation for the latency-bound batch sizes (R <= 1776 rows):
// the WEIGHT rows sit on the 128 lanes of the tcgen05 M axis and a small group of RT <= 48 tokens on the N axis
//       h^T  [256-slice x RT]  = W1[slice, :]   . X^T          (phase A, two 128-row M tiles)
//       out^T[256       x RT]  = W2[:, slice]   . h[:, slice]^T (phase B, K-split partial)
// so a cluster of only CL = 4 CTAs owns RT tokens (not 128 rows as in ffn_cluster.cuh) and 1280 rows spread over 27 clusters /
// 108 SMs.  What bounds the fused FFN on B200 is the distributed-shared-memory exchange (~16 B/clk per SM measured,
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
#include "ffn_cluster.cuh"
#include "linear.cuh"
#include "tc_ptx.cuh"

template <int NSPLIT>
struct SwapCfg {
  static constexpr int CL = 4, D = 256, FF = 1024, HS = FF / CL, BK = 64;
  static constexpr int U = 128 * BK * 2;       // one [128 x 64] bf16 weight block (16 KB)
  static constexpr int STG = NSPLIT * U;       // ring stage: all planes of one weight block
  static constexpr int RT_MAX = 48;
  static constexpr int EPI_WARPS = 8, THREADS = 64 + EPI_WARPS * 32;
  static constexpr int TMEM_COLS = 256;        // acc A: [0, 2 RT)   acc B: [2 RT, 4 RT)
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
__device__ __forceinline__ void st_cluster_v4u(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int NSPLIT>
__global__ void __cluster_dims__(1, 4, 1) __launch_bounds__(SwapCfg<NSPLIT>::THREADS, 1)
k_ffn_swap(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1a,
           const __grid_constant__ CUtensorMap tmW2a, const __grid_constant__ CUtensorMap tmW1b,
           const __grid_constant__ CUtensorMap tmW2b, const FfnArgs p) {
  using C = SwapCfg<NSPLIT>;
  constexpr int U = C::U, STG = C::STG;
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int rt = p.rt, tpc = rt >> 2;       // tokens per cluster / owned per CTA
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
  uint64_t* hfull = bars + 14;              // [2]  128 arrivals: h operand k-blocks 2g, 2g+1 written
  uint64_t* accb = bars + 16;               // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  // warp index through a shuffle: provably warp-uniform, so the MMA / TMA operands of the role branches stay in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int total = p.npairs * 16;          // weight stages: per pair 8 of W1 (g, kb) then 8 of W2 (g, kb)

  auto issue_load = [&](int j) {
    const int slot = j & 3, pr = j >> 4, r = j & 15, ph = r >> 3, g = (r >> 2) & 1, kb = r & 3;
    const CUtensorMap* m;
    int kcol, row, prow;
    if (ph == 0) {
      m = pr == 0 ? &tmW1a : &tmW1b;
      kcol = kb * C::BK;
      row = static_cast<int>(rank) * C::HS + g * 128;
      prow = p.w1_plane_rows[pr];
    } else {
      m = pr == 0 ? &tmW2a : &tmW2b;
      kcol = static_cast<int>(rank) * C::HS + kb * C::BK;
      row = g * 128;
      prow = p.w2_plane_rows[pr];
    }
    tc::mbar_expect_tx(&full[slot], STG);
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl) tc::tma_load_2d(ring + slot * STG + pl * U, m, &full[slot], kcol, pl * prow + row);
  };

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) {
        tc::mbar_init(&full[i], 1);
        tc::mbar_init(&empty[i], 1);
        tc::mbar_init(&x_full[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&acca[i], 1);
        tc::mbar_init(&hfull[i], 128);
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
    if (tc::elect_one()) {
      for (int kb = 0; kb < 4; ++kb) {
        tc::mbar_expect_tx(&x_full[kb], NSPLIT * XT);
        for (int pl = 0; pl < NSPLIT; ++pl)
          for (int i = 0; i < rt / 16; ++i)
            tc::tma_load_2d(xop + (kb * NSPLIT + pl) * XT + i * 2048, &tmX, &x_full[kb], kb * C::BK, pl * p.x_plane_rows + row0 + i * 16);
      }
    }
    __syncwarp();
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
      tc::cluster_sync();  // #1
      tc::cluster_sync();  // #2
      if (pr + 1 < p.npairs) tc::cluster_sync();  // #3
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =====
    const uint32_t idesc = tc::idesc_bf16_f32(128, rt);
    const uint32_t ring_u = tc::smem_u32(ring), xop_u = tc::smem_u32(xop), hr_u = tc::smem_u32(hr);
    for (int pr = 0; pr < p.npairs; ++pr) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int slot = r & 3, ph = r >> 3, g = (r >> 2) & 1, kb = r & 3;   // 16 stages per pair: slot = (16 pr + r) & 3
        if (ph == 0) {
          if (pr == 0) tc::mbar_wait(&x_full[kb], 0);
        } else if (g == 0) {
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
        const uint32_t d = tmem_base + (ph ? 2 * rt : 0) + g * rt;
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
              const uint64_t w_lo = tc::smem_desc_sw128(sw + U + koff), x_lo = tc::smem_desc_sw128(sx + XT + koff);
              tc::mma_bf16_ss(d, w_hi, x_lo, idesc, acc0);
              tc::mma_bf16_ss(d, w_lo, x_hi, idesc, 1u);
              tc::mma_bf16_ss(d, w_hi, x_hi, idesc, 1u);
            }
          }
          tc::mma_commit(&empty[slot]);
          if (ph == 0 && kb == 3) tc::mma_commit(&acca[g]);
          if (r == 15) tc::mma_commit(accb);
        }
        __syncwarp();
      }
      if (lane == 0) SSTAMP(4 + 20 * pr);
      tc::cluster_sync();  // #1
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
    const int c0 = lane * 8;               // owner phase: 8 features per lane
    const float b1v0 = __ldg(p.b1[0] + rank * C::HS + f);
    const float b1v1 = p.npairs > 1 ? __ldg(p.b1[1] + rank * C::HS + f) : 0.f;
    tc::pdl_wait();
    // owner-side operands of pair 0 (residual, hoisted ca_block delta) straight into registers while the mainloop runs
    float rs[2][8], ad[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) rs[u][k] = ad[u][k] = 0.f;
      const int tl = e + 8 * u;
      const long row = static_cast<long>(row0) + rank * tpc + tl;
      if (tl < tpc && row < M && p.kind[0] == EPI_LN) {
        if (p.res) {
          const float4 a = *reinterpret_cast<const float4*>(p.res + row * C::D + c0), b = *reinterpret_cast<const float4*>(p.res + row * C::D + c0 + 4);
          rs[u][0] = a.x; rs[u][1] = a.y; rs[u][2] = a.z; rs[u][3] = a.w; rs[u][4] = b.x; rs[u][5] = b.y; rs[u][6] = b.z; rs[u][7] = b.w;
        }
        if (p.addv) {
          const float* ap = p.addv + static_cast<long>(__ldg(p.add_idx + row)) * p.ld_add + c0;
          const float4 a = *reinterpret_cast<const float4*>(ap), b = *reinterpret_cast<const float4*>(ap + 4);
          ad[u][0] = a.x; ad[u][1] = a.y; ad[u][2] = a.z; ad[u][3] = a.w; ad[u][4] = b.x; ad[u][5] = b.y; ad[u][6] = b.z; ad[u][7] = b.w;
        }
      }
    }

    for (int pr = 0; pr < p.npairs; ++pr) {
      const uint32_t par = pr & 1;
      const bool ln_res = p.kind[pr] == EPI_LN;
      const bool last = pr + 1 >= p.npairs;
      // (1) hidden activation: accA^T -> + b1 -> act -> bf16 hi/lo planes of the phase-B operand (token-major, K-major swizzled)
      tc::mbar_wait(&acca[g], par);
      tc::tc_fence_after();
      if (pr == 0 && e == 0 && lane == 0) trace_mark(p.trace, 2);
      if (e == 0 && lane == 0) SSTAMP(8 + 20 * pr);
      {
        const int kbh = f >> 6, within = f & 63;
        uint8_t* hb = hr + (kbh * NSPLIT) * XT + (within & 7) * 2;
        const int chunk = within >> 3;
        const bool relu = p.act[pr] == EPI_RELU;
        const float b1 = pr ? b1v1 : b1v0;
        for (int c = 0; c < rt / 16; ++c) {
          float v[16];
          tmem_ld16(tmem_base + tlane + g * rt + c * 16, v);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int t = c * 16 + k;
            float x = v[k] + b1;
            x = relu ? fmaxf(x, 0.f) : gelu_erf_fast(x);
            __nv_bfloat16 hi, lo;
            split_bf16(x, hi, lo);
            const int off = (t >> 3) * 1024 + (t & 7) * 128 + ((chunk ^ (t & 7)) << 4);
            *reinterpret_cast<__nv_bfloat16*>(hb + off) = hi;
            if (NSPLIT == 2) *reinterpret_cast<__nv_bfloat16*>(hb + XT + off) = lo;
          }
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&hfull[g]);
      }
      if (e == 0 && lane == 0) SSTAMP(9 + 20 * pr);
      // per-feature vectors of the owner phase (weights: L2 hits), loaded while phase B runs
      float b2v[8], gv[8], bv[8], m1[8], m2[8];
      {
        const float4 *pb2 = reinterpret_cast<const float4*>(p.b2[pr] + c0), *pg = reinterpret_cast<const float4*>(p.ln_g[pr] + c0),
                     *pb = reinterpret_cast<const float4*>(p.ln_b[pr] + c0);
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          const float4 a = __ldg(pb2 + hlf), b = __ldg(pg + hlf), c = __ldg(pb + hlf);
          b2v[4 * hlf] = a.x; b2v[4 * hlf + 1] = a.y; b2v[4 * hlf + 2] = a.z; b2v[4 * hlf + 3] = a.w;
          gv[4 * hlf] = b.x; gv[4 * hlf + 1] = b.y; gv[4 * hlf + 2] = b.z; gv[4 * hlf + 3] = b.w;
          bv[4 * hlf] = c.x; bv[4 * hlf + 1] = c.y; bv[4 * hlf + 2] = c.z; bv[4 * hlf + 3] = c.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) m1[k] = m2[k] = 0.f;
        if (!ln_res) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            m1[k] = 1.f + p.mod[pr][c0 + k];
            m2[k] = p.mod[pr][C::D + c0 + k];
          }
        }
      }
      // (2) partial outputs -> owners (tokens tpc*k .. of the group belong to CTA k); 32 lanes = 128 contiguous bytes
      tc::mbar_wait(accb, par);
      tc::tc_fence_after();
      if (e == 0 && lane == 0) SSTAMP(10 + 20 * pr);
      tc::cluster_sync();  // #1: every CTA has retired its phase-B MMAs -> the h region is free to receive
      if (e == 0 && lane == 0) SSTAMP(11 + 20 * pr);
      {
        const uint32_t mine = (rank * tpc * C::D + f) * 4;
        int dst = 0, tl = 0;
        uint32_t base = tc::mapa(recv_local, 0) + mine;
        for (int c = 0; c < rt / 16; ++c) {
          float v[16];
          tmem_ld16(tmem_base + tlane + 2 * rt + g * rt + c * 16, v);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            tc::st_cluster_f32(base + tl * (C::D * 4), v[k]);
            if (++tl == tpc) {
              tl = 0;
              ++dst;
              base = tc::mapa(recv_local, dst & 3) + mine;
            }
          }
        }
      }
      tc::tc_fence_before();
      if (e == 0 && lane == 0) SSTAMP(12 + 20 * pr);
      tc::cluster_sync();  // #2: all partials have landed
      if (e == 0 && lane == 0) SSTAMP(13 + 20 * pr);
      // (3) owner: fixed-order sum of the 4 partials + bias (+ residual), LayerNorm over the token (one warp), epilogue math
      const Act& o = p.out[pr];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int tl = e + 8 * u;
        if (tl < tpc) {
        const int tcl = static_cast<int>(rank) * tpc + tl;  // token inside the cluster's group
        const long row = static_cast<long>(row0) + tcl;
        const bool valid = row < M;
        float y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = 0.f;
#pragma unroll
        for (int s = 0; s < C::CL; ++s) {
          const float* rp = reinterpret_cast<const float*>(hr) + (s * tpc + tl) * C::D + c0;
          const float4 a = *reinterpret_cast<const float4*>(rp), b = *reinterpret_cast<const float4*>(rp + 4);
          y[0] += a.x; y[1] += a.y; y[2] += a.z; y[3] += a.w; y[4] += b.x; y[5] += b.y; y[6] += b.z; y[7] += b.w;
        }
        float sm = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          y[k] += b2v[k] + (pr == 0 ? rs[u][k] : 0.f);
          sm += y[k];
        }
        const float mean = warp_sum(sm) * (1.f / 256.f);
        float q2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float dd = y[k] - mean;
          q2 += dd * dd;
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q2) * (1.f / 256.f) + LD_EPS);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          y[k] = (y[k] - mean) * rstd * gv[k] + bv[k];
          if (ln_res) y[k] += (pr == 0 ? ad[u][k] : 0.f);
          else y[k] = silu(y[k] * m1[k] + m2[k]);
        }
        uint4 hi4, lo4;
        split2_bf16(y[0], y[1], hi4.x, lo4.x);
        split2_bf16(y[2], y[3], hi4.y, lo4.y);
        split2_bf16(y[4], y[5], hi4.z, lo4.z);
        split2_bf16(y[6], y[7], hi4.w, lo4.w);
        if (valid) {
          if (o.f32) {
            *reinterpret_cast<float4*>(o.f32 + row * o.ld + c0) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(o.f32 + row * o.ld + c0 + 4) = make_float4(y[4], y[5], y[6], y[7]);
          }
          if (o.pl && p.out_planes > 0) {
            __nv_bfloat16* dh = o.pl + row * o.ld + c0;
            *reinterpret_cast<uint4*>(dh) = hi4;
            if (p.out_planes > 1) *reinterpret_cast<uint4*>(dh + static_cast<long>(o.rows_alloc) * o.ld) = lo4;
          }
        }
        if (!last) {
          // next X operand into every CTA of the cluster: features c0..c0+7 = one 16-byte swizzle chunk of k-block c0 / 64
          const int kb = c0 >> 6, chunk = (c0 & 63) >> 3;
          const uint32_t off = (kb * NSPLIT) * XT + (tcl >> 3) * 1024 + (tcl & 7) * 128 + ((chunk ^ (tcl & 7)) << 4);
#pragma unroll
          for (int k = 0; k < C::CL; ++k) {
            const uint32_t base = tc::mapa(xop_local, k) + off;
            st_cluster_v4u(base, hi4);
            if (NSPLIT == 2) st_cluster_v4u(base + XT, lo4);
          }
        }
        }
      }
      if (e == 0 && lane == 0) SSTAMP(16 + 20 * pr);
      if (!last) {
        fence_proxy_async_all();
        tc::cluster_sync();  // #3
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
