// The fused linear  out = epilogue(A[M,K] . W[N,K]^T + bias)  in two arithmetic backends:
//   k_linear_simt : fp32 FFMA (exact fp32 products; parity path)
//   k_linear_tc   : tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) fed by TMA, optional hi/lo operand split
//                   (3 products per k-step: ~fp32-grade), warp-specialised: TMA producer / MMA issuer /
//                   4 epilogue warps that own one accumulator row per thread (LayerNorm is thread-local).
// Both share LinArgs and the same epilogue set (common.cuh: Epi).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

struct LinArgs {
  // problem: rows = min(M_max, *M_dev) when M_dev != null (ragged row counts live on the device so that a
  // captured CUDA graph is valid for every length distribution)
  int M_max;
  const int* M_dev;
  int N, K;
  // fp32 operands (SIMT backend).  Two-source A: columns [0,K1) from A, [K1,K) from A2 (skip-merge, no concat)
  const float* A;
  int lda;
  const float* A2;
  int lda2;
  int K1;
  const float* Wt;  // [K][ldw] fp32 (transposed nn.Linear weight)
  int ldw;
  // epilogue
  const float* bias;
  int epi;
  const float* res;
  int ldres;
  const float* ln_g;
  const float* ln_b;
  const float* mod;     // EPI_LN_MOD_SILU: [scale(256) | shift(256)]
  const float* addv;    // EPI_LN: optional broadcast add  out += addv[add_idx[row]*ld_add + col]
  const int* add_idx;
  int ld_add;
  const int* row_map;   // optional scatter of the fp32 output rows (dst row = row_map[row])
  Act out;
  int out_planes;       // 0 / 1 / 2 bf16 planes written
  int n_store;          // columns actually stored (<= N; 263 of 264..)
  // tensor-core backend
  int a_plane_rows, a2_plane_rows, w_plane_rows;  // row offset of the lo plane inside each tensor map
};

// ------------------------------------------------------------------------------------------------
// scalar epilogue shared by both backends (non-LN kinds)
__device__ __forceinline__ float epi_pointwise(const LinArgs& p, float acc, long row, int col) {
  float x = acc + (p.bias ? __ldg(p.bias + col) : 0.f);
  switch (p.epi) {
    case EPI_RELU: x = fmaxf(x, 0.f); break;
    case EPI_GELU: x = gelu_erf(x); break;
    case EPI_SILU: x = silu(x); break;
    case EPI_RES: x += p.res[row * p.ldres + col]; break;
    default: break;
  }
  return x;
}

// ------------------------------------------------------------------------------------------------
// fp32 SIMT backend.  CTA = 8 warps, tile = (8*RPW) rows x 256 columns; warp w owns rows w*RPW.., lane owns
// columns lane + 32 j (j < 8) so a full 256-wide row lives in one warp (LayerNorm = warp shuffles).
template <int RPW>
__global__ void __launch_bounds__(256) k_linear_simt(const LinArgs p) {
  constexpr int BM = 8 * RPW, BK = 32;
  __shared__ float As[BM][BK + 1];
  __shared__ float Ws[BK][256];
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int row0 = blockIdx.x * BM;
  if (row0 >= M) return;
  const int n0 = blockIdx.y * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[RPW][8];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      const int r = i / BK, k = i % BK;
      const long gr = row0 + r;
      const int gk = k0 + k;
      float v = 0.f;
      if (gr < M) v = (gk < p.K1) ? p.A[gr * p.lda + gk] : p.A2[gr * p.lda2 + (gk - p.K1)];
      As[r][k] = v;
    }
    for (int i = threadIdx.x; i < BK * 256; i += 256) {
      const int k = i >> 8, n = i & 255;
      const int gn = n0 + n;
      Ws[k][n] = (gn < p.N) ? __ldg(p.Wt + static_cast<long>(k0 + k) * p.ldw + gn) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < BK; ++k) {
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = Ws[k][lane + 32 * j];
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const float a = As[warp * RPW + r][k];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(a, w[j], acc[r][j]);
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const long row = row0 + warp * RPW + r;
    if (row >= M) continue;  // warp-uniform
    const long drow = p.row_map ? p.row_map[row] : row;
    if (p.epi == EPI_LN || p.epi == EPI_LN_MOD_SILU) {
      float v[8], s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = lane + 32 * j;
        v[j] = acc[r][j] + (p.bias ? __ldg(p.bias + col) : 0.f);
        if (p.epi == EPI_LN && p.res) v[j] += p.res[row * p.ldres + col];
        s += v[j];
      }
      const float mean = warp_sum(s) * (1.f / 256.f);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[j] - mean;
        q += d * d;
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.f / 256.f) + LD_EPS);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = lane + 32 * j;
        float y = (v[j] - mean) * rstd * __ldg(p.ln_g + col) + __ldg(p.ln_b + col);
        if (p.epi == EPI_LN_MOD_SILU) {
          y = silu(y * (1.f + __ldg(p.mod + col)) + __ldg(p.mod + 256 + col));
        } else if (p.addv) {
          y += p.addv[static_cast<long>(p.add_idx[row]) * p.ld_add + col];
        }
        act_store(p.out, p.out_planes, drow, col, y);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = n0 + lane + 32 * j;
        if (col < p.n_store) act_store(p.out, p.out_planes, drow, col, epi_pointwise(p, acc[r][j], row, col));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 backend
template <int BN, int NSPLIT>
struct TcCfg {
  static constexpr int BM = 128, BK = 64, UMMA_K = 16;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = NSPLIT * (A_BYTES + W_BYTES);
  static constexpr int STAGES_FIT = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 4 ? 4 : STAGES_FIT;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
};

__device__ __forceinline__ uint4 pack8_bf16(const __nv_bfloat16* h) {
  uint4 u;
  u.x = (static_cast<uint32_t>(__bfloat16_as_ushort(h[1])) << 16) | __bfloat16_as_ushort(h[0]);
  u.y = (static_cast<uint32_t>(__bfloat16_as_ushort(h[3])) << 16) | __bfloat16_as_ushort(h[2]);
  u.z = (static_cast<uint32_t>(__bfloat16_as_ushort(h[5])) << 16) | __bfloat16_as_ushort(h[4]);
  u.w = (static_cast<uint32_t>(__bfloat16_as_ushort(h[7])) << 16) | __bfloat16_as_ushort(h[6]);
  return u;
}

// store 32 consecutive finished values of one row (columns col0..col0+31)
__device__ __forceinline__ void tc_store32(const LinArgs& p, long drow, int col0, const float (&y)[32]) {
  const bool vec = (p.row_map == nullptr) && (col0 + 32 <= p.n_store) && ((p.out.ld & 7) == 0);
  if (vec) {
    if (p.out.f32) {
      float4* d = reinterpret_cast<float4*>(p.out.f32 + drow * p.out.ld + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
    }
    if (p.out.pl && p.out_planes > 0) {
      __nv_bfloat16 hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) split_bf16(y[j], hi[j], lo[j]);
      uint4* dh = reinterpret_cast<uint4*>(p.out.pl + drow * p.out.ld + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) dh[j] = pack8_bf16(hi + 8 * j);
      if (p.out_planes > 1) {
        uint4* dl = reinterpret_cast<uint4*>(p.out.pl + (static_cast<long>(p.out.rows_alloc) + drow) * p.out.ld + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) dl[j] = pack8_bf16(lo + 8 * j);
      }
    }
  } else {
    for (int j = 0; j < 32; ++j)
      if (col0 + j < p.n_store) act_store(p.out, p.out_planes, drow, col0 + j, y[j]);
  }
}

template <int BN, int NSPLIT>
__global__ void __launch_bounds__(192, 1)
k_linear_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmW, const LinArgs p) {
  using C = TcCfg<BN, NSPLIT>;
  constexpr int STAGES = C::STAGES;
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int tile_m = blockIdx.x;
  if (tile_m * C::BM >= M) return;  // CTA-uniform, before any barrier / allocation
  const int n0 = blockIdx.y * BN;
  const int nkb = p.K / C::BK;
  const int nkb1 = p.K1 / C::BK;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * C::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(accum_full, 1);
    tc::fence_barrier_init();
    tc::fence_proxy_async();
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmW);
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, BN);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        tc::mbar_expect_tx(&full[s], C::STAGE_BYTES);
        uint8_t* st = smem + s * C::STAGE_BYTES;
        const bool first = kb < nkb1;
        const CUtensorMap* ma = first ? &tmA : &tmA2;
        const int kcol = (first ? kb : kb - nkb1) * C::BK;
        const int prow = first ? p.a_plane_rows : p.a2_plane_rows;
#pragma unroll
        for (int pl = 0; pl < NSPLIT; ++pl)
          tc::tma_load_2d(st + pl * C::A_BYTES, ma, &full[s], kcol, pl * prow + tile_m * C::BM);
#pragma unroll
        for (int pl = 0; pl < NSPLIT; ++pl)
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tc::tma_load_2d(st + NSPLIT * C::A_BYTES + pl * C::W_BYTES + c * 8192, &tmW, &full[s], kb * C::BK,
                            pl * p.w_plane_rows + n0 + c * 64);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (single thread) =====
      constexpr uint32_t idesc = tc::idesc_bf16_f32(C::BM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + s * C::STAGE_BYTES);
        const uint32_t sw = sa + NSPLIT * C::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < C::BK / C::UMMA_K; ++kk) {
          const uint32_t koff = kk * C::UMMA_K * 2;  // bytes inside the 128B swizzle atom
          const uint64_t a_hi = tc::smem_desc_sw128(sa + koff);
          const uint64_t w_hi = tc::smem_desc_sw128(sw + koff);
          const uint32_t acc0 = (kb | kk) != 0;
          if (NSPLIT == 1) {
            tc::mma_bf16_ss(tmem_base, a_hi, w_hi, idesc, acc0);
          } else {
            const uint64_t a_lo = tc::smem_desc_sw128(sa + C::A_BYTES + koff);
            const uint64_t w_lo = tc::smem_desc_sw128(sw + C::W_BYTES + koff);
            tc::mma_bf16_ss(tmem_base, a_lo, w_hi, idesc, acc0);  // small terms first
            tc::mma_bf16_ss(tmem_base, a_hi, w_lo, idesc, 1u);
            tc::mma_bf16_ss(tmem_base, a_hi, w_hi, idesc, 1u);
          }
        }
        tc::mma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
      }
      tc::mma_commit(accum_full);
    }
  } else {
    // ===== epilogue: 4 warps, thread <-> accumulator row =====
    tc::mbar_wait(accum_full, 0);
    tc::tc_fence_after();
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int r = wq * 32 + lane;
    const long row = static_cast<long>(tile_m) * C::BM + r;
    const bool valid = row < M;
    const long drow = (valid && p.row_map) ? p.row_map[row] : row;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float v[32];
    if (p.epi == EPI_LN || p.epi == EPI_LN_MOD_SILU) {
      // BN == 256 == N: whole row in this thread.  3 passes over TMEM (exact two-pass variance).
      float s = 0.f;
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c * 32 + j;
          float x = v[j] + (p.bias ? __ldg(p.bias + col) : 0.f);
          if (p.epi == EPI_LN && p.res && valid) x += p.res[row * p.ldres + col];
          v[j] = x;
          s += x;
        }
        tc::tmem_st32(trow + c * 32, v);
      }
      const float mean = s * (1.f / 256.f);
      float q = 0.f;
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = v[j] - mean;
          q += d * d;
        }
      }
      const float rstd = 1.0f / sqrtf(q * (1.f / 256.f) + LD_EPS);
      const float* addrow = (p.epi == EPI_LN && p.addv && valid) ? p.addv + static_cast<long>(p.add_idx[row]) * p.ld_add : nullptr;
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c * 32 + j;
          float y = (v[j] - mean) * rstd * __ldg(p.ln_g + col) + __ldg(p.ln_b + col);
          if (p.epi == EPI_LN_MOD_SILU) y = silu(y * (1.f + __ldg(p.mod + col)) + __ldg(p.mod + 256 + col));
          else if (addrow) y += addrow[col];
          v[j] = y;
        }
        if (valid) tc_store32(p, drow, c * 32, v);
      }
    } else {
      for (int c = 0; c < BN / 32; ++c) {
        tc::tmem_ld32(trow + c * 32, v);
        const int col0 = n0 + c * 32;
        if (valid && col0 < p.n_store) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            v[j] = (col < p.N) ? epi_pointwise(p, v[j], row, col) : 0.f;
          }
          tc_store32(p, drow, col0, v);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, BN);
}
