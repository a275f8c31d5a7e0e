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
  // fp32 operands (SIMT backend).  Multi-source A: columns [0,K1) from A, [K1,K2) from A2, [K2,K) from A3 (skip merge /
  // folded residual branches: no concat is ever materialised)
  const float* A;
  int lda;
  const float* A2;
  int lda2;
  const float* A3;
  int lda3;
  int K1, K2;
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
  int a_plane_rows, a2_plane_rows, a3_plane_rows, w_plane_rows;  // row offset of the lo plane inside each tensor map
  int dbg_flags;        // experiments (profiling only)
  long long* dbg;       // optional per-CTA clock stamps [ncta][16] (profiling builds of ladiff_linear_bench)
  unsigned long long* trace;  // optional per-launch %globaltimer record (LADIFF_TRACE=1): [0] min CTA start, [1..3] ~max of
                              // {dependency wait returned, accumulator ready, CTA done}
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// slot 0 keeps the minimum, slots >= 1 the maximum (stored complemented so that one memset(0xFF) initialises both)
__device__ __forceinline__ void trace_mark(unsigned long long* tr, int slot) {
  if (!tr) return;
  const unsigned long long t = gtimer();
  atomicMin(tr + slot, slot == 0 ? t : ~t);
}

// ------------------------------------------------------------------------------------------------
// scalar epilogue shared by both backends (non-LN kinds)
template <int EPI>
__device__ __forceinline__ float epi_pointwise_t(const LinArgs& p, float acc, long row, int col) {
  float x = acc + (p.bias ? __ldg(p.bias + col) : 0.f);
  if (EPI == EPI_RELU) x = fmaxf(x, 0.f);
  else if (EPI == EPI_GELU) x = gelu_erf(x);
  else if (EPI == EPI_SILU) x = silu(x);
  else if (EPI == EPI_RES) x += p.res[row * p.ldres + col];
  return x;
}

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
  pdl_prologue();
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
      if (gr < M)
        v = (gk < p.K1) ? p.A[gr * p.lda + gk] : (gk < p.K2 ? p.A2[gr * p.lda2 + (gk - p.K1)] : p.A3[gr * p.lda3 + (gk - p.K2)]);
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

// A packed weight carries THREE 16-bit planes of n_pad rows: fp16 hi | fp16 lo | bf16 (kernels.cuh k_pack_weight).
// The x3 mode reads planes 0 and 1, the bf16 mode plane 2.
template <int NSPLIT>
__device__ __forceinline__ constexpr int w_plane(int pl) { return NSPLIT == 2 ? pl : 2; }

// ------------------------------------------------------------------------------------------------
// tcgen05 backend
//
// Shared memory:  [ STAGES x { A planes | W planes } ]  [ per-column vectors 5 KB ]  [ mbarriers ]
// After the last MMA retires the stage memory is dead, so the epilogue aliases its transposition tiles onto it.
template <int BN, int NSPLIT, int EW = 8>
struct TcCfg {
  static constexpr int BM = 128, BK = 64, UMMA_K = 16;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = NSPLIT * (A_BYTES + W_BYTES);
  static constexpr int STAGES_FIT = (196 * 1024) / STAGE_BYTES;
  // BN == 64 is the tile of the latency-bound denoiser links.  In plain bf16 mode its 4-stage ring is 96 KB, so two CTAs
  // fit on an SM; the hi/lo split mode keeps the whole K = 256 operand set in flight (4 x 48 KB), one CTA per SM
  // (a 2-stage ring measured slower: den_qkv 6.4 -> 7.9 us).
  static constexpr int MIN_CTAS = (BN == 64 && NSPLIT == 1) ? 2 : 1;
  static constexpr int STAGES = STAGES_FIT > 4 ? 4 : STAGES_FIT;
  static constexpr int VEC_OFF = STAGES * STAGE_BYTES;
  static constexpr int VEC_BYTES = 1280 * 4;  // bias[256] | ln_g[256] | ln_b[256] | scale[256] | shift[256]
  static constexpr int BAR_OFF = VEC_OFF + VEC_BYTES;
  static constexpr int STAT_BYTES = 8192;                    // LayerNorm row statistics exchanged between epilogue warps / cluster CTAs
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + STAT_BYTES + 1024 /*align slack*/;
  static constexpr int EPI_WARPS = EW;                       // EW / 4 warps per TMEM lane quarter; warp group g takes the 32-column chunks g, g + EW/4, ...
  static constexpr int CGROUPS = EW / 4;
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr int TILE_LD = 36;                         // floats; 144 B rows keep float4 accesses conflict-free
  static constexpr int TILE_BYTES = 32 * TILE_LD * 4;        // one warp-private 32x32 fp32 tile
  static_assert(EW * TILE_BYTES <= STAGES * STAGE_BYTES, "epilogue tiles must fit in the dead stage memory");
  static_assert(EW == 8 || EW == 16, "8 epilogue warps (latency-bound launches) or 16 (multi-wave launches: the epilogue is issue / latency bound)");
};

// slow path: per-thread scalar stores (scatter / unaligned / partial column chunk)
__device__ __forceinline__ void tc_store32_slow(const LinArgs& p, long drow, int col0, const float (&y)[32]) {
  for (int j = 0; j < 32; ++j)
    if (col0 + j < p.n_store) act_store(p.out, p.out_planes, drow, col0 + j, y[j]);
}

// Warp-cooperative, coalesced load of a [32 rows x 32 cols] fp32 block into the warp-private tile, then every thread
// picks up its own row.  Rows row_base .. row_base+nvalid-1 of `base` (row stride ld) are read, the rest is zero.
// When `idx` is given the source row of local row rr is idx[row_base + rr] (broadcast add of per-sequence vectors).
__device__ __forceinline__ void warp_load_rows(float (*tile)[36], const float* __restrict__ base, long ld,
                                               const int* __restrict__ idx, long row_base, int nvalid, int col0, int lane,
                                               float (&v)[32]) {
  float4 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr < nvalid) {
      const long sr = idx ? static_cast<long>(__ldg(idx + row_base + rr)) : row_base + rr;
      a[i] = *reinterpret_cast<const float4*>(base + sr * ld + col0 + (lane & 7) * 4);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(&tile[i * 4 + (lane >> 3)][(lane & 7) * 4]) = a[i];
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(&tile[lane][4 * q]);
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
  __syncwarp();
}

// The same load in two halves, so that the global round trip of chunk c + 1 overlaps the arithmetic of chunk c (and, for the
// first chunk, the mainloop): issue() only requests the 8 float4 per lane, finish() transposes them through the warp tile.
__device__ __forceinline__ void warp_load_issue(float4 (&a)[8], const float* __restrict__ base, long ld, long row_base, int nvalid,
                                                int col0, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rr < nvalid) a[i] = *reinterpret_cast<const float4*>(base + (row_base + rr) * ld + col0 + (lane & 7) * 4);
  }
}
__device__ __forceinline__ void warp_load_finish(float (*tile)[36], const float4 (&a)[8], int lane, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(&tile[i * 4 + (lane >> 3)][(lane & 7) * 4]) = a[i];
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(&tile[lane][4 * q]);
    v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
  }
  __syncwarp();
}

// Warp-cooperative, coalesced store of 32 finished rows x 32 cols (thread <-> row) through the warp-private tile.
// Every store instruction writes whole 128-byte lines (fp32) / whole 32-byte sectors (bf16 planes).
__device__ __forceinline__ void warp_store_act(float (*tile)[36], const Act& out, int out_planes, long row_base, int nvalid, int col0,
                                               int lane, const float (&y)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(&tile[lane][4 * q]) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
  __syncwarp();
  float4 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(&tile[i * 4 + (lane >> 3)][(lane & 7) * 4]);
  const long off0 = (row_base + (lane >> 3)) * out.ld + col0 + (lane & 7) * 4;
  const long step = 4L * out.ld;
  if (out.f32) {
    float* d = out.f32 + off0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i * 4 + (lane >> 3) < nvalid) *reinterpret_cast<float4*>(d + i * step) = a[i];
  }
  if (out.pl && out_planes > 0) {
    op16* dh = out.pl + off0;
    op16* dl = dh + static_cast<long>(out.rows_alloc) * out.ld;
    const bool two = out_planes > 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t h01, l01, h23, l23;
      split2_op(a[i].x, a[i].y, out_planes, h01, l01);
      split2_op(a[i].z, a[i].w, out_planes, h23, l23);
      if (i * 4 + (lane >> 3) < nvalid) {
        *reinterpret_cast<uint2*>(dh + i * step) = make_uint2(h01, h23);
        if (two) *reinterpret_cast<uint2*>(dl + i * step) = make_uint2(l01, l23);
      }
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_store_rows(float (*tile)[36], const LinArgs& p, long row_base, int nvalid, int col0,
                                                int lane, const float (&y)[32]) {
  warp_store_act(tile, p.out, p.out_planes, row_base, nvalid, col0, lane, y);
}

template <int BN, int NSPLIT, int EPI, int EW = 8>
__global__ void __launch_bounds__(TcCfg<BN, NSPLIT, EW>::THREADS, TcCfg<BN, NSPLIT, EW>::MIN_CTAS)
k_linear_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmW, const LinArgs p) {
  using C = TcCfg<BN, NSPLIT, EW>;
  constexpr int STAGES = C::STAGES;
  constexpr bool LN = (EPI == EPI_LN || EPI == EPI_LN_MOD_SILU);
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int tile_m = blockIdx.x;
  tc::pdl_launch_dependents();
  if (tile_m * C::BM >= M || (p.dbg_flags & 2)) return;  // CTA-uniform, before any barrier / allocation
  const int n0 = blockIdx.y * BN;
  const int nkb = p.K / C::BK;
  const int nkb1 = p.K1 / C::BK;
  const int nkb2 = p.K2 / C::BK;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle; pointer arithmetic (no integer round trip) keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  float* vec = reinterpret_cast<float*>(smem + C::VEC_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  // warp index through a shuffle: provably warp-uniform, so the role branches below are uniform and the TMA / MMA operands
  // stay in uniform registers (a lane-0 branch makes ptxas wrap every UTCHMMA / UTMALDG in an ELECT + R2UR waterfall loop)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  long long* dbg = p.dbg ? p.dbg + (blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
#define STAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) STAMP(0);
  if (threadIdx.x == 0) trace_mark(p.trace, 0);

  // weights never depend on the previous grid, activations do: the W half of a stage may be requested before pdl_wait
  auto produce_w = [&](int kb) {
    const int s = kb % STAGES;
    tc::mbar_expect_tx(&full[s], C::STAGE_BYTES);
    uint8_t* st = smem + s * C::STAGE_BYTES;
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl)
      tc::tma_load_2d(st + NSPLIT * C::A_BYTES + pl * C::W_BYTES, &tmW, &full[s], kb * C::BK, w_plane<NSPLIT>(pl) * p.w_plane_rows + n0);
  };
  auto produce_a = [&](int kb) {
    const int s = kb % STAGES;
    uint8_t* st = smem + s * C::STAGE_BYTES;
    const int src = kb < nkb1 ? 0 : (kb < nkb2 ? 1 : 2);
    const CUtensorMap* ma = src == 0 ? &tmA : (src == 1 ? &tmA2 : &tmA3);
    const int kcol = (src == 0 ? kb : (src == 1 ? kb - nkb1 : kb - nkb2)) * C::BK;
    const int prow = src == 0 ? p.a_plane_rows : (src == 1 ? p.a2_plane_rows : p.a3_plane_rows);
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl) tc::tma_load_2d(st + pl * C::A_BYTES, ma, &full[s], kcol, pl * prow + tile_m * C::BM);
  };

  const int npre = nkb < STAGES ? nkb : STAGES;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        tc::mbar_init(&full[s], 1);
        tc::mbar_init(&empty[s], 1);
      }
      tc::mbar_init(accum_full, 1);
      tc::fence_barrier_init();
      tc::fence_proxy_async();
    }
    __syncwarp();
    // the first ring pass needs no consumer hand-shake: start the loads before the CTA-wide setup barrier
    if (tc::elect_one())
      for (int kb = 0; kb < npre; ++kb) produce_w(kb);
    __syncwarp();
    tc::pdl_wait();
    if (lane == 0) trace_mark(p.trace, 1);
    if (tc::elect_one())
      for (int kb = 0; kb < npre; ++kb) produce_a(kb);
    if (lane == 0) STAMP(2);
    __syncwarp();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, BN);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) STAMP(1);

  if (warp == 0) {
    // ===== TMA producer (remaining k-blocks): the whole warp runs the uniform loop, one elected lane issues =====
    for (int kb = npre; kb < nkb; ++kb) {
      tc::mbar_wait(&empty[kb % STAGES], ((kb / STAGES) & 1) ^ 1);
      if (tc::elect_one()) {
        produce_w(kb);
        produce_a(kb);
      }
      __syncwarp();
    }
    if (lane == 0) STAMP(3);
    __syncwarp();  // reconverge before the CTA barrier (bar.sync counts per warp)
  } else if (warp == 1) {
    {
      // ===== MMA issuer (whole warp in the uniform loop, one elected lane issues) =====
      constexpr uint32_t idesc = tc::idesc_op<NSPLIT>(C::BM, BN);
      const uint32_t smem_u = tc::smem_u32(smem);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        if (lane == 0) {
          if (kb == 0) STAMP(4);
          if (kb == nkb - 1) STAMP(5);
        }
        const uint32_t sa = smem_u + s * C::STAGE_BYTES;
        const uint32_t sw = sa + NSPLIT * C::A_BYTES;
        if (tc::elect_one()) {
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
        if (kb == nkb - 1) tc::mma_commit(accum_full);
        }
        __syncwarp();
      }
      if (lane == 0) STAMP(6);
    }
    __syncwarp();
  } else {
    // ===== epilogue: EW warps; thread <-> accumulator row, the EW / 4 warps of a TMEM lane quarter split the 32-column chunks
    // between them (chunk c belongs to column group c % (EW / 4)) =====
    // (1) while the mainloop runs: per-column vectors -> shared memory (read back as broadcast float4)
    const int et = threadIdx.x - 64;
    for (int i = et; i < BN; i += C::EPI_THREADS) vec[i] = (p.bias && n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
    if (LN) {
      for (int i = et; i < 256; i += C::EPI_THREADS) {
        vec[256 + i] = __ldg(p.ln_g + i);
        vec[512 + i] = __ldg(p.ln_b + i);
      }
    }
    tc::pdl_wait();  // everything below may belong to earlier grids (modulation table, residuals, outputs)
    if (EPI == EPI_LN_MOD_SILU) {
      for (int i = et; i < 256; i += C::EPI_THREADS) {
        vec[768 + i] = 1.f + p.mod[i];
        vec[1024 + i] = p.mod[256 + i];
      }
    }
    constexpr int CG = C::CGROUPS;
    const int ew = warp - 2;  // 0..EW-1
    const int wq = warp & 3;  // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int ch = ew >> 2;   // column group
    const int r = wq * 32 + lane;
    const long row = static_cast<long>(tile_m) * C::BM + r;
    const bool valid = row < M;
    const long row_base = static_cast<long>(tile_m) * C::BM + wq * 32;
    const int nvalid = static_cast<int>(min(32L, static_cast<long>(M) - row_base));  // may be <= 0
    const long drow = (valid && p.row_map) ? p.row_map[row] : row;
    float* xs = reinterpret_cast<float*>(smem + C::BAR_OFF + 256);  // [2 passes][CG column groups][128 rows] LayerNorm partials
    asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_THREADS) : "memory");  // the epilogue warps only
    // residual of the first LayerNorm chunk: requested while the mainloop runs
    float4 rpre[8];
    const bool ln_res = LN && EPI == EPI_LN && p.res != nullptr;
    if (ln_res) warp_load_issue(rpre, p.res, p.ldres, row_base, nvalid, ch * 32, lane);
    // (2) accumulator ready
    tc::mbar_wait(accum_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) STAMP(7);
    if (threadIdx.x == 64) trace_mark(p.trace, 2);
    float (*tile)[36] = reinterpret_cast<float (*)[36]>(smem + ew * C::TILE_BYTES);   // aliases dead stage memory
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const bool fast = (p.row_map == nullptr) && ((p.out.ld & 7) == 0);
    float v[32], t[32];
    if (LN) {
      // BN == 256 == N: the row lives in two threads (column halves).  3 passes over TMEM (exact two-pass variance),
      // partial sums exchanged through shared memory.
      float s = 0.f;
#pragma unroll 1
      for (int c = ch; c < BN / 32; c += CG) {
        if (ln_res) {
          warp_load_finish(tile, rpre, lane, t);
          if (c + CG < BN / 32) warp_load_issue(rpre, p.res, p.ldres, row_base, nvalid, (c + CG) * 32, lane);   // next chunk in flight
        }
        tc::tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(vec + c * 32 + 4 * q);
          v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
        }
        if (ln_res) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += t[j];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) s += v[j];
        tc::tmem_st32(trow + c * 32, v);
      }
      xs[ch * 128 + r] = s;
      asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_THREADS) : "memory");
      float msum = 0.f;
#pragma unroll
      for (int k = 0; k < CG; ++k) msum += xs[k * 128 + r];
      const float mean = msum * (1.f / 256.f);
      if (threadIdx.x == 64) STAMP(10);
      float q2 = 0.f;
#pragma unroll 1
      for (int c = ch; c < BN / 32; c += CG) {
        tc::tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float d = v[j] - mean;
          q2 += d * d;
        }
      }
      xs[(CG + ch) * 128 + r] = q2;
      asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_THREADS) : "memory");
      float qsum = 0.f;
#pragma unroll
      for (int k = 0; k < CG; ++k) qsum += xs[(CG + k) * 128 + r];
      const float rstd = 1.0f / sqrtf(qsum * (1.f / 256.f) + LD_EPS);
      if (threadIdx.x == 64) STAMP(11);
#pragma unroll 1
      for (int c = ch; c < BN / 32; c += CG) {
        if (EPI == EPI_LN && p.addv) warp_load_rows(tile, p.addv, p.ld_add, p.add_idx, row_base, nvalid, c * 32, lane, t);
        tc::tmem_ld32(trow + c * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 g4 = *reinterpret_cast<const float4*>(vec + 256 + c * 32 + 4 * q);
          const float4 b4 = *reinterpret_cast<const float4*>(vec + 512 + c * 32 + 4 * q);
          v[4 * q] = (v[4 * q] - mean) * rstd * g4.x + b4.x;
          v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * g4.y + b4.y;
          v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * g4.z + b4.z;
          v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * g4.w + b4.w;
        }
        if (EPI == EPI_LN_MOD_SILU) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 s4 = *reinterpret_cast<const float4*>(vec + 768 + c * 32 + 4 * q);
            const float4 h4 = *reinterpret_cast<const float4*>(vec + 1024 + c * 32 + 4 * q);
            v[4 * q] = silu(v[4 * q] * s4.x + h4.x);
            v[4 * q + 1] = silu(v[4 * q + 1] * s4.y + h4.y);
            v[4 * q + 2] = silu(v[4 * q + 2] * s4.z + h4.z);
            v[4 * q + 3] = silu(v[4 * q + 3] * s4.w + h4.w);
          }
        } else if (p.addv) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += t[j];
        }
        if (fast) warp_store_rows(tile, p, row_base, nvalid, c * 32, lane, v);
        else if (valid) tc_store32_slow(p, drow, c * 32, v);
      }
    } else {
#pragma unroll 1
      for (int c = ch; c < BN / 32; c += CG) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.n_store) break;  // warp-uniform
        const bool full_chunk = fast && (col0 + 32 <= p.n_store);
        if (EPI == EPI_RES) {
          if (full_chunk) {
            warp_load_rows(tile, p.res, p.ldres, nullptr, row_base, nvalid, col0, lane, t);
          } else {
            for (int j = 0; j < 32; ++j) t[j] = (valid && col0 + j < p.N) ? p.res[row * p.ldres + col0 + j] : 0.f;
          }
        }
        if (threadIdx.x == 64 && c == 0) STAMP(10);
        tc::tmem_ld32(trow + c * 32, v);
        if (threadIdx.x == 64 && c == 0) STAMP(11);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = *reinterpret_cast<const float4*>(vec + c * 32 + 4 * q);
          v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (EPI == EPI_RELU) v[j] = fmaxf(v[j], 0.f);
          else if (EPI == EPI_GELU) v[j] = gelu_erf_fast(v[j]);
          else if (EPI == EPI_SILU) v[j] = silu(v[j]);
          else if (EPI == EPI_RES) v[j] += t[j];
        }
        if (threadIdx.x == 64 && c == 0) STAMP(12);
        if (full_chunk) warp_store_rows(tile, p, row_base, nvalid, col0, lane, v);
        else if (valid) tc_store32_slow(p, drow, col0, v);
        if (threadIdx.x == 64 && c == 0) STAMP(13);
      }
    }
    if (threadIdx.x == 64) STAMP(8);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p.trace, 3);
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, BN);
  }
#undef STAMP
}

// ------------------------------------------------------------------------------------------------
// LayerNorm-epilogue linears (N == 256) split over a cluster of CL CTAs along N: CTA `rank` owns 256/CL output columns
// (its own W slice, TMEM slice and epilogue columns) and the per-row LayerNorm statistics are exchanged through
// distributed shared memory (two exact passes: sum -> mean, centred sum of squares -> rstd).  CL x more SMs work on
// every 128-row tile than with a whole-row CTA, which is what the latency-bound denoiser loop (10 row tiles) needs.
template <int CL, int NSPLIT, int EPI>
__global__ void __cluster_dims__(1, CL, 1) __launch_bounds__(TcCfg<256 / CL, NSPLIT>::THREADS, 1)
k_linear_tc_ln(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmW, const LinArgs p) {
  constexpr int BN = 256 / CL;
  static_assert(BN == 64 || BN == 128, "cluster split supports 2 or 4 CTAs");
  using C = TcCfg<BN, NSPLIT>;
  constexpr int STAGES = C::STAGES;
  constexpr int NCH = BN / 32;
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int tile_m = blockIdx.x;
  tc::pdl_launch_dependents();
  if (tile_m * C::BM >= M || (p.dbg_flags & 2)) return;  // cluster-uniform
  const uint32_t rank = tc::cluster_ctarank();
  if (threadIdx.x == 0) trace_mark(p.trace, 0);
  const int n0 = static_cast<int>(rank) * BN;
  const int nkb = p.K / C::BK;
  const int nkb1 = p.K1 / C::BK;
  const int nkb2 = p.K2 / C::BK;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  float* vec = reinterpret_cast<float*>(smem + C::VEC_OFF);   // [0,BN) bias | [256,..) g | [512,..) b | [768,..) 1+scale | [1024,..) shift
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* accum_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
  // row statistics exchanged across the cluster: [2 passes][CL ranks][128 rows]; lives after the epilogue tiles in the
  // (by then dead) stage memory is NOT possible -- peers write it while our mainloop may still run -> own region
  float* stat = reinterpret_cast<float*>(smem + C::BAR_OFF + 256);   // 2 passes * 2 CL slots * 128 floats (<= 8 KB)

  // warp index through a shuffle: provably warp-uniform, so the role branches below are uniform and the TMA / MMA operands
  // stay in uniform registers (a lane-0 branch makes ptxas wrap every UTCHMMA / UTMALDG in an ELECT + R2UR waterfall loop)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  // weights never depend on the previous grid, activations do: the W half of a stage may be requested before pdl_wait
  auto produce_w = [&](int kb) {
    const int s = kb % STAGES;
    tc::mbar_expect_tx(&full[s], C::STAGE_BYTES);
    uint8_t* st = smem + s * C::STAGE_BYTES;
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl)
      tc::tma_load_2d(st + NSPLIT * C::A_BYTES + pl * C::W_BYTES, &tmW, &full[s], kb * C::BK, w_plane<NSPLIT>(pl) * p.w_plane_rows + n0);
  };
  auto produce_a = [&](int kb) {
    const int s = kb % STAGES;
    uint8_t* st = smem + s * C::STAGE_BYTES;
    const int src = kb < nkb1 ? 0 : (kb < nkb2 ? 1 : 2);
    const CUtensorMap* ma = src == 0 ? &tmA : (src == 1 ? &tmA2 : &tmA3);
    const int kcol = (src == 0 ? kb : (src == 1 ? kb - nkb1 : kb - nkb2)) * C::BK;
    const int prow = src == 0 ? p.a_plane_rows : (src == 1 ? p.a2_plane_rows : p.a3_plane_rows);
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl) tc::tma_load_2d(st + pl * C::A_BYTES, ma, &full[s], kcol, pl * prow + tile_m * C::BM);
  };

  const int npre = nkb < STAGES ? nkb : STAGES;
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        tc::mbar_init(&full[s], 1);
        tc::mbar_init(&empty[s], 1);
      }
      tc::mbar_init(accum_full, 1);
      tc::fence_barrier_init();
      tc::fence_proxy_async();
    }
    __syncwarp();
    if (tc::elect_one())
      for (int kb = 0; kb < npre; ++kb) produce_w(kb);
    __syncwarp();
    tc::pdl_wait();
    if (lane == 0) trace_mark(p.trace, 1);
    if (tc::elect_one())
      for (int kb = 0; kb < npre; ++kb) produce_a(kb);
    __syncwarp();
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
    for (int kb = npre; kb < nkb; ++kb) {
      tc::mbar_wait(&empty[kb % STAGES], ((kb / STAGES) & 1) ^ 1);
      if (tc::elect_one()) {
        produce_w(kb);
        produce_a(kb);
      }
      __syncwarp();
    }
    __syncwarp();
    tc::cluster_sync();
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = tc::idesc_op<NSPLIT>(C::BM, BN);
      const uint32_t smem_u = tc::smem_u32(smem);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        tc::mbar_wait(&full[s], (kb / STAGES) & 1);
        tc::tc_fence_after();
        const uint32_t sa = smem_u + s * C::STAGE_BYTES;
        const uint32_t sw = sa + NSPLIT * C::A_BYTES;
        if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < C::BK / C::UMMA_K; ++kk) {
          const uint32_t koff = kk * C::UMMA_K * 2;
          const uint64_t a_hi = tc::smem_desc_sw128(sa + koff);
          const uint64_t w_hi = tc::smem_desc_sw128(sw + koff);
          const uint32_t acc0 = (kb | kk) != 0;
          if (NSPLIT == 1) {
            tc::mma_bf16_ss(tmem_base, a_hi, w_hi, idesc, acc0);
          } else {
            const uint64_t a_lo = tc::smem_desc_sw128(sa + C::A_BYTES + koff);
            const uint64_t w_lo = tc::smem_desc_sw128(sw + C::W_BYTES + koff);
            tc::mma_bf16_ss(tmem_base, a_lo, w_hi, idesc, acc0);
            tc::mma_bf16_ss(tmem_base, a_hi, w_lo, idesc, 1u);
            tc::mma_bf16_ss(tmem_base, a_hi, w_hi, idesc, 1u);
          }
        }
        tc::mma_commit(&empty[s]);
        if (kb == nkb - 1) tc::mma_commit(accum_full);
        }
        __syncwarp();
      }
    }
    __syncwarp();
    tc::cluster_sync();
  } else {
    const int et = threadIdx.x - 64;
    for (int i = et; i < BN; i += C::EPI_THREADS) {
      vec[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
      vec[256 + i] = __ldg(p.ln_g + n0 + i);
      vec[512 + i] = __ldg(p.ln_b + n0 + i);
    }
    tc::pdl_wait();
    if (EPI == EPI_LN_MOD_SILU) {
      for (int i = et; i < BN; i += C::EPI_THREADS) {
        vec[768 + i] = 1.f + p.mod[n0 + i];
        vec[1024 + i] = p.mod[256 + n0 + i];
      }
    }
    const int ew = warp - 2;
    const int wq = warp & 3;
    const int ch = ew >> 2;        // column half inside this CTA's BN columns: chunks ch, ch+2, ...
    constexpr int NCHT = NCH / 2;  // chunks per thread
    static_assert(NCH % 2 == 0, "two epilogue warps per lane quarter");
    const int r = wq * 32 + lane;
    const long row = static_cast<long>(tile_m) * C::BM + r;
    const bool valid = row < M;
    const long row_base = static_cast<long>(tile_m) * C::BM + wq * 32;
    const int nvalid = static_cast<int>(min(32L, static_cast<long>(M) - row_base));
    // (1) while the mainloop runs: this thread's residual / broadcast-add values straight into registers (each thread reads
    // its own 128-byte row segment; the lines stay in L1 across the 8 vector loads), so that nothing but the TMEM read and the
    // statistics exchange is left on the critical path once the accumulator is ready
    float t[NCHT][32], u[NCHT][32];
#pragma unroll
    for (int ci = 0; ci < NCHT; ++ci) {
      const int c = ch + 2 * ci;
      if (EPI == EPI_LN && p.res) {
        const float4* src = reinterpret_cast<const float4*>(p.res + row * p.ldres + n0 + c * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          t[ci][4 * q] = x.x; t[ci][4 * q + 1] = x.y; t[ci][4 * q + 2] = x.z; t[ci][4 * q + 3] = x.w;
        }
      }
      if (EPI == EPI_LN && p.addv) {
        const long sr = valid ? static_cast<long>(__ldg(p.add_idx + row)) : 0;
        const float4* src = reinterpret_cast<const float4*>(p.addv + sr * p.ld_add + n0 + c * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
          u[ci][4 * q] = x.x; u[ci][4 * q + 1] = x.y; u[ci][4 * q + 2] = x.z; u[ci][4 * q + 3] = x.w;
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    tc::mbar_wait(accum_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) trace_mark(p.trace, 2);
    float (*tile)[36] = reinterpret_cast<float (*)[36]>(smem + ew * C::TILE_BYTES);
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float v[NCHT][32];
    // (2) v = acc + bias (+ residual); local mean and centred sum of squares over this thread's 32 * NCHT columns
    float s = 0.f;
#pragma unroll
    for (int ci = 0; ci < NCHT; ++ci) {
      const int c = ch + 2 * ci;
      tc::tmem_ld32(trow + c * 32, v[ci]);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(vec + c * 32 + 4 * q);
        v[ci][4 * q] += b4.x; v[ci][4 * q + 1] += b4.y; v[ci][4 * q + 2] += b4.z; v[ci][4 * q + 3] += b4.w;
      }
      if (EPI == EPI_LN && p.res) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[ci][j] += t[ci][j];
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) s += v[ci][j];
    }
    constexpr float NLOC = 32.f * NCHT;
    const float mloc = s * (1.f / NLOC);
    float m2 = 0.f;
#pragma unroll
    for (int ci = 0; ci < NCHT; ++ci)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[ci][j] - mloc;
        m2 += d * d;
      }
    // (3) ONE exchange of (local mean, local M2) between the 2 * CL column groups of the row (Chan et al. pairwise
    // combination: exact two-pass variance inside a group, group means combined afterwards), pushed into every CTA
    const uint32_t stat_addr = tc::smem_u32(stat);
    const uint32_t slot = rank * 2 + ch;
#pragma unroll
    for (int k = 0; k < CL; ++k) {
      tc::st_cluster_f32(tc::mapa(stat_addr + ((0 * 2 * CL + slot) * 128 + r) * 4, k), mloc);
      tc::st_cluster_f32(tc::mapa(stat_addr + ((1 * 2 * CL + slot) * 128 + r) * 4, k), m2);
    }
    tc::cluster_sync();
    float msum = 0.f, qsum = 0.f, mk[2 * CL];
#pragma unroll
    for (int k = 0; k < 2 * CL; ++k) {
      mk[k] = stat[(0 * 2 * CL + k) * 128 + r];
      msum += mk[k];
      qsum += stat[(1 * 2 * CL + k) * 128 + r];
    }
    const float mean = msum * (1.f / (2 * CL));
#pragma unroll
    for (int k = 0; k < 2 * CL; ++k) qsum += NLOC * (mk[k] - mean) * (mk[k] - mean);
    const float rstd = 1.0f / sqrtf(qsum * (1.f / 256.f) + LD_EPS);
    // (4) normalise this thread's columns and store
#pragma unroll
    for (int ci = 0; ci < NCHT; ++ci) {
      const int c = ch + 2 * ci;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 g4 = *reinterpret_cast<const float4*>(vec + 256 + c * 32 + 4 * q);
        const float4 b4 = *reinterpret_cast<const float4*>(vec + 512 + c * 32 + 4 * q);
        v[ci][4 * q] = (v[ci][4 * q] - mean) * rstd * g4.x + b4.x;
        v[ci][4 * q + 1] = (v[ci][4 * q + 1] - mean) * rstd * g4.y + b4.y;
        v[ci][4 * q + 2] = (v[ci][4 * q + 2] - mean) * rstd * g4.z + b4.z;
        v[ci][4 * q + 3] = (v[ci][4 * q + 3] - mean) * rstd * g4.w + b4.w;
      }
      if (EPI == EPI_LN_MOD_SILU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s4 = *reinterpret_cast<const float4*>(vec + 768 + c * 32 + 4 * q);
          const float4 h4 = *reinterpret_cast<const float4*>(vec + 1024 + c * 32 + 4 * q);
          v[ci][4 * q] = silu(v[ci][4 * q] * s4.x + h4.x);
          v[ci][4 * q + 1] = silu(v[ci][4 * q + 1] * s4.y + h4.y);
          v[ci][4 * q + 2] = silu(v[ci][4 * q + 2] * s4.z + h4.z);
          v[ci][4 * q + 3] = silu(v[ci][4 * q + 3] * s4.w + h4.w);
        }
      } else if (p.addv) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[ci][j] += u[ci][j];
      }
      warp_store_rows(tile, p, row_base, nvalid, n0 + c * 32, lane, v[ci]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p.trace, 3);
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, BN);
  }
}
