// Ragged self-attention of the LA-VAE decoder / encoder on the 5th-generation tensor cores (operator/cross_attention.py:299-300,
// 366-371: nn.MultiheadAttention, 4 heads of 64, key-padding mask = the L valid rows of the sequence, L <= 256).
//
// CTA = (head, sequence).  Q, K and V^T of the (sequence, head) are converted ONCE from the fp32 in-projection buffer into 16-bit
// operand tiles in shared memory (K-major, 128-byte swizzle -- one row is exactly the 64-wide head, i.e. one swizzle atom), then per
// 128-query tile:
//     S = Q K^T            tcgen05.mma  M = 128, N = Lp (keys, multiple of 16), K = 64, accumulator in TMEM columns [0, Lp)
//     P = softmax(S)       two threads per query row (TMEM lane; warps w and w + 4 take alternate 32-column chunks): row max, exp2,
//                          row sum in fp32 registers; P goes back to TENSOR MEMORY as packed 16-bit pairs (tcgen05.st)
//     O = P V              tcgen05.mma with the A operand read from TMEM, B = V^T tiles; accumulator in TMEM, scaled by 1 / sum
// Scores never touch shared memory or registers beyond one 32-column chunk.  NSPLIT = 2 forms every product from fp16 hi / lo split
// operands (lo*hi + hi*lo + hi*hi, fp32 accumulate) like the x3 GEMMs; NSPLIT = 1 is the plain bf16 mode.  Nothing is computed or
// stored for padded rows; keys in [L, Lp) are zero rows whose probabilities are forced to 0.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

#define AT5_MAXL 256

template <int NSPLIT>
struct At5Cfg {
  static constexpr int THREADS = 256;                       // 8 warps: two per TMEM lane quarter of the 128-query tile
  static constexpr int Q_BYTES = 2 * 128 * 128;             // two query tiles x 128 rows x 128 B          (per plane)
  static constexpr int K_BYTES = AT5_MAXL * 128;            // keys x 128 B                                (per plane)
  static constexpr int V_BYTES = (AT5_MAXL / 64) * 64 * 128;  // V^T: key blocks of 64 x [64 head dims x 128 B] (per plane)
  static constexpr int RED_BYTES = 2 * 2 * 128 * 4;         // [max | sum][column half][128 rows] softmax partials
  static constexpr int SMEM_BYTES = NSPLIT * (Q_BYTES + K_BYTES + V_BYTES) + RED_BYTES + 64 + 1024 /*alignment slack*/;
  static_assert(V_BYTES == K_BYTES, "the TMA staging path stores V like K: [keys x 128 B]");
  // tensor memory columns: S [0, 256) fp32 | P hi [256, 384) | P lo [384, 512) | O [0, 64) over the (by then dead) scores
  static constexpr int TMEM_COLS = 512, COL_O = 0, COL_PHI = 256, COL_PLO = 384;
};

namespace at5 {
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (here the probabilities) is read from tensor memory
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-bit element (row, col) inside a K-major tile with 128-byte rows and the 128-byte swizzle (col < 64)
__device__ __forceinline__ uint32_t sw128(int row, int col) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) * 2));
}
// NR rows per thread-batch: rows x 64 columns of an fp32 matrix (row stride 768) -> K-major swizzled 16-bit tile(s).  One CTA per SM
// means the staging is a pure latency problem: every thread first ISSUES all its 16-byte loads of the phase (lane <-> 4 columns,
// 16 lanes per row, ITEMS float4 per thread = up to 64 KB in flight per SM), then converts and stores (8-byte stores, two rows per
// warp instruction -> 2-way bank conflict, negligible).
template <int NSPLIT, int ITEMS>
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, int L, int nrows, float scale, uint8_t* dst, int plane_bytes,
                                           int tid, int nthreads) {
  float4 v[ITEMS];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    const int i = tid + u * nthreads, row = i >> 4, c4 = i & 15;
    v[u] = (row < L) ? __ldg(reinterpret_cast<const float4*>(src + static_cast<long>(row) * 768) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    const int i = tid + u * nthreads, row = i >> 4, c4 = i & 15;
    if (row < nrows) {
      uint32_t h0, l0, h1, l1;
      split2_op<NSPLIT>(v[u].x * scale, v[u].y * scale, h0, l0);
      split2_op<NSPLIT>(v[u].z * scale, v[u].w * scale, h1, l1);
      const uint32_t o = sw128(row, 4 * c4);
      *reinterpret_cast<uint2*>(dst + o) = make_uint2(h0, h1);
      if (NSPLIT == 2) *reinterpret_cast<uint2*>(dst + plane_bytes + o) = make_uint2(l0, l1);
    }
  }
}
}  // namespace at5

// TMAQ = true: q | k | v arrive as 16-bit operand planes written by the in-projection GEMM's epilogue ([planes][rows_alloc][768], the
// tensor map of that buffer has boxes of 128 rows x 64 columns = one head of 128 rows): the three operands of the (sequence, head) are
// pulled by a dozen TMA boxes straight into their swizzled tiles -- no fp32 read, no conversion, no transposition (V is consumed
// as an MN-major B operand: row = key, 128 bytes = the 64 head dims), and the 1/8 of the scores moves into the softmax exponent.
// Rows behind the sequence inside a box belong to the next sequence or to the zero-initialised padding of the buffer: finite values,
// whose keys get probability 0.  TMAQ = false: the fp32 staging path (q k v as one fp32 [rows, 768] buffer).
template <int NSPLIT, bool TMAQ>
__global__ void __launch_bounds__(At5Cfg<NSPLIT>::THREADS, 1)
k_attn_self_t5(const __grid_constant__ CUtensorMap tmQKV, int plane_rows, const float* __restrict__ qkv, const int* __restrict__ foff,
               Act out, int planes, long long* dbg) {
  using C = At5Cfg<NSPLIT>;
#define ASTAMP(i) do { if (dbg && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) dbg[i] = clock64(); } while (0)
  ASTAMP(0);
  pdl_prologue();
  ASTAMP(1);
  const int b = blockIdx.y, h = blockIdx.x;
  const int r0 = foff[b], L = min(foff[b + 1] - r0, AT5_MAXL);
  if (L <= 0) return;
  const int Lp = (L + 15) & ~15;   // MMA N of the score GEMM / K extent of the P V GEMM
  extern __shared__ uint8_t at5_raw[];
  uint8_t* smem = at5_raw + ((1024u - (tc::smem_u32(at5_raw) & 1023u)) & 1023u);
  uint8_t* Qs = smem;                              // [plane][256 x 128 B]
  uint8_t* Ks = Qs + NSPLIT * C::Q_BYTES;          // [plane][256 x 128 B]
  uint8_t* Vt = Ks + NSPLIT * C::K_BYTES;          // [plane][key block][64 x 128 B]
  float* red = reinterpret_cast<float*>(Vt + NSPLIT * C::V_BYTES);   // [max | sum][half][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + 2 * 2 * 128);
  uint64_t* bar_ld = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NW = C::THREADS / 32;

  if (warp == 0) {
    if (lane == 0) {
      tc::mbar_init(bar, 1);
      tc::mbar_init(bar_ld, 1);
      tc::fence_barrier_init();
      tc::fence_proxy_async();
      if (TMAQ) {
        // all operand boxes of this (sequence, head) on one barrier; a box always delivers its full 16 KB (rows past the tensor are zero-filled)
        const int nqt_ = (L + 127) >> 7, nkt_ = (Lp + 127) >> 7;
        tc::mbar_expect_tx(bar_ld, static_cast<uint32_t>(NSPLIT * (nqt_ + 2 * nkt_) * 16384));
        for (int pl = 0; pl < NSPLIT; ++pl) {
          const int prow = pl * plane_rows + r0;
          for (int t = 0; t < nqt_; ++t) tc::tma_load_2d(Qs + pl * C::Q_BYTES + t * 16384, &tmQKV, bar_ld, h * 64, prow + t * 128);
          for (int t = 0; t < nkt_; ++t) {
            tc::tma_load_2d(Ks + pl * C::K_BYTES + t * 16384, &tmQKV, bar_ld, 256 + h * 64, prow + t * 128);
            tc::tma_load_2d(Vt + pl * C::V_BYTES + t * 16384, &tmQKV, bar_ld, 512 + h * 64, prow + t * 128);
          }
        }
      }
    }
    __syncwarp();
    tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc::tmem_relinquish();
  }
  const int nqt = (L + 127) >> 7;                  // query tiles
  if (!TMAQ) {
  // ---- operand staging: q (pre-scaled by 1/8, exact) and k as row-major tiles, v transposed.  Rows >= L are zero.
  const float* base = qkv + static_cast<long>(r0) * 768 + h * 64;
  // (256 query rows + 256 key rows) x 16 float4 / 256 threads = 16 + 16 per thread, all in flight at once
  at5::stage_rows<NSPLIT, 16>(base, L, nqt * 128, 0.125f, Qs, C::Q_BYTES, tid, C::THREADS);
  ASTAMP(21);
  at5::stage_rows<NSPLIT, 16>(base + 256, L, Lp, 1.0f, Ks, C::K_BYTES, tid, C::THREADS);
  ASTAMP(22);
  // V^T: item = (key block of 64, quad of head dims); lane <-> key pair (2 l, 2 l + 1) of the block, so the 32 four-byte stores of an
  // instruction fill one 128-byte row (head dim d) of the block: conflict-free.  The strided fp32 reads (one 16-byte segment of a
  // different row per lane) hit each 128-byte line repeatedly -> L1.  A warp owns the items warp, warp + 8, ...: 8 items of 2 float4.
  const int nkb = (Lp + 63) >> 6;
  {
    float4 va[8], vc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = warp + u * NW, kb = i >> 4, d0 = (i & 15) * 4, k0 = kb * 64 + 2 * lane;
      va[u] = vc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < nkb * 16) {
        if (k0 < L) va[u] = __ldg(reinterpret_cast<const float4*>(base + static_cast<long>(k0) * 768 + 512 + d0));
        if (k0 + 1 < L) vc[u] = __ldg(reinterpret_cast<const float4*>(base + static_cast<long>(k0 + 1) * 768 + 512 + d0));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = warp + u * NW, kb = i >> 4, d0 = (i & 15) * 4;
      if (i < nkb * 16) {
        const float a[4] = {va[u].x, va[u].y, va[u].z, va[u].w}, c[4] = {vc[u].x, vc[u].y, vc[u].z, vc[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hi, lo;
          split2_op<NSPLIT>(a[j], c[j], hi, lo);
          const uint32_t o = kb * 8192 + at5::sw128(d0 + j, 2 * lane);
          *reinterpret_cast<uint32_t*>(Vt + o) = hi;
          if (NSPLIT == 2) *reinterpret_cast<uint32_t*>(Vt + C::V_BYTES + o) = lo;
        }
      }
    }
  }
  } else {
    __syncthreads();                // the barrier was initialised (and the loads issued) by one lane of warp 0
    tc::mbar_wait(bar_ld, 0);
  }
  ASTAMP(2);
  tc::fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  ASTAMP(3);
  const uint32_t tmem = *tmem_slot;
  const uint32_t qs_u = tc::smem_u32(Qs), ks_u = tc::smem_u32(Ks), vt_u = tc::smem_u32(Vt);
  // TMAQ: V tiles are [key][64 head dims] = an MN-major B operand (instruction descriptor bit 16), 8-key groups 1024 B apart
  const uint32_t idesc_s = tc::idesc_op<NSPLIT>(128, Lp), idesc_o = tc::idesc_op<NSPLIT>(128, 64) | (TMAQ ? (1u << 16) : 0u);
  const int wq = warp & 3, half = warp >> 2;                          // TMEM lane quarter / column half of this warp
  const int row = wq * 32 + lane;                                     // query row inside the tile
  const uint32_t tlane = static_cast<uint32_t>(wq * 32) << 16;
  const float LOG2E = TMAQ ? 0.125f * 1.4426950408889634f : 1.4426950408889634f;   // TMAQ: q is not pre-scaled, 1/sqrt(64) goes here
  const int nch = (Lp + 31) >> 5;
  uint32_t phase = 0;

  for (int qt = 0; qt < nqt; ++qt) {
    // ---- S = Q K^T
    if (warp == 0) {
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t q_hi = tc::smem_desc_sw128(qs_u + qt * 16384 + kk * 32), k_hi = tc::smem_desc_sw128(ks_u + kk * 32);
          if (NSPLIT == 1) {
            tc::mma_bf16_ss(tmem, q_hi, k_hi, idesc_s, kk != 0);
          } else {
            const uint64_t q_lo = tc::smem_desc_sw128(qs_u + C::Q_BYTES + qt * 16384 + kk * 32);
            const uint64_t k_lo = tc::smem_desc_sw128(ks_u + C::K_BYTES + kk * 32);
            tc::mma_bf16_ss(tmem, q_lo, k_hi, idesc_s, kk != 0);
            tc::mma_bf16_ss(tmem, q_hi, k_lo, idesc_s, 1u);
            tc::mma_bf16_ss(tmem, q_hi, k_hi, idesc_s, 1u);
          }
        }
        tc::mma_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::tc_fence_after();
    ASTAMP(4 + 8 * qt);
    // ---- softmax: this thread's chunks (half, half + 2, ...) of its query row; row max / sum combined with the partner warp
    float mx = -INFINITY;
    for (int c = half; c < nch; c += 2) {
      float v[32];
      tc::tmem_ld32(tmem + tlane + c * 32, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < L) mx = fmaxf(mx, v[j]);
    }
    red[half * 128 + row] = mx;
    __syncthreads();
    ASTAMP(5 + 8 * qt);
    const float mxl = fmaxf(red[row], red[128 + row]) * LOG2E;
    float sum = 0.f;
    for (int c = half; c < nch; c += 2) {
      float v[32];
      tc::tmem_ld32(tmem + tlane + c * 32, v);
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int key = c * 32 + 2 * j;
        const float p0 = key < L ? at5::ex2(fmaf(v[2 * j], LOG2E, -mxl)) : 0.f;
        const float p1 = key + 1 < L ? at5::ex2(fmaf(v[2 * j + 1], LOG2E, -mxl)) : 0.f;
        sum += p0 + p1;
        split2_op<NSPLIT>(p0, p1, ph[j], pl[j]);
      }
      at5::tmem_st16(tmem + tlane + C::COL_PHI + c * 16, ph);         // column j of a plane holds keys 2 j, 2 j + 1
      if (NSPLIT == 2) at5::tmem_st16(tmem + tlane + C::COL_PLO + c * 16, pl);
    }
    ASTAMP(6 + 8 * qt);
    red[256 + half * 128 + row] = sum;
    at5::tmem_wait_st();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    ASTAMP(7 + 8 * qt);
    const float inv = 1.0f / (red[256 + row] + red[384 + row]);
    // ---- O = P V   (A = P from tensor memory: 16 keys = 8 columns per MMA; B = V^T tiles); O overwrites the dead scores
    if (warp == 0) {
      if (tc::elect_one()) {
        const int nks = Lp >> 4;
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t vo = TMAQ ? ks * 2048 : (ks >> 2) * 8192 + (ks & 3) * 32;
          const uint64_t v_hi = tc::smem_desc_sw128(vt_u + vo);
          const uint32_t p_hi = tmem + C::COL_PHI + ks * 8;
          if (NSPLIT == 1) {
            at5::mma_ts(tmem + C::COL_O, p_hi, v_hi, idesc_o, ks != 0);
          } else {
            const uint64_t v_lo = tc::smem_desc_sw128(vt_u + C::V_BYTES + vo);
            const uint32_t p_lo = tmem + C::COL_PLO + ks * 8;
            at5::mma_ts(tmem + C::COL_O, p_lo, v_hi, idesc_o, ks != 0);
            at5::mma_ts(tmem + C::COL_O, p_hi, v_lo, idesc_o, 1u);
            at5::mma_ts(tmem + C::COL_O, p_hi, v_hi, idesc_o, 1u);
          }
        }
        tc::mma_commit(bar);
      }
      __syncwarp();
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1;
    tc::tc_fence_after();
    ASTAMP(8 + 8 * qt);
    // ---- normalise and store: this thread's 32 columns (column half) of its output row
    {
      const int qr = qt * 128 + row;
      float v[32];
      tc::tmem_ld32(tmem + tlane + C::COL_O + half * 32, v);
      if (qr < L) {
        const long o = static_cast<long>(r0 + qr) * out.ld + h * 64 + half * 32;
        if (out.f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(out.f32 + o + 4 * j) = make_float4(v[4 * j] * inv, v[4 * j + 1] * inv, v[4 * j + 2] * inv, v[4 * j + 3] * inv);
        }
        if (out.pl && planes > 0 && out.f32) {   // (a tensor with an fp32 master as well: direct row stores)
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split2_op(v[2 * j] * inv, v[2 * j + 1] * inv, planes, hi[j], lo[j]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(out.pl + o + 8 * j) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            if (planes > 1)
              *reinterpret_cast<uint4*>(out.pl + static_cast<long>(out.rows_alloc) * out.ld + o + 8 * j) =
                  make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        }
      }
      // planes only (what the plans use): through the shared memory of this query tile's (dead) Q operand, so that every global
      // store instruction writes whole 128-byte row segments instead of 32 rows x 16 bytes
      if (out.pl && planes > 0 && !out.f32) {
        uint8_t* scr = Qs + qt * 16384;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split2_op(v[2 * j] * inv, v[2 * j + 1] * inv, planes, hi[j], lo[j]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t so = row * 128 + ((((half * 4 + k) ^ (row & 7))) << 4);
          *reinterpret_cast<uint4*>(scr + so) = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
          if (NSPLIT == 2 && planes > 1) *reinterpret_cast<uint4*>(scr + C::Q_BYTES + so) = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 32 + warp * 4 + (lane >> 3), ch = lane & 7;
          if (qt * 128 + rr < L) {
            const uint32_t so = rr * 128 + ((ch ^ (rr & 7)) << 4);
            const long go = static_cast<long>(r0 + qt * 128 + rr) * out.ld + h * 64 + ch * 8;
            *reinterpret_cast<uint4*>(out.pl + go) = *reinterpret_cast<const uint4*>(scr + so);
            if (NSPLIT == 2 && planes > 1)
              *reinterpret_cast<uint4*>(out.pl + static_cast<long>(out.rows_alloc) * out.ld + go) = *reinterpret_cast<const uint4*>(scr + C::Q_BYTES + so);
          }
        }
      }
    }
    ASTAMP(9 + 8 * qt);
    tc::tc_fence_before();
    __syncthreads();          // the next tile's score GEMM overwrites S / O, its softmax the partials: everyone is done reading
    tc::tc_fence_after();
  }
  if (warp == 0) {
    __syncwarp();
    tc::tmem_dealloc(tmem, C::TMEM_COLS);
  }
  ASTAMP(20);
#undef ASTAMP
}
