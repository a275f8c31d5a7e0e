// Shared device-side types and helpers.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define LD_D 256      // latent / model width
#define LD_H 4        // heads
#define LD_HD 64      // head dim
#define LD_EPS 1e-5f  // LayerNorm eps (torch default; nn.LayerNorm at cross_attention.py:277, mdiff_transformer.py:145)

// 16-bit tensor-core operand element.  The FORMAT follows the arithmetic mode, i.e. the number of planes a tensor carries:
//   two planes (mode "x3")  : fp16 hi / lo split, v ~= hi + lo to 22 mantissa bits; the three products a_lo w_hi + a_hi w_lo +
//                             a_hi w_hi of the GEMMs then carry ~2^-21 relative error (fp32-grade: the 50-step CFG loop amplifies
//                             operand rounding ~100x, a bf16 hi/lo split (16 bits) measured 1.2e-3 on the B = 128 headline config
//                             against the 1e-3 contract, profiles/r02a).  Range: |v| < 65504 (activations here are LayerNorm'd
//                             streams and O(100) latents; weights O(0.1)); values below 2^-14 use fp16 subnormals (abs error 3e-8).
//   one plane (mode "bf16") : bf16, round to nearest.
typedef uint16_t op16;

// An activation tensor [rows, ld]: fp32 master and/or 16-bit operand planes for the tensor-core GEMMs.
// Planes are stacked: hi plane rows [0, rows_alloc), lo plane rows [rows_alloc, 2*rows_alloc); value ~= hi + lo.
struct Act {
  float* f32;           // may be null
  op16* pl;             // may be null
  int ld;
  int rows_alloc;       // multiple of 128 so a 128-row TMA box never straddles the planes
};

enum Epi : int {
  EPI_BIAS = 0,
  EPI_RELU = 1,
  EPI_GELU = 2,      // exact erf GELU (F.gelu default; cross_attention.py:477-478, mdiff_transformer.py:255)
  EPI_RES = 3,       // res + acc + bias
  EPI_LN = 4,        // LayerNorm((res?) + acc + bias) * g + b  [+ addv[add_idx[row]]]   (N == 256)
  EPI_LN_MOD_SILU = 5,  // SiLU(LayerNorm(acc + bias) * (1 + scale) + shift)              (N == 256) StylizationBlock :161-162
  EPI_SILU = 6
};

// Programmatic dependent launch: every kernel of the sampling plans starts with pdl_prologue() -- signal that the next
// grid may begin its own prologue, then wait until the previous grid's results are visible.
//
// INVARIANT for everything a kernel reads BEFORE its griddepcontrol.wait (the row counts *M_dev, the sequence offsets off[],
// per-column vectors, weights): because every grid signals launch_dependents at entry, pre-wait code can overlap grids several
// launches back, so such data must be complete before the FIRST kernel of the PDL chain starts.  The plans guarantee it
// structurally: counts / offsets / row maps are produced by k_scan_counts + k_fill_rows, which are launched with LAUNCH (a full
// stream dependency, no PDL attribute) ahead of the chain (enqueue_offsets / the plan prologues in ladiff_b200.cu), and every
// table built per call (time / text k-v, modulation, ca_block delta) is written by kernels that precede the step loop and are
// followed by at least one full-dependency launch (the first kernel after a cross-stream join uses g_skip_pdl_once).  A new
// producer of pre-wait data must keep that rule: end the producing section with a non-PDL launch.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
// Branch-free erf-GELU for the tensor-core epilogues (erff's range split serialises the 32 independent elements a thread
// owns).  erf by Abramowitz & Stegun 7.1.26: |erf error| <= 1.5e-7, far below the bf16x3 product error (~1e-5 relative).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;  // 1 / (1 + p z): MUFU.RCP (<= 1 ulp here; __frcp_rn would add a branchy IEEE fix-up path)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = 1.0f - p * t * __expf(-z * z);  // erf(|x|/sqrt2)
  return 0.5f * x * (1.0f + copysignf(e, x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- operand splits -----------------------------------------------------------------------------------------------
// fp16 hi/lo split of two neighbouring columns (x in the low half): F2FP.F16.F32.PACK_AB runs on the full-rate pipes.
__device__ __forceinline__ void split2_f16(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// bf16 round of two neighbouring columns
__device__ __forceinline__ uint32_t pack2_bf16_rn(float x, float y) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// bf16 hi/lo split (decoder / encoder self-attention keeps its mma.sync bf16 fragments: its error is 5e-5 on the features)
__device__ __forceinline__ void split2_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack2_bf16_rn(x, y);
  const float xf = __uint_as_float(hi << 16), yf = __uint_as_float(hi & 0xffff0000u);
  lo = pack2_bf16_rn(x - xf, y - yf);
}
// the operand format of a tensor with NPL planes (see op16): NPL == 2 -> fp16 hi / lo, else bf16 (lo unused)
template <int NPL>
__device__ __forceinline__ void split2_op(float x, float y, uint32_t& hi, uint32_t& lo) {
  if (NPL == 2) {
    split2_f16(x, y, hi, lo);
  } else {
    hi = pack2_bf16_rn(x, y);
    lo = 0u;
  }
}
__device__ __forceinline__ void split2_op(float x, float y, int nplanes, uint32_t& hi, uint32_t& lo) {
  if (nplanes > 1) split2_op<2>(x, y, hi, lo);
  else split2_op<1>(x, y, hi, lo);
}
__device__ __forceinline__ void split_op(float v, int nplanes, op16& hi, op16& lo) {
  if (nplanes > 1) {
    const __half h = __float2half_rn(v);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
  } else {
    hi = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    lo = 0;
  }
}

// nplanes: 0 none, 1 bf16, 2 fp16 hi + lo
__device__ __forceinline__ void act_store(const Act& a, int nplanes, long row, int col, float v) {
  if (a.f32) a.f32[row * a.ld + col] = v;
  if (a.pl && nplanes > 0) {
    op16 hi, lo;
    split_op(v, nplanes, hi, lo);
    a.pl[row * a.ld + col] = hi;
    if (nplanes > 1) a.pl[(static_cast<long>(a.rows_alloc) + row) * a.ld + col] = lo;
  }
}
