// Shared device-side types and helpers.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define LD_D 256      // latent / model width
#define LD_H 4        // heads
#define LD_HD 64      // head dim
#define LD_EPS 1e-5f  // LayerNorm eps (torch default; nn.LayerNorm at cross_attention.py:277, mdiff_transformer.py:145)

// An activation tensor [rows, ld]: fp32 master and/or bf16 operand planes for the tensor-core GEMMs.
// Planes are stacked: hi plane rows [0, rows_alloc), lo plane rows [rows_alloc, 2*rows_alloc); value ~= hi + lo.
struct Act {
  float* f32;           // may be null
  __nv_bfloat16* pl;    // may be null
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

// hi/lo bf16 split of an fp32 value (lo = bf16(v - float(hi))): v ~= hi + lo to ~16 mantissa bits.
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Packed variant for two neighbouring columns (x in the low half): F2FP.BF16.F32.PACK_AB runs on the full-rate pipes,
// the scalar F2F.BF16.F32 conversion does not.  Same round-to-nearest-even results as split_bf16.
__device__ __forceinline__ void split2_bf16(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float xf = __uint_as_float(hi << 16), yf = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - xf, y - yf);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// nplanes: 0 none, 1 hi only, 2 hi + lo
__device__ __forceinline__ void act_store(const Act& a, int nplanes, long row, int col, float v) {
  if (a.f32) a.f32[row * a.ld + col] = v;
  if (a.pl && nplanes > 0) {
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    a.pl[row * a.ld + col] = hi;
    if (nplanes > 1) a.pl[(static_cast<long>(a.rows_alloc) + row) * a.ld + col] = lo;
  }
}
