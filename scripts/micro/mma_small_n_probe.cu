// Micro-probe: clocks per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N when both operands come from
// shared memory (SS form), against the A-operand-in-TMEM form.  The "swapped" feed-forward kernel of the denoiser (k_ffn_swap)
// issues M = 128 x N = 48 MMAs: the tensor pipe's floor for that shape is 128 * 48 / 256 = 24 clk, but every MMA re-reads its
// 4 KB A tile (128 weight rows x 16 k) and 1.5 KB B tile from shared memory -- if the operand fetch is bound by the 128 B/clk of
// the shared-memory crossbar, the MMA costs ~44 clk whatever the pipe could do.  N = 96 = [x_hi ; x_lo] concatenated on the N axis
// (one A read for two of the three products of the x3 split) is the candidate fix measured here.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../ladiff_b200/csrc -o mma_small_n_probe mma_small_n_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a_tmem, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS, the same A / B tiles every time; mode 1: SS, A walks over 8 different 16 KB tiles (a weight ring); mode 2: A from TMEM
__global__ void __launch_bounds__(128, 1) k_probe(int N, int mode, int iters, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_tiles = smem;                 // 8 x [128 x 64] 16-bit = 8 x 16 KB
  uint8_t* b_tile = smem + 8 * 16384;      // [256 x 64] 16-bit = 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_tile + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (8 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
    tc::fence_proxy_async();
  }
  if (warp == 1) {
    tc::tmem_alloc(slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    const uint32_t idesc = tc::idesc_f16_f32(128, N);
    const uint32_t a_u = tc::smem_u32(a_tiles), b_u = tc::smem_u32(b_tile);
    long long t0 = 0;
    for (int rep = 0; rep < 2; ++rep) {    // rep 0 warms up
      t0 = clock64();
      if (tc::elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t a_addr = a_u + (mode == 1 ? (i & 7) * 16384 : 0) + kk * 32;
            const uint64_t bd = tc::smem_desc_sw128(b_u + kk * 32);
            if (mode == 2) mma_f16_ts(tmem, tmem + 256 + kk * 8, bd, idesc, 1u);
            else tc::mma_bf16_ss(tmem, tc::smem_desc_sw128(a_addr), bd, idesc, 1u);
          }
        }
        tc::mma_commit(bar);
      }
      __syncwarp();
      tc::mbar_wait(bar, rep & 1);
      tc::tc_fence_after();
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* out; cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 256;
  printf("# clk per tcgen05.mma M=128 K=16 kind::f16 (%d MMAs back to back, one CTA); floor = 128 N / 256\n", iters * 4);
  printf("%-28s %5s %9s %7s %12s\n", "operands", "N", "clk/MMA", "floor", "smem B/clk");
  for (int mode : {0, 1, 2})
    for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
      k_probe<<<1, 128, 180 * 1024>>>(N, mode, iters, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      unsigned long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
      const double per = static_cast<double>(c) / (iters * 4);
      const double bytes = (mode == 2 ? 0 : 4096) + N * 32.0;
      printf("%-28s %5d %9.1f %7.1f %12.1f\n", mode == 0 ? "SS same tiles" : (mode == 1 ? "SS A over 8 tiles" : "A in TMEM, B smem"), N, per, 128.0 * N / 256, bytes / per);
    }
  return 0;
}
