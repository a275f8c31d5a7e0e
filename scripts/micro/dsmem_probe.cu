// Micro-probe: how fast can the 8 CTAs of a cluster exchange 96 KB each (12 KB to every peer) on B200?
//   mode 0: st.shared::cluster.v4.f32, consecutive lanes -> consecutive 16 B
//   mode 1: st.shared::cluster.v4.f32, consecutive lanes -> 1 KB stride (the row-per-thread pattern of an accumulator dump)
//   mode 2: cp.async.bulk.shared::cluster.shared::cta (one 12 KB bulk copy per peer, remote mbarrier complete_tx)
//   mode 3: st.global (coalesced) + cluster barrier + cp.async.bulk global -> shared (through L2)
//   mode 4: like 2 but 48 bulk copies of 2 KB per CTA (finer pipelining granularity)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int PEER_BYTES = 12 * 1024, CL = 8, TOT = PEER_BYTES * CL;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0; int spins = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\nselp.b32 %0,1,0,P;\n}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    if (++spins > (1 << 24)) __trap();
  }
}
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 1) k_probe(int mode, int reps, long long* out, float* scratch) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* src = sm; uint8_t* dst = sm + TOT; uint64_t* bar = (uint64_t*)(sm + 2 * TOT);
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int t = threadIdx.x, cluster = blockIdx.x / CL;
  for (int i = t; i < TOT / 4; i += 256) ((float*)src)[i] = rank * 1000.f + i;
  if (t == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  csync();
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (mode == 0 || mode == 1) {
      for (int k = 0; k < CL; ++k) {
        const uint32_t d = mapa(s32(dst), (rank + k) % CL) + rank * PEER_BYTES;
        const uint8_t* s = src + ((rank + k) % CL) * PEER_BYTES;
        for (int i = 0; i < PEER_BYTES / 16 / 256; ++i) {
          int idx = mode == 0 ? i * 256 + t : ((t & 31) * 24 + (t >> 5) * 3 + i) ;  // mode 1: lane stride 24*16 B = 384 B
          float4 v = *(const float4*)(s + idx * 16);
          asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(d + idx * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
      csync();
    } else if (mode == 2 || mode == 4) {
      if (t == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(TOT) : "memory");
      }
      csync();  // every receiver armed (stands for barrier #1 of the real kernel)
      if (t == 0) {
        const int piece = mode == 2 ? PEER_BYTES : 2048;
        for (int off = 0; off < PEER_BYTES; off += piece)
          for (int k = 0; k < CL; ++k) {
            const uint32_t pr = (rank + k) % CL;
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(mapa(s32(dst), pr) + rank * PEER_BYTES + off), "r"(s32(src + pr * PEER_BYTES + off)), "r"(piece), "r"(mapa(s32(bar), pr)) : "memory");
          }
      }
      mwait(bar, rep & 1);
    } else {
      float* g = scratch + (size_t)cluster * CL * TOT / 4;
      for (int k = 0; k < CL; ++k) {
        const int pr = (rank + k) % CL;
        float4* gd = (float4*)(g + ((size_t)pr * TOT + rank * PEER_BYTES) / 4);
        const float4* s = (const float4*)(src + pr * PEER_BYTES);
        for (int i = t; i < PEER_BYTES / 16; i += 256) gd[i] = s[i];
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
      csync();
      if (t == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(TOT) : "memory");
        for (int k = 0; k < CL; ++k)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(s32(dst + k * PEER_BYTES)), "l"(g + ((size_t)rank * TOT + k * PEER_BYTES) / 4), "r"(PEER_BYTES), "r"(s32(bar)) : "memory");
      }
      mwait(bar, rep & 1);
      csync();
    }
  }
  long long t1 = clock64();
  csync();
  if (t == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = (long long)((float*)dst)[((rank + 1) % CL) * PEER_BYTES / 4 + 5]; }
}
int main() {
  long long* d; cudaMalloc(&d, 2 * 256 * 8);
  float* scratch; cudaMalloc(&scratch, (size_t)32 * CL * TOT);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TOT + 64);
  const int reps = 8;
  for (int clusters : {1, 10, 18}) for (int mode = 0; mode < 5; ++mode) {
    k_probe<<<clusters * CL, 256, 2 * TOT + 64>>>(mode, reps, d, scratch);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    long long h[2 * 256]; cudaMemcpy(h, d, sizeof(long long) * 2 * clusters * CL, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < clusters * CL; ++i) mx = h[2 * i] > mx ? h[2 * i] : mx;
    printf("clusters %2d mode %d: %lld cycles per exchange of %d KB out per CTA -> %.1f B/clk/CTA (check %lld)\n", clusters, mode, mx / reps, TOT / 1024, (double)TOT * reps / mx, h[1]);
  }
  return 0;
}
