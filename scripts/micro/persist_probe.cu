// Micro-probe for the "CUDA graph vs persistent kernel" decision (north_star): what does ONE dependency boundary between two
// grid-wide phases cost when the phases live inside one persistent kernel (software grid barrier: release-add on a global counter,
// acquire-spin) compared with a kernel boundary inside a captured graph with programmatic dependent launch (pdl_probe.cu)?
// The reverse loop has 27 such boundaries per step (in-projection -> attention -> feed-forward, 9 layers); between them every
// CTA needs data written by OTHER CTAs (the GEMM tiles cut the rows differently from the token groups), so a barrier + memory
// visibility is the minimum a persistent variant has to pay per boundary.
//   nvcc -arch=sm_100a -O3 -o persist_probe persist_probe.cu && ./persist_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// kind 0: hand-written barrier (red.release.gpu + ld.acquire.gpu spin by one thread, bar.sync around it)
// kind 1: cooperative_groups grid.sync()
__global__ void k_persist(unsigned int* counter, float* data, int phases, int spin_ns, int kind, unsigned long long* out, int* err, int smem_touch) {
  extern __shared__ char sm[];
  if (smem_touch) sm[threadIdx.x] = 1;
  cg::grid_group grid = cg::this_grid();
  const unsigned int G = gridDim.x;
  unsigned long long t_begin = gtime();
  float acc = 0.f;
  for (int p = 0; p < phases; ++p) {
    // "work": consume what another CTA produced in the previous phase, spin, produce
    acc += data[((blockIdx.x + 37) % G) * 32 + (threadIdx.x & 31)];
    const unsigned long long t1 = gtime();
    while (gtime() - t1 < (unsigned long long)spin_ns) {}
    if (threadIdx.x < 32) data[blockIdx.x * 32 + threadIdx.x] = acc + p;
    if (kind == 1) {
      grid.sync();
    } else {
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        const unsigned int target = G * (p + 1);
        unsigned int v = 0;
        const unsigned long long t0 = gtime();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          if (gtime() - t0 > 1000000000ull) { *err = 1; break; }   // 1 s: never hang the box
        } while (v < target);
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) { out[2 * blockIdx.x] = t_begin; out[2 * blockIdx.x + 1] = gtime(); }
  if (acc == 123.456f) out[0] = 0;
}

int main() {
  unsigned int* counter; float* data; unsigned long long* out; int* err;
  cudaMalloc(&counter, 4); cudaMalloc(&data, 148 * 32 * 4); cudaMalloc(&out, 2 * 148 * 8); cudaMalloc(&err, 4);
  cudaMemset(data, 0, 148 * 32 * 4);
  cudaStream_t st; cudaStreamCreate(&st);
  const int phases = 270, spin = 3000;
  cudaFuncSetAttribute(k_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int smem : {0, 190 * 1024})
    for (int G : {108, 132, 148})
      for (int kind = 0; kind < 2; ++kind) {
        float best = 1e9f;
        int herr = 0;
        for (int rep = 0; rep < 3; ++rep) {
          cudaMemsetAsync(counter, 0, 4, st); cudaMemsetAsync(err, 0, 4, st);
          int ph = phases, sp = spin, kd = kind, touch = smem > 0;
          void* args[] = {&counter, &data, &ph, &sp, &kd, &out, &err, &touch};
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0, st);
          cudaError_t le = cudaLaunchCooperativeKernel((void*)k_persist, dim3(G), dim3(320), args, smem, st);
          cudaEventRecord(e1, st);
          cudaError_t se = cudaStreamSynchronize(st);
          if (le != cudaSuccess || se != cudaSuccess) { printf("launch error %s / %s\n", cudaGetErrorString(le), cudaGetErrorString(se)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
          cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
        }
        printf("persistent smem %6d grid %3d barrier %-9s: %.2f us/phase (spin 3.00) -> boundary %.2f us%s\n", smem, G,
               kind ? "grid.sync" : "red+spin", best * 1e3 / phases, best * 1e3 / phases - 3.0, herr ? "  [TIMEOUT]" : "");
      }
  return 0;
}
