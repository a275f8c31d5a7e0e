// Micro-probe: what does a kernel boundary cost on this box, with and without programmatic dependent launch?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void k_chain(unsigned long long* stamps, int idx, int spin_ns, int early, int smem_touch) {
  extern __shared__ char sm[];
  unsigned long long t0 = gtime();
  if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (smem_touch) sm[threadIdx.x] = 1;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long t1 = gtime();
  while (gtime() - t1 < (unsigned long long)spin_ns) {}
  unsigned long long t2 = gtime();
  if (!early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) { stamps[3 * idx] = t0; stamps[3 * idx + 1] = t1; stamps[3 * idx + 2] = t2; }
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 3 * 64 * 8);
  unsigned long long h[3 * 64];
  cudaStream_t st; cudaStreamCreate(&st);
  const int N = 32;
  for (int smem : {0, 100 * 1024, 190 * 1024}) {
    cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {40, 120, 148}) {
      for (int pdl = 0; pdl < 2; ++pdl) {
        for (int graph = 0; graph < 2; ++graph) {
          auto body = [&](cudaStream_t s) {
            for (int i = 0; i < N; ++i) {
              cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = s;
              cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl;
              cfg.attrs = at; cfg.numAttrs = 1;
              cudaLaunchKernelEx(&cfg, k_chain, d, i, 3000, 1, smem > 0 ? 1 : 0);
            }
          };
          float ms = 0; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          if (graph) {
            cudaGraph_t g; cudaGraphExec_t ge;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal); body(st); cudaStreamEndCapture(st, &g);
            cudaGraphInstantiate(&ge, g, 0);
            cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
            cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
          } else {
            body(st); cudaStreamSynchronize(st);
            cudaEventRecord(e0, st); body(st); cudaEventRecord(e1, st); cudaStreamSynchronize(st);
          }
          cudaEventElapsedTime(&ms, e0, e1);
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          // gap between kernel i's end of spin (t2) and kernel i+1's wait-return (t1); and how early i+1 started (t0) relative to i's t2
          double gap = 0, early = 0;
          for (int i = 1; i < N; ++i) { gap += (double)h[3 * i + 1] - (double)h[3 * (i - 1) + 2]; early += (double)h[3 * (i - 1) + 2] - (double)h[3 * i]; }
          printf("smem %6d grid %3d pdl %d graph %d: %.2f us/kernel (spin 3.00) | end->next-wait-return %.2f us | next started %.2f us before prev end | err %s\n",
                 smem, grid, pdl, graph, ms * 1e3 / N, gap / (N - 1) / 1e3, early / (N - 1) / 1e3, cudaGetErrorString(cudaGetLastError()));
        }
      }
    }
  }
  return 0;
}
