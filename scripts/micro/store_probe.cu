// Micro-probe: what does the output epilogue of a 128-row x 256-column tile cost when 148 CTAs store at once?
// Every CTA writes `tiles` tiles of 128 rows: an fp32 master [rows, 256] plus two 16-bit planes [rows, 256] (what the decoder kernels
// write per activation: 256 KB per tile), either
//   mode 0: with the warp-cooperative pattern of warp_store_act (linear.cuh): per instruction 4 rows x 128 B (fp32) / 4 rows x 64 B
//           (a plane), 8 warps, each owning a 32-row x 32-column chunk at a time;
//   mode 1: with TMA tensor stores out of shared memory: boxes of [128 rows x 128 B] (fp32: 32 columns, plane: 64 columns);
//   mode 2: mode 0 without the fp32 master (planes only);   mode 3: mode 1 without the fp32 master.
// The data is synthetic (registers / an uninitialised smem tile); only the store path is timed (clock64 per CTA + the grid time).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu && ./store_probe
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void tma_store_2d(const void* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(256, 1) k_store(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmP,
                                                  float* f32, uint16_t* pl, long plane_elems, int tiles, int mode, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool master = mode < 2;
  const long long t0 = clock64();
  for (int it = 0; it < tiles; ++it) {
    const long row0 = (static_cast<long>(blockIdx.x) + static_cast<long>(it) * gridDim.x) * 128;
    if ((mode & 1) == 0) {
      // warp (wq, ch): rows 32 wq .., columns 128 ch + 32 c (c = 0..3), like the LayerNorm epilogues
      const int wq = warp & 3, ch = warp >> 2;
      for (int c = 0; c < 4; ++c) {
        const int col0 = ch * 128 + c * 32;
        const long off0 = (row0 + wq * 32 + (lane >> 3)) * 256 + col0 + (lane & 7) * 4;
        const float4 v = make_float4(1.f + lane, 2.f, 3.f, 4.f + it);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (master) *reinterpret_cast<float4*>(f32 + off0 + i * 1024) = v;
          *reinterpret_cast<uint2*>(pl + off0 + i * 1024) = make_uint2(lane, i);
          *reinterpret_cast<uint2*>(pl + plane_elems + off0 + i * 1024) = make_uint2(i, lane);
        }
      }
    } else {
      // the tile is assumed to be in shared memory already (the epilogue writes it there instead of to global): 8 fp32 boxes of
      // [128 x 32 columns] + 2 planes x 4 boxes of [128 x 64 columns], 16 KB each, issued by one thread
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (master)
          for (int b = 0; b < 8; ++b) tma_store_2d(&tmF, smem + (b & 7) * 16384, b * 32, static_cast<int>(row0));
        for (int p = 0; p < 2; ++p)
          for (int b = 0; b < 4; ++b)
            tma_store_2d(&tmP, smem + ((p * 4 + b) & 7) * 16384, b * 64, static_cast<int>(p * (plane_elems / 256) + row0));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be overwritten (the next tile's epilogue)
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    if (mode & 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    out[blockIdx.x] = static_cast<unsigned long long>(clock64() - t0);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
  const long rows = 25088 + 128 * 148;   // enough for 148 CTAs x 2 tiles and the decoder's 196 tiles
  float* f32; uint16_t* pl;
  cudaMalloc(&f32, rows * 256 * 4); cudaMalloc(&pl, 2 * rows * 256 * 2);
  unsigned long long* out; cudaMalloc(&out, 148 * 8);
  CUtensorMap tmF, tmP;
  {
    cuuint64_t dims[2] = {256, static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {256 * 4};
    cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
    CUresult r = encode(&tmF, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, f32, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode f32 failed %d\n", (int)r); return 1; }
    cuuint64_t dimsp[2] = {256, static_cast<cuuint64_t>(2 * rows)};
    cuuint64_t stridesp[1] = {256 * 2};
    cuuint32_t boxp[2] = {64, 128};
    r = encode(&tmP, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl, dimsp, stridesp, boxp, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode planes failed %d\n", (int)r); return 1; }
  }
  cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<unsigned long long> h(148);
  printf("# output epilogue store paths, 148 CTAs (one per SM), 128-row tiles of 256 columns; per tile: fp32 128 KB + two planes 128 KB\n");
  printf("%-34s %6s %10s %12s %12s %10s\n", "mode", "tiles", "grid us", "clk/tile", "B/clk/SM", "chip TB/s");
  const char* names[4] = {"st.global coalesced, master+planes", "TMA store, master+planes", "st.global coalesced, planes only", "TMA store, planes only"};
  for (int mode = 0; mode < 4; ++mode)
    for (int tiles : {1, 2}) {
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_store<<<148, 256, 136 * 1024>>>(tmF, tmP, f32, pl, rows * 256, tiles, mode, out);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      cudaMemcpy(h.data(), out, 148 * 8, cudaMemcpyDeviceToHost);
      double clk = 0;
      for (auto v : h) clk += static_cast<double>(v);
      clk /= 148.0 * tiles;
      const double bytes = (mode < 2 ? 256.0 : 128.0) * 1024;
      printf("%-34s %6d %10.2f %12.0f %12.1f %10.2f\n", names[mode], tiles, best * 1e3, clk, bytes / clk, bytes * tiles * 148 / (best * 1e-3) / 1e12);
    }
  return 0;
}
