// Micro-probe: how many bytes per clock can ONE SM pull from L2 into shared memory, as a function of
//   * the mechanism: 2-D tensor boxes [128 rows x 128 B] out of a row-major matrix with the 128-byte swizzle (what the GEMM kernels
//     use: every box row is its own 128-byte L2 request, rows `ld` bytes apart)  vs  ONE contiguous cp.async.bulk per stage out of
//     a pre-tiled copy of the same data (the stage image already laid out in global memory),
//   * the stage size and the ring depth (bytes in flight),
//   * how many SMs pull at once (1, 4, 27 x 4, 148) and whether they all pull the SAME bytes (weights re-streamed by every token
//     group) or disjoint ones.
// The consumer only waits for `full` and releases the slot, so this is the ceiling of a weight-streaming mainloop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest_probe tma_ingest_probe.cu && ./tma_ingest_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t spins = 0;
  while (!mbar_try(b, par)) if (++spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_2d(void* dst, const void* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: 2-D boxes of [box_rows x 64] 16-bit elements (box_rows * 128 B each), `boxes` of them per stage
// mode 1: one contiguous bulk copy of stage_bytes per stage
// mode 2: contiguous, split into `boxes` bulk copies per stage
struct Args {
  int mode, stages, boxes, box_rows, nstage_total, same;
  int rows_total, kcols;          // matrix geometry (2-D mode): rows x kcols 16-bit elements
  const uint8_t* flat;            // contiguous copy
  long flat_bytes;
  unsigned long long* out;        // per CTA: clocks
};

__global__ void __launch_bounds__(64, 1) k_ingest(const __grid_constant__ CUtensorMap tm, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int box_bytes = a.box_rows * 128, stage_bytes = a.boxes * box_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + a.stages * stage_bytes);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int kb_per_row = a.kcols / 64;                     // k-blocks per matrix row
  const uint32_t ntiles = static_cast<uint32_t>(a.rows_total / 128) * kb_per_row;   // 16 KB tiles of the matrix (power of two)
  const long cta_off = a.same ? 0 : blockIdx.x;            // disjoint: CTAs start at different tiles
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int j = 0; j < a.nstage_total; ++j) {
      const int slot = j % a.stages, use = j / a.stages;
      if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
      mbar_expect_tx(&full[slot], stage_bytes);
      uint8_t* dst = smem + slot * stage_bytes;
      // 32-bit index arithmetic with power-of-two wraps only: a 64-bit modulo per stage costs more than the copy it addresses
      const uint32_t tile = (static_cast<uint32_t>(cta_off) * 7u + static_cast<uint32_t>(j) * a.boxes) & (ntiles - 1u);   // 16 KB units
      if (a.mode == 0) {
        for (int b = 0; b < a.boxes; ++b) {
          const uint32_t t = (tile + b) & (ntiles - 1u);
          tma_2d(dst + b * box_bytes, &tm, &full[slot], (t & (kb_per_row - 1)) * 64, (t / kb_per_row) * 128);
        }
      } else {
        const uint32_t off = (tile * 16384u) & (static_cast<uint32_t>(a.flat_bytes) / 2 - 1u);   // stays inside the first half + one stage
        if (a.mode == 1) bulk_1d(dst, a.flat + off, stage_bytes, &full[slot]);
        else for (int b = 0; b < a.boxes; ++b) bulk_1d(dst + b * box_bytes, a.flat + off + b * box_bytes, box_bytes, &full[slot]);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int j = 0; j < a.nstage_total; ++j) {
      const int slot = j % a.stages, use = j / a.stages;
      mbar_wait(&full[slot], use & 1);
      mbar_arrive(&empty[slot]);
    }
    t1 = clock64();
    a.out[blockIdx.x * 2 + 1] = t1;
  }
  if (threadIdx.x == 0) a.out[blockIdx.x * 2] = t0;
}


// Burst test: ONE CTA issues `n` copies of `bytes` each at once, every copy on its own mbarrier (or all on one), and records when
// each barrier completes: do independent copies overlap their latencies, or does the engine retire them one after the other?
__global__ void __launch_bounds__(64, 1) k_burst(const __grid_constant__ CUtensorMap tm, const uint8_t* flat, int n, int bytes, int mode2d, int onebar,
                                                 long long* stamps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + n * bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    if (onebar) mbar_expect_tx(&full[0], n * bytes);
    for (int i = 0; i < n; ++i) {
      uint64_t* b = onebar ? &full[0] : &full[i];
      if (!onebar) mbar_expect_tx(b, bytes);
      if (mode2d) tma_2d(smem + i * bytes, &tm, b, (i % 16) * 64, (i / 16) * (bytes / 128));
      else bulk_1d(smem + i * bytes, flat + static_cast<long>(i) * bytes, bytes, b);
    }
    stamps[0] = clock64() - t0;   // issue time of the burst
    for (int i = 0; i < (onebar ? 1 : n); ++i) {
      mbar_wait(&full[i], 0);
      stamps[1 + i] = clock64() - t0;
    }
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
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  // "weights": 2048 rows x 1024 16-bit columns = 4 MB (the hi / lo planes of one layer's four FFN matrices); L2 resident
  const int rows = 2048, kcols = 1024;
  const long bytes = static_cast<long>(rows) * kcols * 2;
  uint8_t* w; cudaMalloc(&w, bytes); cudaMemset(w, 1, bytes);
  unsigned long long* out; cudaMalloc(&out, 148 * 16);
  cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  std::vector<unsigned long long> h(148 * 2);

  {
    long long* st; cudaMalloc(&st, 64 * 8);
    cudaFuncSetAttribute(k_burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    CUtensorMap tm;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(kcols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kcols) * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("# burst: one CTA, n copies issued at once; clk from first issue to [burst issued | barrier i complete ...]\n");
    for (int mode2d : {0, 1})
      for (int onebar : {0, 1})
        for (int bytes : {4096, 16384})
          for (int n : {1, 2, 4, 8}) {
            if (mode2d && bytes != 16384) continue;
            long long hs[16] = {0};
            for (int rep = 0; rep < 2; ++rep) k_burst<<<1, 64, n * bytes + 2048>>>(tm, w, n, bytes, mode2d, onebar, st);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("burst failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(hs, st, 16 * 8, cudaMemcpyDeviceToHost);
            printf("%-6s %-8s %6d B x %d :", mode2d ? "box2d" : "bulk", onebar ? "one-bar" : "own-bar", bytes, n);
            for (int i = 0; i < 1 + (onebar ? 1 : n); ++i) printf(" %lld", hs[i]);
            printf("\n");
          }
  }
  printf("# per-SM TMA ingest from L2 (4 MB source, warm); B/clk per SM = bytes per CTA / (last full observed - first issue)\n");
  printf("%-8s %6s %7s %6s %9s %6s %5s | %9s %9s %10s\n", "mode", "boxKB", "stageKB", "depth", "flightKB", "ctas", "same", "B/clk/SM", "min", "chip KB/clk");
  for (int box_rows : {128}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(kcols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kcols) * 2};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int mode : {0, 1, 2}) {
      if (box_rows == 256 && mode != 0) continue;
      for (int boxes : {1, 2, 4, 6}) {
        for (int stages : {1, 2, 3, 4, 6, 8}) {
          const int stage_bytes = boxes * box_rows * 128;
          if (stages * stage_bytes > 200 * 1024) continue;
          for (int ctas : {1, 108, 148}) {
            for (int same : {1, 0}) {
              if (ctas == 1 && same == 0) continue;
              if ((ctas == 4) && same == 0) continue;
              Args a;
              a.mode = mode; a.stages = stages; a.boxes = boxes; a.box_rows = box_rows; a.same = same;
              a.nstage_total = (1 << 20) / stage_bytes;   // 1 MB per CTA, like k_ffn_swap
              a.rows_total = rows; a.kcols = kcols; a.flat = w; a.flat_bytes = bytes; a.out = out;
              const int smem = stages * stage_bytes + 1024 + 512;
              for (int rep = 0; rep < 3; ++rep) k_ingest<<<ctas, 64, smem>>>(tm, a);
              if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
              cudaMemcpy(h.data(), out, ctas * 16, cudaMemcpyDeviceToHost);
              double sum = 0, mn = 1e30;
              const double per_cta = static_cast<double>(a.nstage_total) * stage_bytes;
              for (int c = 0; c < ctas; ++c) {
                const double bpc = per_cta / static_cast<double>(h[2 * c + 1] - h[2 * c]);
                sum += bpc; if (bpc < mn) mn = bpc;
              }
              printf("%-8s %6d %7d %6d %9d %6d %5d | %9.1f %9.1f %10.2f\n", mode == 0 ? "box2d" : (mode == 1 ? "bulk1" : "bulkN"),
                     box_rows * 128 / 1024, stage_bytes / 1024, stages, stages * stage_bytes / 1024, ctas, same, sum / ctas, mn, sum / 1024);
            }
          }
        }
      }
    }
  }
  return 0;
}
