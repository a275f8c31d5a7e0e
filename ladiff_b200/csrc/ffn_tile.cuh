// Fused feed-forward block on 128-row tiles for the throughput-bound launches (decoder / encoder: 25 088 frame rows; denoiser at
// >= 2560 latent rows), where the two separate fused linears are bound by what they move, not by the tensor pipe: the 1024-wide
// hidden activation costs 100 - 150 MB of HBM writes and the same again in reads per layer, and every 128 x 256 output tile
// re-streams its operands from L2 (profiles/r02_gemm_stamps.txt).  Here the hidden activation never leaves the SM:
//
//   per tile of 128 rows            X [128 x 256] resident in shared memory (16-bit operand planes, loaded once by TMA)
//   for each hidden chunk j of 128: acc_h  = X . W1[j]^T                    tcgen05.mma, A / B from shared memory  (TMEM cols [0, 128))
//                                   h      = act(acc_h + b1[j])             epilogue warps: TMEM -> registers -> packed 16-bit pairs
//                                                                           back into TENSOR MEMORY (hi | lo planes, cols [128, 256))
//                                   acc_o += h . W2[:, j]^T                 tcgen05.mma with the A operand read from TMEM (cols [256, 512))
//   out = LayerNorm-kind epilogue(acc_o + b2 [+ residual])                  one TMEM read, row statistics in registers
//
// so a tile ingests X once and every weight byte once (2.1 MB in the x3 mode, about the time its 3 x 403 MFLOP take on the tensor
// pipe) instead of 2 x 384 KB + 1.5 MB per 128 x 256 output tile plus the round trip of the hidden through HBM.  The weight blocks
// ([128 rows x 64 k] per plane) stream through a 3-stage TMA ring in exactly the order the MMA warp consumes them:
// G1(0), then G1(j + 1) before G2(j) -- the score accumulator is drained early by the epilogue warps (TMEM -> registers), so the next
// chunk's first GEMM runs under the activation math of the current one.  CTAs are persistent over tiles (grid = min(tiles, #SMs)).
// NSPLIT = 2: fp16 hi / lo operands, 3 products (x3 mode); NSPLIT = 1: bf16.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "linear.cuh"
#include "attn_tc5.cuh"
#include "tc_ptx.cuh"

struct FtArgs {
  int M_max;
  const int* M_dev;
  const float *b1, *b2;          // [1024], [256]
  int act;                       // EPI_RELU / EPI_GELU
  int kind;                      // EPI_LN: LN(acc + b2 + res) g + b [+ addv[add_idx[row]]];  EPI_LN_MOD_SILU: SiLU(LN(acc + b2) g + b) (1 + scale) + shift)
  const float *ln_g, *ln_b;
  const float* res;              // fp32 rows, ld 256 (EPI_LN; may be null)
  const float* addv;             // optional broadcast add after the LayerNorm
  const int* add_idx;
  int ld_add;
  const float* mod;              // [scale(256) | shift(256)]
  Act out;
  int out_planes;
  int x_plane_rows, w1_plane_rows, w2_plane_rows;
  long long* dbg;                // optional clock64 stamps of CTA 0, first tile (profiling: LADIFF_FT_DBG=1)
};

template <int NSPLIT>
struct FtCfg {
  static constexpr int BM = 128, D = 256, FF = 1024, HC = 128, BK = 64, NCH = FF / HC;
  static constexpr int U = 128 * BK * 2;                    // one [128 x 64] 16-bit block: 16 KB
  static constexpr int X_BYTES = NSPLIT * (D / BK) * U;     // resident X tile: [k-block][plane]
  static constexpr int STG = NSPLIT * U;                    // ring stage: one weight block, all planes
  static constexpr int NSTG = NSPLIT == 2 ? 3 : 6;
  static constexpr int RED_BYTES = 2 * 128 * 4;             // row statistics exchanged between the two column halves
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = X_BYTES + NSTG * STG + RED_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
  static constexpr int EPI_WARPS = 8, EPI_THREADS = 256, THREADS = 64 + EPI_THREADS;
  static constexpr int STAGES_PER_TILE = NCH * 8;           // per chunk: 4 blocks of W1, 2 k-blocks x 2 output halves of W2
  // tensor memory columns
  static constexpr int ACC_H = 0, H_HI = 128, H_LO = 192, ACC_O = 256, TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int NSPLIT>
__global__ void __launch_bounds__(FtCfg<NSPLIT>::THREADS, 1)
k_ffn_tile(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
           const FtArgs p) {
  using C = FtCfg<NSPLIT>;
  constexpr int U = C::U, STG = C::STG, NSTG = C::NSTG;
  const int M = p.M_dev ? min(p.M_max, *p.M_dev) : p.M_max;
  const int ntiles = (M + C::BM - 1) / C::BM;
  tc::pdl_launch_dependents();
  if (static_cast<int>(blockIdx.x) >= ntiles) return;
  const int my_tiles = (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  extern __shared__ uint8_t ft_raw[];
  uint8_t* smem = ft_raw + ((1024u - (tc::smem_u32(ft_raw) & 1023u)) & 1023u);
  uint8_t* xs = smem;                                   // [kb][plane][128 x 128 B]
  uint8_t* ring = xs + C::X_BYTES;
  float* red = reinterpret_cast<float*>(ring + NSTG * STG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + C::RED_BYTES);
  uint64_t* full = bars;                  // [NSTG]  TMA -> MMA
  uint64_t* empty = bars + NSTG;          // [NSTG]  MMA -> TMA
  uint64_t* x_full = bars + 2 * NSTG;     // X tile landed
  uint64_t* x_empty = x_full + 1;         // last G1 of the tile retired: X may be overwritten
  uint64_t* acch_full = x_full + 2;       // G1(j) retired
  uint64_t* acch_empty = x_full + 3;      // 256 arrivals: acc_h is in registers
  uint64_t* h_full = x_full + 4;          // 256 arrivals: h(j) is in tensor memory
  uint64_t* h_empty = x_full + 5;         // G2(j) retired: h may be overwritten
  uint64_t* acco_full = x_full + 6;       // last G2 of the tile retired
  uint64_t* acco_empty = x_full + 7;      // 256 arrivals: acc_o is in registers
  uint64_t* x_free = x_full + 8;          // 256 arrivals: the output epilogue no longer uses the X region as its transposition scratch
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 9);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  long long* dbg = (p.dbg && blockIdx.x == 0) ? p.dbg : nullptr;
#define FTS(i) do { if (dbg && lane == 0) dbg[i] = clock64(); } while (0)
  if (threadIdx.x == 0) FTS(0);

  // ring stage s of a tile (0 .. 63): which weight block.  Order = MMA issue order: G1(0); then for j: G1(j + 1), G2(j); G2(7).
  //   returns kind (0 = W1 block of chunk j, k-block kb; 1 = W2 block of chunk j, k-block kb (of 2), output half nh)
  auto decode = [](int s, int& kind, int& j, int& kb, int& nh) {
    if (s < 4) { kind = 0; j = 0; kb = s; nh = 0; return; }
    const int t = s - 4, grp = t >> 3, r = t & 7;       // groups of 8: G1(grp + 1) [4], G2(grp) [4]   (grp 0 .. 6), then G2(7)
    if (grp < C::NCH - 1) {
      if (r < 4) { kind = 0; j = grp + 1; kb = r; nh = 0; }
      else { kind = 1; j = grp; kb = (r - 4) >> 1; nh = (r - 4) & 1; }
    } else {
      kind = 1; j = C::NCH - 1; kb = r >> 1; nh = r & 1;
    }
  };
  auto issue_stage = [&](long gs) {       // gs: global stage counter of this CTA
    const int slot = static_cast<int>(gs % NSTG);
    int kind, j, kb, nh;
    decode(static_cast<int>(gs % C::STAGES_PER_TILE), kind, j, kb, nh);
    tc::mbar_expect_tx(&full[slot], STG);
#pragma unroll
    for (int pl = 0; pl < NSPLIT; ++pl) {
      if (kind == 0)
        tc::tma_load_2d(ring + slot * STG + pl * U, &tmW1, &full[slot], kb * C::BK, w_plane<NSPLIT>(pl) * p.w1_plane_rows + j * C::HC);
      else
        tc::tma_load_2d(ring + slot * STG + pl * U, &tmW2, &full[slot], j * C::HC + kb * C::BK, w_plane<NSPLIT>(pl) * p.w2_plane_rows + nh * 128);
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmX);
      tc::tma_prefetch_desc(&tmW1);
      tc::tma_prefetch_desc(&tmW2);
      for (int i = 0; i < NSTG; ++i) {
        tc::mbar_init(&full[i], 1);
        tc::mbar_init(&empty[i], 1);
      }
      tc::mbar_init(x_full, 1);
      tc::mbar_init(x_empty, 1);
      tc::mbar_init(acch_full, 1);
      tc::mbar_init(acch_empty, C::EPI_THREADS);
      tc::mbar_init(h_full, C::EPI_THREADS);
      tc::mbar_init(h_empty, 1);
      tc::mbar_init(acco_full, 1);
      tc::mbar_init(acco_empty, C::EPI_THREADS);
      tc::mbar_init(x_free, C::EPI_THREADS);
      tc::fence_barrier_init();
      tc::fence_proxy_async();
    }
    __syncwarp();
    if (tc::elect_one())
      for (int s = 0; s < NSTG; ++s) issue_stage(s);       // weights never depend on the previous grid
    __syncwarp();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long total_stages = static_cast<long>(my_tiles) * C::STAGES_PER_TILE;

  if (warp == 0) {
    // ===== TMA producer =====
    tc::pdl_wait();
    auto load_x = [&](int it) {
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      if (tc::elect_one()) {
        tc::mbar_expect_tx(x_full, C::X_BYTES);
        for (int kb = 0; kb < C::D / C::BK; ++kb)
          for (int pl = 0; pl < NSPLIT; ++pl)
            tc::tma_load_2d(xs + (kb * NSPLIT + pl) * U, &tmX, x_full, kb * C::BK, pl * p.x_plane_rows + tile * C::BM);
      }
      __syncwarp();
    };
    load_x(0);
    for (long gs = NSTG; gs < total_stages; ++gs) {
      const long prev = gs - NSTG;
      tc::mbar_wait(&empty[prev % NSTG], (prev / NSTG) & 1);
      if (gs < 64) FTS(64 + gs);
      if (tc::elect_one()) issue_stage(gs);
      __syncwarp();
      // the next tile's X goes in once this tile's output epilogue has released the X region (it transposes its rows through it
      // for coalesced global access), i.e. after the tile's last weight stage has been issued
      const int it = static_cast<int>(gs / C::STAGES_PER_TILE);
      if (gs % C::STAGES_PER_TILE == C::STAGES_PER_TILE - 1 && it + 1 < my_tiles) {
        tc::mbar_wait(x_empty, it & 1);
        tc::mbar_wait(x_free, it & 1);
        load_x(it + 1);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = tc::idesc_op<NSPLIT>(128, 128);
    const uint32_t xs_u = tc::smem_u32(xs), ring_u = tc::smem_u32(ring);
    long gs = 0;
    int cnt_h = 0;      // chunks whose h has been consumed so far (parity source of h_full / acch_empty)
    auto g1 = [&](int j) {
      for (int kb = 0; kb < 4; ++kb, ++gs) {
        const int slot = static_cast<int>(gs % NSTG);
        tc::mbar_wait(&full[slot], (gs / NSTG) & 1);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t sw = ring_u + slot * STG, sx = xs_u + kb * NSPLIT * U;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t x_hi = tc::smem_desc_sw128(sx + kk * 32), w_hi = tc::smem_desc_sw128(sw + kk * 32);
            const uint32_t acc = (kb | kk) != 0;
            if (NSPLIT == 1) {
              tc::mma_bf16_ss(tmem + C::ACC_H, x_hi, w_hi, idesc, acc);
            } else {
              const uint64_t x_lo = tc::smem_desc_sw128(sx + U + kk * 32), w_lo = tc::smem_desc_sw128(sw + U + kk * 32);
              tc::mma_bf16_ss(tmem + C::ACC_H, x_lo, w_hi, idesc, acc);
              tc::mma_bf16_ss(tmem + C::ACC_H, x_hi, w_lo, idesc, 1u);
              tc::mma_bf16_ss(tmem + C::ACC_H, x_hi, w_hi, idesc, 1u);
            }
          }
          tc::mma_commit(&empty[slot]);
          if (kb == 3) {
            tc::mma_commit(acch_full);
            if (j == C::NCH - 1) tc::mma_commit(x_empty);    // last reader of the X tile
          }
        }
        __syncwarp();
      }
    };
    auto g2 = [&](int j) {
      for (int r = 0; r < 4; ++r, ++gs) {
        const int kb = r >> 1, nh = r & 1;
        const int slot = static_cast<int>(gs % NSTG);
        tc::mbar_wait(&full[slot], (gs / NSTG) & 1);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t sw = ring_u + slot * STG;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t w_hi = tc::smem_desc_sw128(sw + kk * 32);
            const uint32_t h_hi = tmem + C::H_HI + kb * 32 + kk * 8;
            const uint32_t d = tmem + C::ACC_O + nh * 128;
            const uint32_t acc = (j | kb | kk) != 0;
            if (NSPLIT == 1) {
              at5::mma_ts(d, h_hi, w_hi, idesc, acc);
            } else {
              const uint64_t w_lo = tc::smem_desc_sw128(sw + U + kk * 32);
              const uint32_t h_lo = tmem + C::H_LO + kb * 32 + kk * 8;
              at5::mma_ts(d, h_lo, w_hi, idesc, acc);
              at5::mma_ts(d, h_hi, w_lo, idesc, 1u);
              at5::mma_ts(d, h_hi, w_hi, idesc, 1u);
            }
          }
          tc::mma_commit(&empty[slot]);
          if (r == 3) {
            tc::mma_commit(h_empty);
            if (j == C::NCH - 1) tc::mma_commit(acco_full);  // the tile's output accumulator is complete
          }
        }
        __syncwarp();
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      tc::mbar_wait(x_full, it & 1);
      tc::tc_fence_after();
      if (it == 0) FTS(1);
      if (it == 1) FTS(5);
      // acc_h of the previous tile's last chunk has been drained (its h_full was waited for below)
      g1(0);
      for (int j = 0; j < C::NCH; ++j) {
        if (j + 1 < C::NCH) {
          tc::mbar_wait(acch_empty, (cnt_h + j) & 1);      // epilogue holds acc_h(j) in registers
          tc::tc_fence_after();
          if (it == 0) FTS(8 + 4 * j);
          g1(j + 1);
          if (it == 0) FTS(9 + 4 * j);
        }
        tc::mbar_wait(h_full, (cnt_h + j) & 1);            // h(j) stored (and acc_h(j) drained)
        tc::tc_fence_after();
        if (j == 0 && it > 0) {                            // acc_o of the previous tile is in registers
          tc::mbar_wait(acco_empty, (it - 1) & 1);
          tc::tc_fence_after();
        }
        if (it == 0) FTS(10 + 4 * j);
        g2(j);
        if (it == 0) FTS(11 + 4 * j);
      }
      cnt_h += C::NCH;
    }
  } else {
    // ===== epilogue warps: thread <-> tile row (TMEM lane), two threads per row split the columns =====
    const int ew = warp - 2, wq = warp & 3, ch = ew >> 2;
    const int r = wq * 32 + lane;
    const uint32_t tlane = static_cast<uint32_t>(wq * 32) << 16;
    tc::pdl_wait();
    int cnt = 0;      // chunk counter (parity of acch_full / h_empty)
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const long row = static_cast<long>(tile) * C::BM + r;
      const bool valid = row < M;
      for (int j = 0; j < C::NCH; ++j, ++cnt) {
        // ---- h(j) = act(acc_h + b1): this thread's 64 hidden columns [ch * 64, +64) of the chunk
        tc::mbar_wait(acch_full, cnt & 1);
        tc::tc_fence_after();
        if (it == 0 && ew == 0) FTS(128 + 4 * j);
        float va[32], vb[32];
        tc::tmem_ld32(tmem + tlane + C::ACC_H + ch * 64, va);
        tc::tmem_ld32(tmem + tlane + C::ACC_H + ch * 64 + 32, vb);
        tc::tc_fence_before();
        tc::mbar_arrive(acch_empty);
        const float4* bp = reinterpret_cast<const float4*>(p.b1 + j * C::HC + ch * 64);
        uint32_t ha[16], la[16], hb[16], lb[16];
        const bool relu = p.act == EPI_RELU;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(bp + q), c4 = __ldg(bp + 8 + q);
          float a0 = va[4 * q] + b4.x, a1 = va[4 * q + 1] + b4.y, a2 = va[4 * q + 2] + b4.z, a3 = va[4 * q + 3] + b4.w;
          float e0 = vb[4 * q] + c4.x, e1 = vb[4 * q + 1] + c4.y, e2 = vb[4 * q + 2] + c4.z, e3 = vb[4 * q + 3] + c4.w;
          if (relu) {
            a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f);
            e0 = fmaxf(e0, 0.f); e1 = fmaxf(e1, 0.f); e2 = fmaxf(e2, 0.f); e3 = fmaxf(e3, 0.f);
          } else {
            a0 = gelu_erf_fast(a0); a1 = gelu_erf_fast(a1); a2 = gelu_erf_fast(a2); a3 = gelu_erf_fast(a3);
            e0 = gelu_erf_fast(e0); e1 = gelu_erf_fast(e1); e2 = gelu_erf_fast(e2); e3 = gelu_erf_fast(e3);
          }
          split2_op<NSPLIT>(a0, a1, ha[2 * q], la[2 * q]);
          split2_op<NSPLIT>(a2, a3, ha[2 * q + 1], la[2 * q + 1]);
          split2_op<NSPLIT>(e0, e1, hb[2 * q], lb[2 * q]);
          split2_op<NSPLIT>(e2, e3, hb[2 * q + 1], lb[2 * q + 1]);
        }
        if (it == 0 && ew == 0) FTS(129 + 4 * j);
        if (cnt > 0) {                                       // G2 of the previous chunk has finished reading h
          tc::mbar_wait(h_empty, (cnt - 1) & 1);
          tc::tc_fence_after();
        }
        if (it == 0 && ew == 0) FTS(130 + 4 * j);
        at5::tmem_st16(tmem + tlane + C::H_HI + ch * 32, ha);
        at5::tmem_st16(tmem + tlane + C::H_HI + ch * 32 + 16, hb);
        if (NSPLIT == 2) {
          at5::tmem_st16(tmem + tlane + C::H_LO + ch * 32, la);
          at5::tmem_st16(tmem + tlane + C::H_LO + ch * 32 + 16, lb);
        }
        at5::tmem_wait_st();
        tc::tc_fence_before();
        tc::mbar_arrive(h_full);
        if (it == 0 && ew == 0) FTS(131 + 4 * j);
      }
      // ---- output rows: acc_o + b2 (+ residual) -> LayerNorm kind -> fp32 master / operand planes.  This thread: columns
      // [ch * 128, +128) of its row, all in registers (one TMEM read), statistics exchanged with the partner warp.
      tc::mbar_wait(acco_full, it & 1);
      tc::tc_fence_after();
      if (it == 0 && ew == 0) FTS(2);
      if (it == 1 && ew == 0) FTS(6);
      // Three passes over the accumulator in tensor memory (sum -> centred sum of squares -> normalise + store), 32 columns at a
      // time, so that the epilogue stays within the register budget of 320 threads (128 live accumulator values per thread spill,
      // and with the whole shared-memory carve-out taken there is no L1 to catch local memory).  Global rows are read / written
      // through a warp-private transposition tile in the (now dead) X region: every instruction touches whole 128-byte lines.
      const int col0 = ch * 128;
      const uint32_t tacc = tmem + tlane + C::ACC_O + col0;
      const bool ln_res = p.kind == EPI_LN && p.res != nullptr;
      float (*tt)[36] = reinterpret_cast<float (*)[36]>(xs + ew * (32 * 36 * 4));
      const long row_base = static_cast<long>(blockIdx.x + it * gridDim.x) * C::BM + wq * 32;
      const int nvalid = static_cast<int>(min(32L, static_cast<long>(M) - row_base));
      float s = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32], t[32];
        if (ln_res) warp_load_rows(tt, p.res, C::D, nullptr, row_base, nvalid, col0 + c * 32, lane, t);
        tc::tmem_ld32(tacc + c * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + c * 32) + q);
          v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
        }
        if (ln_res) {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] += t[k];
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) s += v[k];
        tc::tmem_st32(tacc + c * 32, v);
      }
      if (it == 0 && ew == 0) FTS(201);
      red[ch * 128 + r] = s;
      asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_THREADS) : "memory");
      const float mean = (red[r] + red[128 + r]) * (1.f / 256.f);
      asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_THREADS) : "memory");
      float q2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tc::tmem_ld32(tacc + c * 32, v);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float d = v[k] - mean;
          q2 += d * d;
        }
      }
      red[ch * 128 + r] = q2;
      asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_THREADS) : "memory");
      const float rstd = 1.0f / sqrtf((red[r] + red[128 + r]) * (1.f / 256.f) + LD_EPS);
      asm volatile("bar.sync 1, %0;" ::"n"(C::EPI_THREADS) : "memory");
      if (it == 0 && ew == 0) FTS(203);
      const bool add = p.kind == EPI_LN && p.addv != nullptr;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32], t[32];
        if (add) warp_load_rows(tt, p.addv, p.ld_add, p.add_idx, row_base, nvalid, col0 + c * 32, lane, t);
        tc::tmem_ld32(tacc + c * 32, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col0 + c * 32) + q), b4 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col0 + c * 32) + q);
          float z0 = (v[4 * q] - mean) * rstd * g4.x + b4.x, z1 = (v[4 * q + 1] - mean) * rstd * g4.y + b4.y;
          float z2 = (v[4 * q + 2] - mean) * rstd * g4.z + b4.z, z3 = (v[4 * q + 3] - mean) * rstd * g4.w + b4.w;
          if (p.kind == EPI_LN_MOD_SILU) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.mod + col0 + c * 32) + q), h4 = __ldg(reinterpret_cast<const float4*>(p.mod + C::D + col0 + c * 32) + q);
            z0 = silu(z0 * (1.f + s4.x) + h4.x); z1 = silu(z1 * (1.f + s4.y) + h4.y);
            z2 = silu(z2 * (1.f + s4.z) + h4.z); z3 = silu(z3 * (1.f + s4.w) + h4.w);
          } else if (add) {
            z0 += t[4 * q]; z1 += t[4 * q + 1]; z2 += t[4 * q + 2]; z3 += t[4 * q + 3];
          }
          v[4 * q] = z0; v[4 * q + 1] = z1; v[4 * q + 2] = z2; v[4 * q + 3] = z3;
        }
        warp_store_act(tt, p.out, p.out_planes, row_base, nvalid, col0 + c * 32, lane, v);
      }
      tc::tc_fence_before();
      tc::mbar_arrive(acco_empty);      // the next tile's second GEMMs may overwrite the accumulator
      if (ew == 0) FTS(it == 0 ? 4 : 7);
      tc::fence_proxy_async();          // generic-proxy use of the X region before the next tile's TMA (async proxy) writes
      tc::mbar_arrive(x_free);
    }
  }
  if (threadIdx.x == 64) FTS(3);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem, C::TMEM_COLS);
  }
#undef FTS
}
