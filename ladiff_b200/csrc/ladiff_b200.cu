// C-ABI implementation (include/ladiff_b200.h): handle, weight packing, plans (workspace + CUDA graph) and the
// orchestration of the hoisted, ragged sampling path described in DESIGN.md.
#include "../../include/ladiff_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "linear.cuh"
#include "ffn_swap.cuh"
#include "attn_tc5.cuh"
#include "ffn_tile.cuh"

namespace {

constexpr int NL = 9;  // layers (4 input + middle + 4 output blocks)

// ------------------------------------------------------------------------------------------------
struct Err {
  std::string msg;
  int set(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    msg = buf;
    return code;
  }
};
thread_local std::string g_create_error;
bool g_use_pdl = true;
bool g_skip_pdl_once = false;
long long* g_ffn_dbg = nullptr;  // ladiff_ffn_test with LADIFF_DBG_STAMPS=1  // next launch_pdl() uses a full dependency (kernel right after a cross-stream join)

#define CK(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return h->err.set(LADIFF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)
#define CKS(expr)                \
  do {                           \
    int _s = (expr);             \
    if (_s != LADIFF_OK) return _s; \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Raw {
  float* dev = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// One packed nn.Linear: fp32 transposed copy for the SIMT backend, zero-padded bf16 hi/lo planes + TMA map for tcgen05.
struct Weight {
  int N = 0, K = 0, n_pad = 0;
  float* Wt = nullptr;
  op16* pl = nullptr;
  float* bias = nullptr;
  op16* swap_img = nullptr;           // k_ffn_swap streaming image (k_pack_swap_image): x3 stages, then bf16 stages; FFN shapes only
  CUtensorMap map64, map128, map256;  // TMA boxes of 64 / 128 / 256 weight rows (one load per plane and k-block)
};

struct ActBuf {
  Act act{nullptr, nullptr, 0, 0};
  CUtensorMap map;    // TMA box of 128 rows x 64 columns
  CUtensorMap map16, map32, map48;  // boxes of 16 / 32 / 48 rows (k_ffn_swap token groups)
  bool has_map = false;
};

struct Arena {
  std::vector<void*> blocks;
  ~Arena() { clear(); }
  void clear() {
    for (void* p : blocks) cudaFree(p);
    blocks.clear();
  }
  cudaError_t alloc(void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e == cudaSuccess) blocks.push_back(*p);
    return e;
  }
};

struct DenLayerW {
  Weight qkv, ff1, ff2, ca_value, ca_out, gff1, gff2, ffn_out;
  float* inx = nullptr;       // extended in-projection [1536,256]: q | k | 4 x (W_o[:,head] W_v[head])  (k_attn_ln)
  float* inx_bias = nullptr;  // [1536]
  float* out_bias = nullptr;  // out_proj.bias
  float *n1g, *n1b, *n2g, *n2b, *ca_tn_g, *ca_tn_b, *ca_sn_g, *ca_sn_b, *ffn_sn_g, *ffn_sn_b;
};
struct DecLayerW {
  Weight qkv, out, ff1, ff2;
  float *n1g, *n1b, *n2g, *n2b, *n3g, *n3b;
  float* out2_bias = nullptr;   // multihead_attn.out_proj.bias (the projections themselves are folded into memx_all)
};

struct EncLayerW {
  Weight qkv, out, ff1, ff2;
  float *n1g, *n1b, *n2g, *n2b;
};

struct DenoisePlan;
struct ReversePlan;
struct DecodePlan;
struct EncodePlan;
constexpr int MAX_CHAINS = 8;

}  // namespace

struct ladiff_handle {
  ladiff_config cfg;
  Err err;
  EncodeTiledFn encode = nullptr;
  int device = 0;
  int64_t launches = 0;
  std::map<std::string, Raw> raw;
  Arena warena_den, warena_dec;  // packed weights
  Arena* warena = &warena_den;   // arena the pack_* helpers currently fill
  bool den_ready = false, dec_ready = false;
  // denoiser
  DenLayerW den[NL];
  Weight den_skip[4], time1, time2, embproj, timekv_all, textkv_all, mod_all;
  // folded transitions: qkv of layer l+1 straight from (x3, s[, skip]) of layer l, and the skip merge
  // straight from (x3, s, skip) -- the residual GEMM and the skip Linear leave the critical path
  Weight den_qkv_fold[NL];   // [l] for l = 1..8: N = 1792 (q | k | 4 x v' | X), K = 512 (l <= 4) or 768 (l >= 5)
  float *den_fg = nullptr, *den_fb = nullptr, *den_pe = nullptr;
  // decoder
  DecLayerW dec[NL];
  Weight dec_skip[4], dec_final, memx_all;   // memx_all: folded cross-attention table projection, [9 x CX_LD, 256] (k_cross_ln)
  float *dec_fg = nullptr, *dec_fb = nullptr, *dec_pe = nullptr;
  // LA-VAE encoder (packed with the decoder when its keys are present)
  EncLayerW enc[NL];
  Weight enc_skip[4], skel;
  float *enc_fg = nullptr, *enc_fb = nullptr, *enc_pe = nullptr, *enc_gmt = nullptr;
  int skel_kp = 0;
  bool enc_ready = false;
  std::map<std::string, std::unique_ptr<EncodePlan>> enc_plans;
  std::map<std::string, std::unique_ptr<DenoisePlan>> den_plans;
  std::map<std::string, std::unique_ptr<ReversePlan>> rev_plans;
  std::map<std::string, std::unique_ptr<DecodePlan>> dec_plans;
  uint64_t use_clock = 0;   // LRU stamp of the plan caches
  cudaStream_t cap_stream = nullptr;
  // LADIFF_TRACE=1: per-launch %globaltimer records of the fused linears (ladiff_trace_read)
  unsigned long long* trace = nullptr;
  int trace_n = 0, trace_cap = 0;
  std::vector<std::string> trace_names;
  cudaStream_t side[MAX_CHAINS] = {};       // forked streams of the chained reverse loop
  cudaStream_t aux[MAX_CHAINS] = {};        // per-chain side branch (off-critical-path residual / skip GEMMs)
  cudaEvent_t ev_aux_fork[MAX_CHAINS] = {}, ev_aux_join[MAX_CHAINS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHAINS] = {};
};

namespace {

typedef ladiff_handle H;

const char* block_name(int l) {
  static const char* n[NL] = {"input_blocks.0", "input_blocks.1", "input_blocks.2", "input_blocks.3", "middle_block",
                              "output_blocks.0", "output_blocks.1", "output_blocks.2", "output_blocks.3"};
  return n[l];
}

int make_map(H* h, CUtensorMap* m, const op16* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(op16)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<op16*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return h->err.set(LADIFF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r,
                                          (unsigned long long)rows, (unsigned long long)cols);
  return LADIFF_OK;
}

int roundup(int x, int m) { return (x + m - 1) / m * m; }

int alloc_act(H* h, Arena& ar, ActBuf* b, int rows, int ld, bool f32, bool planes) {
  b->act.ld = ld;
  b->act.rows_alloc = roundup(rows < 1 ? 1 : rows, 128);
  b->act.f32 = nullptr;
  b->act.pl = nullptr;
  b->has_map = false;
  const size_t n = static_cast<size_t>(b->act.rows_alloc) * ld;
  if (f32) {
    CK(ar.alloc(reinterpret_cast<void**>(&b->act.f32), n * sizeof(float)));
    CK(cudaMemset(b->act.f32, 0, n * sizeof(float)));
  }
  if (planes) {
    CK(ar.alloc(reinterpret_cast<void**>(&b->act.pl), 2 * n * sizeof(op16)));
    CK(cudaMemset(b->act.pl, 0, 2 * n * sizeof(op16)));
    CKS(make_map(h, &b->map, b->act.pl, 2ull * b->act.rows_alloc, ld, ld, 128));
    CKS(make_map(h, &b->map16, b->act.pl, 2ull * b->act.rows_alloc, ld, ld, 16));
    CKS(make_map(h, &b->map32, b->act.pl, 2ull * b->act.rows_alloc, ld, ld, 32));
    CKS(make_map(h, &b->map48, b->act.pl, 2ull * b->act.rows_alloc, ld, ld, 48));
    b->has_map = true;
  }
  return LADIFF_OK;
}

// Launch with the programmatic-stream-serialization attribute: the kernel may start while its predecessor drains and
// synchronises itself with griddepcontrol.wait (every plan kernel begins with pdl_prologue / pdl_wait).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = (g_use_pdl && !g_skip_pdl_once) ? 1 : 0;
  g_skip_pdl_once = false;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------------------------------------------------------------
// launching the fused linear
struct LinCall {
  const ActBuf* A = nullptr;
  const ActBuf* A2 = nullptr;  // second K source (skip merge / folded branches)
  const ActBuf* A3 = nullptr;  // third K source
  bool no_pdl = false;         // launch with a full dependency (first kernel after a cross-stream join)
  const Weight* W = nullptr;
  int M_max = 0;
  const int* M_dev = nullptr;
  int epi = EPI_BIAS;
  const float* res = nullptr;
  int ldres = 256;
  const float *ln_g = nullptr, *ln_b = nullptr, *mod = nullptr, *addv = nullptr;
  const int* add_idx = nullptr;
  int ld_add = 256;
  const int* row_map = nullptr;
  Act out{nullptr, nullptr, 0, 0};
  int out_planes = 0;
  int n_store = -1;
  long long* dbg = nullptr;
};

template <int BN, int NS, int EPI>
int launch_tc_e(H* h, cudaStream_t st, const LinCall& c, const LinArgs& a, int tiles_m) {
  using Cfg = TcCfg<BN, NS>;
  dim3 grid(tiles_m, (c.W->N + BN - 1) / BN);
  CK(launch_pdl(k_linear_tc<BN, NS, EPI>, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, c.A->map, c.A2 ? c.A2->map : c.A->map,
                c.A3 ? c.A3->map : c.A->map, BN == 64 ? c.W->map64 : (BN == 128 ? c.W->map128 : c.W->map256), a));
  h->launches++;
  return LADIFF_OK;
}

constexpr int LN_CL = 4;  // CTAs per cluster (along N) for the LayerNorm-epilogue linears

template <int NS, int EPI>
int launch_tc_ln(H* h, cudaStream_t st, const LinCall& c, const LinArgs& a, int tiles_m) {
  using Cfg = TcCfg<256 / LN_CL, NS>;
  dim3 grid(tiles_m, LN_CL);
  CK(launch_pdl(k_linear_tc_ln<LN_CL, NS, EPI>, grid, dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, c.A->map, c.A2 ? c.A2->map : c.A->map,
                c.A3 ? c.A3->map : c.A->map, LN_CL == 4 ? c.W->map64 : c.W->map128, a));
  h->launches++;
  return LADIFF_OK;
}

template <int BN, int NS>
int launch_tc(H* h, cudaStream_t st, const LinCall& c, const LinArgs& a, int tiles_m) {
  switch (c.epi) {
    case EPI_BIAS: return launch_tc_e<BN, NS, EPI_BIAS>(h, st, c, a, tiles_m);
    case EPI_RELU: return launch_tc_e<BN, NS, EPI_RELU>(h, st, c, a, tiles_m);
    case EPI_GELU: return launch_tc_e<BN, NS, EPI_GELU>(h, st, c, a, tiles_m);
    case EPI_RES: return launch_tc_e<BN, NS, EPI_RES>(h, st, c, a, tiles_m);
    case EPI_SILU: return launch_tc_e<BN, NS, EPI_SILU>(h, st, c, a, tiles_m);
    case EPI_LN:
      if (BN == 256 && c.out.ld % 8 == 0 && !c.row_map && (tiles_m * LN_CL <= 2 * 148 || getenv("LADIFF_LN_CLUSTER_ALL")) && !getenv("LADIFF_NO_CLUSTER"))
        return launch_tc_ln<NS, EPI_LN>(h, st, c, a, tiles_m);
      if (BN == 256) return launch_tc_e<256, NS, EPI_LN>(h, st, c, a, tiles_m);
      break;
    case EPI_LN_MOD_SILU:
      if (BN == 256 && c.out.ld % 8 == 0 && !c.row_map && (tiles_m * LN_CL <= 2 * 148 || getenv("LADIFF_LN_CLUSTER_ALL")) && !getenv("LADIFF_NO_CLUSTER"))
        return launch_tc_ln<NS, EPI_LN_MOD_SILU>(h, st, c, a, tiles_m);
      if (BN == 256) return launch_tc_e<256, NS, EPI_LN_MOD_SILU>(h, st, c, a, tiles_m);
      break;
    default: break;
  }
  return h->err.set(LADIFF_ERR_INVALID, "unsupported epilogue %d for BN %d", c.epi, BN);
}

int launch_linear(H* h, cudaStream_t st, int mode, const LinCall& c) {
  if (c.no_pdl) g_skip_pdl_once = true;
  LinArgs a;
  memset(&a, 0, sizeof(a));
  const Weight& W = *c.W;
  a.M_max = c.M_max;
  a.M_dev = c.M_dev;
  a.N = W.N;
  a.K = W.K;
  a.K1 = c.A2 ? c.A->act.ld : W.K;
  a.K2 = c.A3 ? a.K1 + c.A2->act.ld : W.K;
  if (c.A3 && !c.A2) return h->err.set(LADIFF_ERR_INVALID, "linear: third source without a second");
  if (c.A->act.ld + (c.A2 ? c.A2->act.ld : 0) + (c.A3 ? c.A3->act.ld : 0) != W.K)
    return h->err.set(LADIFF_ERR_INVALID, "linear: source widths do not add up to K %d", W.K);
  a.A = c.A->act.f32;
  a.lda = c.A->act.ld;
  a.A2 = c.A2 ? c.A2->act.f32 : nullptr;
  a.lda2 = c.A2 ? c.A2->act.ld : 0;
  a.A3 = c.A3 ? c.A3->act.f32 : nullptr;
  a.lda3 = c.A3 ? c.A3->act.ld : 0;
  a.Wt = W.Wt;
  a.ldw = W.N;
  a.bias = W.bias;
  a.epi = c.epi;
  a.res = c.res;
  a.ldres = c.ldres;
  a.ln_g = c.ln_g;
  a.ln_b = c.ln_b;
  a.mod = c.mod;
  a.addv = c.addv;
  a.add_idx = c.add_idx;
  a.ld_add = c.ld_add;
  a.row_map = c.row_map;
  a.out = c.out;
  a.out_planes = c.out_planes;
  a.n_store = c.n_store < 0 ? W.N : c.n_store;
  a.dbg = c.dbg;
  a.trace = nullptr;
  if (h->trace && h->trace_n < h->trace_cap) {
    a.trace = h->trace + 8ull * h->trace_n;
    char nm[96];
    snprintf(nm, sizeof(nm), "lin M%d N%d K%d epi%d", c.M_max, W.N, W.K, c.epi);
    if (static_cast<int>(h->trace_names.size()) <= h->trace_n) h->trace_names.resize(h->trace_n + 1);
    h->trace_names[h->trace_n] = nm;
    h->trace_n++;
  }
  a.dbg_flags = getenv("LADIFF_DBG_FLAGS") ? atoi(getenv("LADIFF_DBG_FLAGS")) : 0;
  const bool ln = (c.epi == EPI_LN || c.epi == EPI_LN_MOD_SILU);
  if (ln && W.N != 256) return h->err.set(LADIFF_ERR_INVALID, "LayerNorm epilogue needs N == 256");
  if (c.M_max <= 0) return LADIFF_OK;
  if (mode == LADIFF_MODE_FP32) {
    if (!a.A || (c.A2 && !a.A2) || (c.A3 && !a.A3)) return h->err.set(LADIFF_ERR_STATE, "fp32 linear: missing fp32 operand");
    if (c.M_max >= 4096) {
      dim3 grid((c.M_max + 31) / 32, (W.N + 255) / 256);
      CK(launch_pdl(k_linear_simt<4>, grid, dim3(256), 0, st, a));
    } else {
      dim3 grid((c.M_max + 15) / 16, (W.N + 255) / 256);
      CK(launch_pdl(k_linear_simt<2>, grid, dim3(256), 0, st, a));
    }
    h->launches++;
    return LADIFF_OK;
  }
  if (!c.A->has_map || (c.A2 && !c.A2->has_map) || (c.A3 && !c.A3->has_map))
    return h->err.set(LADIFF_ERR_STATE, "tensor-core linear: operand has no bf16 planes");
  if (W.K % 64 != 0 || a.K1 % 64 != 0 || a.K2 % 64 != 0) return h->err.set(LADIFF_ERR_INVALID, "tensor-core linear: K must be a multiple of 64");
  a.a_plane_rows = c.A->act.rows_alloc;
  a.a2_plane_rows = c.A2 ? c.A2->act.rows_alloc : 0;
  a.a3_plane_rows = c.A3 ? c.A3->act.rows_alloc : 0;
  a.w_plane_rows = W.n_pad;
  const int tiles_m = (c.M_max + 127) / 128;
  // one CTA per SM (the stage ring takes ~190 KB): keep the grid within ONE wave of the 148 SMs, as narrow as that allows
  int bn = 256;
  if (!ln) {
    const int sms = getenv("LADIFF_TILE_SMS") ? atoi(getenv("LADIFF_TILE_SMS")) : 148;
    if (tiles_m * ((W.N + 63) / 64) <= (mode == LADIFF_MODE_BF16 ? 2 : 1) * sms) bn = 64;  // bf16: two 64-wide CTAs fit per SM
    else if (tiles_m * ((W.N + 127) / 128) <= sms) bn = 128;
  }
  const bool split = (mode == LADIFF_MODE_BF16X3);
  if (bn == 256) return split ? launch_tc<256, 2>(h, st, c, a, tiles_m) : launch_tc<256, 1>(h, st, c, a, tiles_m);
  if (bn == 128) return split ? launch_tc<128, 2>(h, st, c, a, tiles_m) : launch_tc<128, 1>(h, st, c, a, tiles_m);
  return split ? launch_tc<64, 2>(h, st, c, a, tiles_m) : launch_tc<64, 1>(h, st, c, a, tiles_m);
}

// ------------------------------------------------------------------------------------------------
// fused FFN pairs on a cluster of 4 CTAs (ffn_swap.cuh): up to two (W1, act, W2, LayerNorm-kind) pairs chained on one token group
struct FfnPair {
  const Weight* W1 = nullptr;
  const Weight* W2 = nullptr;
  int act = EPI_RELU, kind = EPI_LN;
  const float *ln_g = nullptr, *ln_b = nullptr, *mod = nullptr;
  Act out{nullptr, nullptr, 0, 0};
};
struct FfnCall {
  const ActBuf* X = nullptr;
  int M_max = 0;
  const int* M_dev = nullptr;
  int npairs = 0;
  FfnPair pair[2];
  const float* res = nullptr;   // pair 0
  const float* addv = nullptr;  // pair 0
  const int* add_idx = nullptr;
  int ld_add = 256;
  int out_planes = 0;
  long long* dbg = nullptr;
  // fused sa_block attention prologue (k_ffn_swap, rt == 48): X is computed in the kernel instead of being loaded
  bool att = false;
  const float *att_qkvx = nullptr, *att_textkv = nullptr, *att_timekv = nullptr, *att_res = nullptr;
  const float *att_bo = nullptr, *att_g = nullptr, *att_b = nullptr;
  const int *att_off = nullptr, *att_row_seq = nullptr;
  int att_ld_textkv = 0, att_ld_res = 0;
  Act att_x1{nullptr, nullptr, 0, 0}, att_xcopy{nullptr, nullptr, 0, 0};
};

bool ffn_fused_enabled() { return getenv("LADIFF_NO_FFN_FUSED") == nullptr; }

// Token-group size of the swapped kernel (ffn_swap.cuh): the smallest multiple of 16 whose clusters of 4 CTAs are all
// co-resident on the 148 SMs; 0 = too many rows: the plans then run the four separate fused linears (at >= 2560 rows their
// 128-row tiles fill the SMs and beat every cluster-fused variant: B = 256 reverse loop 36.8 vs 47.3 ms, B = 1024 107.6 vs 131 ms).
int ffn_swap_rt(int M_max) {
  if (getenv("LADIFF_NO_FFN_SWAP")) return 0;
  if (const char* e = getenv("LADIFF_FFN_RT")) {
    const int v = atoi(e);
    if (v == 16 || v == 32 || v == 48) return v;
  }
  for (int rt = 16; rt <= 48; rt += 16)
    if (((M_max + rt - 1) / rt) * 4 <= 148) return rt;
  return 0;
}
// the attention prologue of k_ffn_swap stages the keys of its 12 owned rows in the 48 KB h region: token groups of 48 only
// Opt-in (LADIFF_ATT_FUSE=1): measured SLOWER on B200 than the separate k_attn_ln launch (reverse loop 23.9 vs 22.0 ms at B = 128,
// profiles/r01g_trace_att_fused.txt): 12 rows per CTA on 256 threads serialise three dependent global round trips (~1.8 us each).
bool ffn_att_fusable(int M_max, int T) {
  const char* e = getenv("LADIFF_ATT_FUSE");
  return e && atoi(e) == 1 && T <= 5 && ffn_swap_rt(M_max) == 48;
}

int launch_ffn_swap(H* h, cudaStream_t st, int mode, const FfnCall& c) {
  if (mode == LADIFF_MODE_FP32) return h->err.set(LADIFF_ERR_INVALID, "the fused feed-forward kernel is a tensor-core path");
  if (c.npairs < 1 || c.npairs > 2 || !c.X || !c.X->has_map || c.X->act.ld != 256)
    return h->err.set(LADIFF_ERR_INVALID, "ffn swap: bad operands");
  if (c.M_max <= 0) return LADIFF_OK;
  FfnArgs a;
  memset(&a, 0, sizeof(a));
  a.M_max = c.M_max;
  a.M_dev = c.M_dev;
  a.npairs = c.npairs;
  for (int i = 0; i < c.npairs; ++i) {
    const FfnPair& q = c.pair[i];
    if (!q.W1 || !q.W2 || q.W1->N != 1024 || q.W1->K != 256 || q.W2->N != 256 || q.W2->K != 1024)
      return h->err.set(LADIFF_ERR_INVALID, "ffn swap: pair %d must be 256 -> 1024 -> 256", i);
    if (q.kind != EPI_LN && q.kind != EPI_LN_MOD_SILU) return h->err.set(LADIFF_ERR_INVALID, "ffn swap: bad epilogue kind");
    a.act[i] = q.act;
    a.kind[i] = q.kind;
    a.b1[i] = q.W1->bias;
    a.b2[i] = q.W2->bias;
    a.ln_g[i] = q.ln_g;
    a.ln_b[i] = q.ln_b;
    a.mod[i] = q.mod;
    a.out[i] = q.out;
    a.w1_plane_rows[i] = q.W1->n_pad;
    a.w2_plane_rows[i] = q.W2->n_pad;
    if (!q.W1->swap_img || !q.W2->swap_img) return h->err.set(LADIFF_ERR_INVALID, "ffn swap: pair %d has no streaming image", i);
    const size_t img_off = mode == LADIFF_MODE_BF16X3 ? 0 : 64ull * 8192;   // elements: the bf16 stages follow the x3 stages
    a.w1_img[i] = reinterpret_cast<const uint8_t*>(q.W1->swap_img + img_off);
    a.w2_img[i] = reinterpret_cast<const uint8_t*>(q.W2->swap_img + img_off);
  }
  a.res = c.res;
  a.addv = c.addv;
  a.add_idx = c.add_idx;
  a.ld_add = c.ld_add;
  a.out_planes = c.out_planes;
  a.x_plane_rows = c.X->act.rows_alloc;
  a.dbg = c.dbg;
  a.trace = nullptr;
  if (h->trace && h->trace_n < h->trace_cap) {
    a.trace = h->trace + 8ull * h->trace_n;
    char nm[96];
    snprintf(nm, sizeof(nm), "ffn_swap M%d pairs%d", c.M_max, c.npairs);
    if (static_cast<int>(h->trace_names.size()) <= h->trace_n) h->trace_names.resize(h->trace_n + 1);
    h->trace_names[h->trace_n] = nm;
    h->trace_n++;
  }
  const FfnPair& q0 = c.pair[0];
  const FfnPair& q1 = c.pair[c.npairs - 1];
  const int rt = ffn_swap_rt(c.M_max);
  if (rt == 0) return h->err.set(LADIFF_ERR_INVALID, "ffn swap: %d rows exceed the co-resident cluster capacity (callers use the separate linears)", c.M_max);
  if (c.att && rt != 48) return h->err.set(LADIFF_ERR_INVALID, "ffn swap: the fused attention prologue needs token groups of 48");
  a.rt = rt;
  if (c.att) {
    a.att = 1;
    a.att_qkvx = c.att_qkvx; a.att_off = c.att_off; a.att_row_seq = c.att_row_seq; a.att_textkv = c.att_textkv;
    a.att_ld_textkv = c.att_ld_textkv; a.att_timekv = c.att_timekv; a.att_res = c.att_res; a.att_ld_res = c.att_ld_res;
    a.att_bo = c.att_bo; a.att_g = c.att_g; a.att_b = c.att_b; a.att_x1 = c.att_x1; a.att_xcopy = c.att_xcopy;
  }
  if (a.trace) h->trace_names[h->trace_n - 1] = std::string(c.att ? "attn+ffn_swap M" : "ffn_swap M") + std::to_string(c.M_max) + " rt" + std::to_string(rt);
  dim3 grid((c.M_max + rt - 1) / rt, 4);
  const CUtensorMap& mx = rt == 16 ? c.X->map16 : (rt == 32 ? c.X->map32 : c.X->map48);
  auto kern = mode == LADIFF_MODE_BF16X3 ? (c.att ? k_ffn_swap<2, true> : k_ffn_swap<2, false>)
                                         : (c.att ? k_ffn_swap<1, true> : k_ffn_swap<1, false>);
  const int smem = mode == LADIFF_MODE_BF16X3 ? SwapCfg<2>::smem_bytes(rt) : SwapCfg<1>::smem_bytes(rt);
  CK(launch_pdl(kern, grid, dim3(SwapCfg<2>::THREADS), smem, st, mx, q0.W1->map128, q0.W2->map128, q1.W1->map128,
                q1.W2->map128, a));
  h->launches++;
  return LADIFF_OK;
}

// ------------------------------------------------------------------------------------------------
// fused feed-forward block on persistent 128-row tiles (ffn_tile.cuh): Linear(256 -> 1024) -> act -> Linear(1024 -> 256) -> LN kind
struct FtCall {
  const ActBuf* X = nullptr;
  const Weight *W1 = nullptr, *W2 = nullptr;
  int M_max = 0;
  const int* M_dev = nullptr;
  int act = EPI_GELU, kind = EPI_LN;
  const float *ln_g = nullptr, *ln_b = nullptr, *res = nullptr, *addv = nullptr, *mod = nullptr;
  const int* add_idx = nullptr;
  int ld_add = 256;
  Act out{nullptr, nullptr, 0, 0};
  int out_planes = 0;
};

bool ffn_tile_enabled(int mode) { return mode != LADIFF_MODE_FP32 && getenv("LADIFF_NO_FFN_TILE") == nullptr; }

int launch_ffn_tile(H* h, cudaStream_t st, int mode, const FtCall& c) {
  if (mode == LADIFF_MODE_FP32) return h->err.set(LADIFF_ERR_INVALID, "the fused feed-forward tile kernel is a tensor-core path");
  if (!c.X || !c.X->has_map || c.X->act.ld != 256 || !c.W1 || !c.W2 || c.W1->N != 1024 || c.W1->K != 256 || c.W2->N != 256 || c.W2->K != 1024 ||
      c.out.ld != 256 || (c.kind != EPI_LN && c.kind != EPI_LN_MOD_SILU) || (c.act != EPI_RELU && c.act != EPI_GELU))
    return h->err.set(LADIFF_ERR_INVALID, "ffn tile: bad operands (256 -> 1024 -> 256 blocks with a LayerNorm-kind epilogue only)");
  if (c.M_max <= 0) return LADIFF_OK;
  FtArgs a;
  memset(&a, 0, sizeof(a));
  a.M_max = c.M_max; a.M_dev = c.M_dev;
  a.b1 = c.W1->bias; a.b2 = c.W2->bias;
  a.act = c.act; a.kind = c.kind;
  a.ln_g = c.ln_g; a.ln_b = c.ln_b; a.res = c.res; a.addv = c.addv; a.add_idx = c.add_idx; a.ld_add = c.ld_add; a.mod = c.mod;
  a.out = c.out; a.out_planes = c.out_planes;
  a.x_plane_rows = c.X->act.rows_alloc; a.w1_plane_rows = c.W1->n_pad; a.w2_plane_rows = c.W2->n_pad;
  const int tiles = (c.M_max + 127) / 128;
  dim3 grid(tiles < 148 ? tiles : 148);
  static int dbg_left = getenv("LADIFF_FT_DBG") ? 1 : 0;     // profiling: clock stamps of CTA 0 of the first launch (needs LADIFF_NO_GRAPH=1)
  long long* dbg = nullptr;
  if (dbg_left > 0) {
    CK(cudaMalloc(&dbg, 256 * sizeof(long long)));
    CK(cudaMemsetAsync(dbg, 0, 256 * sizeof(long long), st));
    a.dbg = dbg;
  }
  if (mode == LADIFF_MODE_BF16X3)
    CK(launch_pdl(k_ffn_tile<2>, grid, dim3(FtCfg<2>::THREADS), FtCfg<2>::SMEM_BYTES, st, c.X->map, c.W1->map128, c.W2->map128, a));
  else
    CK(launch_pdl(k_ffn_tile<1>, grid, dim3(FtCfg<1>::THREADS), FtCfg<1>::SMEM_BYTES, st, c.X->map, c.W1->map128, c.W2->map128, a));
  h->launches++;
  if (dbg) {
    long long hb[256];
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost));
    cudaFree(dbg);
    --dbg_left;
    fprintf(stderr, "  ffn_tile cta 0 clk: x_full %lld | acco_full %lld | epilogue-0 done %lld | tile-1 x_full %lld | tile-1 acco_full %lld | epilogue-1 done %lld | end %lld\n   mma [wait-acch_empty g1-issued wait-h_full g2-issued] per chunk:", hb[1] - hb[0], hb[2] - hb[0], hb[4] - hb[0], hb[5] - hb[0], hb[6] - hb[0], hb[7] - hb[0], hb[3] - hb[0]);
    for (int j = 0; j < 8; ++j) fprintf(stderr, " | %lld %lld %lld %lld", hb[8 + 4 * j] ? hb[8 + 4 * j] - hb[0] : 0, hb[9 + 4 * j] ? hb[9 + 4 * j] - hb[0] : 0, hb[10 + 4 * j] - hb[0], hb[11 + 4 * j] - hb[0]);
    fprintf(stderr, "\n   epi [acch_full loaded+math h_empty-ok stored] per chunk:");
    for (int j = 0; j < 8; ++j) fprintf(stderr, " | %lld %lld %lld %lld", hb[128 + 4 * j] - hb[0], hb[129 + 4 * j] - hb[0], hb[130 + 4 * j] - hb[0], hb[131 + 4 * j] - hb[0]);
    fprintf(stderr, "\n   final epilogue 0: acc in regs %lld | res added %lld | first barrier %lld | stats done %lld", hb[200] - hb[0], hb[201] - hb[0], hb[202] - hb[0], hb[203] - hb[0]);
    fprintf(stderr, "\n   producer stage issue (every 4th):");
    for (int g = 4; g < 64; g += 4) fprintf(stderr, " %lld", hb[64 + g] - hb[0]);
    fprintf(stderr, "\n");
  }
  return LADIFF_OK;
}

#define LAUNCH(kernel, grid, block, smem, st, ...)  \
  do {                                              \
    kernel<<<grid, block, smem, st>>>(__VA_ARGS__); \
    CK(cudaGetLastError());                         \
    h->launches++;                                  \
  } while (0)

#define LAUNCHP(kernel, grid, block, smem, st, ...)                          \
  do {                                                                       \
    CK(launch_pdl(kernel, dim3(grid), dim3(block), smem, st, __VA_ARGS__));  \
    h->launches++;                                                           \
  } while (0)

inline unsigned cdiv(long a, long b) { return static_cast<unsigned>((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// weights
int get_raw(H* h, const std::string& name, std::initializer_list<int64_t> shape, const Raw** out) {
  auto it = h->raw.find(name);
  if (it == h->raw.end()) return h->err.set(LADIFF_ERR_WEIGHTS, "missing weight '%s'", name.c_str());
  const Raw& r = it->second;
  if (r.shape.size() != shape.size() || !std::equal(shape.begin(), shape.end(), r.shape.begin())) {
    std::string got;
    for (auto s : r.shape) got += std::to_string(s) + ",";
    std::string want;
    for (auto s : shape) want += std::to_string(s) + ",";
    return h->err.set(LADIFF_ERR_WEIGHTS, "size mismatch for %s: got [%s] expected [%s]", name.c_str(), got.c_str(), want.c_str());
  }
  *out = &r;
  return LADIFF_OK;
}

// Packs rows [row0, row0+N) of a [*, K] fp32 matrix (+ bias slice) into a Weight.
int pack_weight(H* h, Arena& ar, cudaStream_t st, Weight* w, const float* W_dev, int ldw, int N, int K, const float* bias_dev) {
  w->N = N;
  w->K = K;
  w->n_pad = roundup(N, 256);
  CK(ar.alloc(reinterpret_cast<void**>(&w->Wt), static_cast<size_t>(N) * K * sizeof(float)));
  CK(ar.alloc(reinterpret_cast<void**>(&w->pl), 3ull * w->n_pad * K * sizeof(op16)));   // fp16 hi | fp16 lo | bf16
  w->bias = nullptr;
  if (bias_dev) {
    CK(ar.alloc(reinterpret_cast<void**>(&w->bias), N * sizeof(float)));
    CK(cudaMemcpyAsync(w->bias, bias_dev, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  LAUNCH(k_pack_weight, cdiv(static_cast<long>(w->n_pad) * K, 256), 256, 0, st, W_dev, ldw, N, K, w->n_pad, w->Wt, w->pl);
  w->swap_img = nullptr;
  if ((N == 1024 && K == 256) || (N == 256 && K == 1024)) {   // a feed-forward matrix: also keep the k_ffn_swap streaming image
    CK(ar.alloc(reinterpret_cast<void**>(&w->swap_img), 3ull * N * K * sizeof(op16)));
    LAUNCH(k_pack_swap_image, cdiv(3L * 32 * 128 * 64, 256), 256, 0, st, w->pl, w->n_pad, K, N == 256 ? 1 : 0, w->swap_img);
  }
  if (K % 64 == 0) {
    CKS(make_map(h, &w->map64, w->pl, 3ull * w->n_pad, K, K, 64));
    CKS(make_map(h, &w->map128, w->pl, 3ull * w->n_pad, K, K, 128));
    CKS(make_map(h, &w->map256, w->pl, 3ull * w->n_pad, K, K, 256));
  }
  return LADIFF_OK;
}

int pack_linear(H* h, cudaStream_t st, Weight* w, const std::string& prefix, int N, int K) {
  const Raw *W, *b;
  CKS(get_raw(h, prefix + ".weight", {N, K}, &W));
  CKS(get_raw(h, prefix + ".bias", {N}, &b));
  return pack_weight(h, *h->warena, st, w, W->dev, K, N, K, b->dev);
}

int get_vec(H* h, const std::string& name, int n, float** out) {
  const Raw* r;
  CKS(get_raw(h, name, {n}, &r));
  *out = r->dev;
  return LADIFF_OK;
}

// concatenates row-slices [row0,row0+rows) of `count` same-shaped matrices (+ bias slices) and packs them as one Weight
int pack_concat(H* h, cudaStream_t st, Weight* w, const std::vector<std::pair<const Raw*, const Raw*>>& parts, int row0,
                int rows, int K) {
  const int N = rows * static_cast<int>(parts.size());
  float *tmpW = nullptr, *tmpB = nullptr;
  CK(cudaMalloc(&tmpW, static_cast<size_t>(N) * K * sizeof(float)));
  CK(cudaMalloc(&tmpB, N * sizeof(float)));
  for (size_t i = 0; i < parts.size(); ++i) {
    CK(cudaMemcpyAsync(tmpW + i * rows * K, parts[i].first->dev + static_cast<size_t>(row0) * K,
                       static_cast<size_t>(rows) * K * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(tmpB + i * rows, parts[i].second->dev + row0, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  int s = pack_weight(h, *h->warena, st, w, tmpW, K, N, K, tmpB);
  CK(cudaStreamSynchronize(st));
  cudaFree(tmpW);
  cudaFree(tmpB);
  return s;
}

// Folded layer transitions.  With Y_l = x3 + s P_l^T + p_l (StylizationBlock out-projection + residual,
// mdiff_transformer.py:162,262) the next layer's input tokens X and their extended in-projection E (q | k | 4 x v', see
// k_attn_ln) are linear in (x3, s[, skip]):
//   l+1 <= 4:  X = Y_l                                  -> rows [ I | P_l ],              bias p_l
//   l+1 >= 5:  X = Y_l S1^T + skip S2^T + sb            -> rows [ S1 | S1 P_l | S2 ],     bias S1 p_l + sb
//              (cat + Linear(512->256), cross_attention.py:79-81)
//   E-part  :  E X-rows, bias E (X-bias) + e
// One GEMM per layer transition (N = 1536 + 256, K = 512 or 768; K sources ordered x3, s, skip) replaces the residual GEMM,
// the skip Linear and the in-projection.  Products are formed once here (fp64 accumulate) and packed like any other weight.
int fold_denoiser_transitions(H* h, cudaStream_t st) {
  const int D = 256, NQ = DQ_LD, NX = DQX_LD;
  const std::string P = "denoiser.encoder.";
  float *tmp = nullptr, *tb = nullptr;
  CK(cudaMalloc(&tmp, static_cast<size_t>(NX) * 768 * sizeof(float)));
  CK(cudaMalloc(&tb, NX * sizeof(float)));
  auto done = [&](int s) {
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    cudaFree(tb);
    return s;
  };
  for (int l = 0; l + 1 < NL; ++l) {
    const Raw *Pw, *Pb;
    const float *E = h->den[l + 1].inx, *e = h->den[l + 1].inx_bias;  // extended in-projection of layer l+1
    int s_;
    if ((s_ = get_raw(h, P + block_name(l) + ".ffn.proj_out.out_layers.2.weight", {D, D}, &Pw)) != LADIFF_OK) return done(s_);
    if ((s_ = get_raw(h, P + block_name(l) + ".ffn.proj_out.out_layers.2.bias", {D}, &Pb)) != LADIFF_OK) return done(s_);
    const int K = (l + 1 <= 4) ? 2 * D : 3 * D;
    float* X = tmp + static_cast<size_t>(NQ) * K;  // X-rows [256, K] live at rows [1536, 1792) of the packed matrix
    float* xb = tb + NQ;
    if (l + 1 <= 4) {
      LAUNCH(k_set_identity, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, X, K, D);
      LAUNCH(k_copy2d, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, Pw->dev, D, D, D, X + D, K);
      CK(cudaMemcpyAsync(xb, Pb->dev, D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      const int i = l - 4;
      const Raw *S, *sb;
      if ((s_ = get_raw(h, P + "linear_blocks." + std::to_string(i) + ".weight", {D, 2 * D}, &S)) != LADIFF_OK) return done(s_);
      if ((s_ = get_raw(h, P + "linear_blocks." + std::to_string(i) + ".bias", {D}, &sb)) != LADIFF_OK) return done(s_);
      LAUNCH(k_copy2d, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, S->dev, 2 * D, D, D, X, K);
      LAUNCH(k_fold_matmul, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, S->dev, 2 * D, Pw->dev, D, D, D, D, X + D, K);
      LAUNCH(k_copy2d, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, S->dev + D, 2 * D, D, D, X + 2 * D, K);
      LAUNCH(k_fold_matvec, 1, 256, 0, st, S->dev, 2 * D, Pb->dev, sb->dev, D, D, xb);
    }
    LAUNCH(k_fold_matmul, cdiv(static_cast<long>(NQ) * K, 256), 256, 0, st, E, D, X, K, NQ, D, K, tmp, K);
    LAUNCH(k_fold_matvec, cdiv(NQ, 256), 256, 0, st, E, D, xb, e, NQ, D, tb);
    if ((s_ = pack_weight(h, *h->warena, st, &h->den_qkv_fold[l + 1], tmp, K, NX, K, tb)) != LADIFF_OK) return done(s_);
    CK(cudaStreamSynchronize(st));
  }
  return done(LADIFF_OK);
}

int finalize_denoiser(H* h, cudaStream_t st) {
  const int D = 256;
  h->den_ready = false;
  h->warena_den.clear();
  h->warena = &h->warena_den;
  const std::string P = "denoiser.";
  CKS(pack_linear(h, st, &h->time1, P + "time_embedding.linear_1", D, 768));
  CKS(pack_linear(h, st, &h->time2, P + "time_embedding.linear_2", D, D));
  CKS(pack_linear(h, st, &h->embproj, P + "emb_proj.1", D, 768));
  const Raw* pe;
  CKS(get_raw(h, P + "query_pos.pe", {500, 1, D}, &pe));
  h->den_pe = pe->dev;
  CKS(get_vec(h, P + "encoder.norm.weight", D, &h->den_fg));
  CKS(get_vec(h, P + "encoder.norm.bias", D, &h->den_fb));
  std::vector<std::pair<const Raw*, const Raw*>> inproj, mods;
  std::vector<Raw> inx_raw(NL), inxb_raw(NL);
  for (int l = 0; l < NL; ++l) {
    const std::string L = P + "encoder." + block_name(l) + ".";
    DenLayerW& w = h->den[l];
    const Raw *ipw, *ipb;
    CKS(get_raw(h, L + "sa_block.self_attn.in_proj_weight", {3 * D, D}, &ipw));
    CKS(get_raw(h, L + "sa_block.self_attn.in_proj_bias", {3 * D}, &ipb));
    // extended in-projection: the out-projection is folded into the per-head values (see k_attn_ln)
    const Raw *ow, *ob;
    CKS(get_raw(h, L + "sa_block.self_attn.out_proj.weight", {D, D}, &ow));
    CKS(get_raw(h, L + "sa_block.self_attn.out_proj.bias", {D}, &ob));
    w.out_bias = ob->dev;
    CK(h->warena->alloc(reinterpret_cast<void**>(&w.inx), static_cast<size_t>(DQ_LD) * D * sizeof(float)));
    CK(h->warena->alloc(reinterpret_cast<void**>(&w.inx_bias), DQ_LD * sizeof(float)));
    CK(cudaMemcpyAsync(w.inx, ipw->dev, 2ull * D * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(w.inx_bias, ipb->dev, 2 * D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int hd = 0; hd < 4; ++hd) {
      LAUNCH(k_fold_matmul, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, ow->dev + hd * 64, D, ipw->dev + (2 * D + hd * 64) * D, D, D, 64, D,
             w.inx + (2 * D + hd * D) * D, D);
      LAUNCH(k_fold_matvec, 1, 256, 0, st, ow->dev + hd * 64, D, ipb->dev + 2 * D + hd * 64, (const float*)nullptr, D, 64,
             w.inx_bias + 2 * D + hd * D);
    }
    CKS(pack_weight(h, *h->warena, st, &w.qkv, w.inx, D, DQ_LD, D, w.inx_bias));
    inx_raw[l].dev = w.inx;
    inx_raw[l].shape = {DQ_LD, D};
    inxb_raw[l].dev = w.inx_bias;
    inxb_raw[l].shape = {DQ_LD};
    inproj.push_back({&inx_raw[l], &inxb_raw[l]});
    CKS(pack_linear(h, st, &w.ff1, L + "sa_block.linear1", 1024, D));
    CKS(pack_linear(h, st, &w.ff2, L + "sa_block.linear2", D, 1024));
    CKS(get_vec(h, L + "sa_block.norm1.weight", D, &w.n1g));
    CKS(get_vec(h, L + "sa_block.norm1.bias", D, &w.n1b));
    CKS(get_vec(h, L + "sa_block.norm2.weight", D, &w.n2g));
    CKS(get_vec(h, L + "sa_block.norm2.bias", D, &w.n2b));
    // ca_block: with one text token softmax over tokens == 1, so query/key/norm are mathematically dead
    // (mdiff_transformer.py:231-245); they must still be present in the state_dict like in the reference.
    const Raw* dead;
    CKS(get_raw(h, L + "ca_block.query.weight", {D, D}, &dead));
    CKS(get_raw(h, L + "ca_block.key.weight", {D, D}, &dead));
    CKS(pack_linear(h, st, &w.ca_value, L + "ca_block.value", D, D));
    CKS(get_vec(h, L + "ca_block.text_norm.weight", D, &w.ca_tn_g));
    CKS(get_vec(h, L + "ca_block.text_norm.bias", D, &w.ca_tn_b));
    CKS(get_vec(h, L + "ca_block.proj_out.norm.weight", D, &w.ca_sn_g));
    CKS(get_vec(h, L + "ca_block.proj_out.norm.bias", D, &w.ca_sn_b));
    CKS(pack_linear(h, st, &w.ca_out, L + "ca_block.proj_out.out_layers.2", D, D));
    CKS(pack_linear(h, st, &w.gff1, L + "ffn.linear1", h->cfg.ff_size, D));
    CKS(pack_linear(h, st, &w.gff2, L + "ffn.linear2", D, h->cfg.ff_size));
    CKS(get_vec(h, L + "ffn.proj_out.norm.weight", D, &w.ffn_sn_g));
    CKS(get_vec(h, L + "ffn.proj_out.norm.bias", D, &w.ffn_sn_b));
    CKS(pack_linear(h, st, &w.ffn_out, L + "ffn.proj_out.out_layers.2", D, D));
    const Raw *cw, *cb, *fw, *fb;
    CKS(get_raw(h, L + "ca_block.proj_out.emb_layers.1.weight", {2 * D, D}, &cw));
    CKS(get_raw(h, L + "ca_block.proj_out.emb_layers.1.bias", {2 * D}, &cb));
    CKS(get_raw(h, L + "ffn.proj_out.emb_layers.1.weight", {2 * D, D}, &fw));
    CKS(get_raw(h, L + "ffn.proj_out.emb_layers.1.bias", {2 * D}, &fb));
    mods.push_back({cw, cb});
    mods.push_back({fw, fb});
  }
  for (int i = 0; i < 4; ++i) CKS(pack_linear(h, st, &h->den_skip[i], P + "encoder.linear_blocks." + std::to_string(i), D, 2 * D));
  CKS(fold_denoiser_transitions(h, st));
  // K/V rows of every layer's in_proj: the conditioning tokens are step- or prompt-invariant
  CKS(pack_concat(h, st, &h->timekv_all, inproj, D, DC_LD, D));  // per layer: k | 4 x v' rows of the extended in-projection
  h->textkv_all = h->timekv_all;  // same matrix, applied to the text projection
  CKS(pack_concat(h, st, &h->mod_all, mods, 0, 2 * D, D));
  h->den_ready = true;
  return LADIFF_OK;
}

int finalize_decoder(H* h, cudaStream_t st) {
  const int D = 256;
  h->dec_ready = false;
  h->warena_dec.clear();
  h->warena = &h->warena_dec;
  const std::string P = "vae.";
  const Raw* pe;
  CKS(get_raw(h, P + "query_pos_decoder.pe", {500, 1, D}, &pe));
  h->dec_pe = pe->dev;
  CKS(get_vec(h, P + "decoder.norm.weight", D, &h->dec_fg));
  CKS(get_vec(h, P + "decoder.norm.bias", D, &h->dec_fb));
  // folded cross-attention projection of the memory latents (k_cross_ln): per layer CX_LD rows
  //   [h * 256 + c]        kq : (W_q[h]^T W_k[h])[c, :]      bias  W_q[h]^T b_k[h]
  //   [1024 + h * 256 + c] v' : (W_o[:, h] W_v[h])[c, :]     bias  W_o[:, h] b_v[h]
  //   [2048 + h]           cq :  b_q[h]^T W_k[h]             bias  b_q[h] . b_k[h]
  float *fm = nullptr, *fb = nullptr;
  CK(cudaMalloc(&fm, static_cast<size_t>(NL) * CX_LD * D * sizeof(float)));
  CK(cudaMalloc(&fb, static_cast<size_t>(NL) * CX_LD * sizeof(float)));
  CK(cudaMemsetAsync(fm, 0, static_cast<size_t>(NL) * CX_LD * D * sizeof(float), st));
  CK(cudaMemsetAsync(fb, 0, static_cast<size_t>(NL) * CX_LD * sizeof(float), st));
  auto fold_done = [&](int s_) {
    cudaStreamSynchronize(st);
    cudaFree(fm);
    cudaFree(fb);
    return s_;
  };
  for (int l = 0; l < NL; ++l) {
    const std::string L = P + "decoder." + block_name(l) + ".";
    DecLayerW& w = h->dec[l];
    const Raw *ipw, *ipb, *mw, *mb, *ow, *ob;
    int s_;
    if ((s_ = get_raw(h, L + "self_attn.in_proj_weight", {3 * D, D}, &ipw)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_raw(h, L + "self_attn.in_proj_bias", {3 * D}, &ipb)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = pack_weight(h, *h->warena, st, &w.qkv, ipw->dev, D, 3 * D, D, ipb->dev)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = pack_linear(h, st, &w.out, L + "self_attn.out_proj", D, D)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_raw(h, L + "multihead_attn.in_proj_weight", {3 * D, D}, &mw)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_raw(h, L + "multihead_attn.in_proj_bias", {3 * D}, &mb)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_raw(h, L + "multihead_attn.out_proj.weight", {D, D}, &ow)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_raw(h, L + "multihead_attn.out_proj.bias", {D}, &ob)) != LADIFF_OK) return fold_done(s_);
    w.out2_bias = ob->dev;
    float* F = fm + static_cast<size_t>(l) * CX_LD * D;
    float* Fb = fb + static_cast<size_t>(l) * CX_LD;
    const float *Wq = mw->dev, *Wk = mw->dev + D * D, *Wv = mw->dev + 2 * D * D;
    const float *bq = mb->dev, *bk = mb->dev + D, *bv = mb->dev + 2 * D;
    for (int hd = 0; hd < 4; ++hd) {
      const int r = hd * 64;
      LAUNCH(k_fold_tn, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, Wq + r * D, D, Wk + r * D, D, D, 64, D, F + (hd * D) * D, D);
      LAUNCH(k_fold_tn, 1, 256, 0, st, Wq + r * D, D, bk + r, 1, D, 64, 1, Fb + hd * D, 1);
      LAUNCH(k_fold_matmul, cdiv(static_cast<long>(D) * D, 256), 256, 0, st, ow->dev + r, D, Wv + r * D, D, D, 64, D, F + (1024 + hd * D) * D, D);
      LAUNCH(k_fold_matvec, 1, 256, 0, st, ow->dev + r, D, bv + r, (const float*)nullptr, D, 64, Fb + 1024 + hd * D);
      LAUNCH(k_fold_matmul, 1, 256, 0, st, bq + r, 64, Wk + r * D, D, 1, 64, D, F + (2048 + hd) * D, D);
      LAUNCH(k_fold_matvec, 1, 256, 0, st, bq + r, 64, bk + r, (const float*)nullptr, 1, 64, Fb + 2048 + hd);
    }
    if ((s_ = pack_linear(h, st, &w.ff1, L + "linear1", h->cfg.ff_size, D)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = pack_linear(h, st, &w.ff2, L + "linear2", D, h->cfg.ff_size)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_vec(h, L + "norm1.weight", D, &w.n1g)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_vec(h, L + "norm1.bias", D, &w.n1b)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_vec(h, L + "norm2.weight", D, &w.n2g)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_vec(h, L + "norm2.bias", D, &w.n2b)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_vec(h, L + "norm3.weight", D, &w.n3g)) != LADIFF_OK) return fold_done(s_);
    if ((s_ = get_vec(h, L + "norm3.bias", D, &w.n3b)) != LADIFF_OK) return fold_done(s_);
  }
  {
    int s_ = pack_weight(h, *h->warena, st, &h->memx_all, fm, D, NL * CX_LD, D, fb);
    if (fold_done(s_) != LADIFF_OK) return s_;
  }
  for (int i = 0; i < 4; ++i) CKS(pack_linear(h, st, &h->dec_skip[i], P + "decoder.linear_blocks." + std::to_string(i), D, 2 * D));
  CKS(pack_linear(h, st, &h->dec_final, P + "final_layer", h->cfg.nfeats, D));
  h->dec_ready = true;
  // ---- encoder half of the VAE (LADiffVae.encode; only when the caller supplied its keys)
  h->enc_ready = false;
  if (h->raw.count(P + "skel_embedding.weight") && h->raw.count(P + "encoder.norm.weight")) {
    const int nf = h->cfg.nfeats, Kp = roundup(nf, 64), T = h->cfg.max_it;
    const Raw *sw, *sb, *gmt, *epe;
    CKS(get_raw(h, P + "skel_embedding.weight", {D, nf}, &sw));
    CKS(get_raw(h, P + "skel_embedding.bias", {D}, &sb));
    CKS(get_raw(h, P + "global_motion_token", {2 * T, D}, &gmt));
    CKS(get_raw(h, P + "query_pos_encoder.pe", {500, 1, D}, &epe));
    h->enc_gmt = gmt->dev;
    h->enc_pe = epe->dev;
    h->skel_kp = Kp;
    float* tmp = nullptr;   // [256, Kp] zero-padded copy: the tensor path needs K % 64 == 0
    CK(cudaMalloc(&tmp, static_cast<size_t>(D) * Kp * sizeof(float)));
    CK(cudaMemsetAsync(tmp, 0, static_cast<size_t>(D) * Kp * sizeof(float), st));
    LAUNCH(k_copy2d, cdiv(static_cast<long>(D) * nf, 256), 256, 0, st, sw->dev, nf, D, nf, tmp, Kp);
    int s_ = pack_weight(h, *h->warena, st, &h->skel, tmp, Kp, D, Kp, sb->dev);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    CKS(s_);
    CKS(get_vec(h, P + "encoder.norm.weight", D, &h->enc_fg));
    CKS(get_vec(h, P + "encoder.norm.bias", D, &h->enc_fb));
    for (int l = 0; l < NL; ++l) {
      const std::string L = P + "encoder." + block_name(l) + ".";
      EncLayerW& w = h->enc[l];
      const Raw *ipw, *ipb;
      CKS(get_raw(h, L + "self_attn.in_proj_weight", {3 * D, D}, &ipw));
      CKS(get_raw(h, L + "self_attn.in_proj_bias", {3 * D}, &ipb));
      CKS(pack_weight(h, *h->warena, st, &w.qkv, ipw->dev, D, 3 * D, D, ipb->dev));
      CKS(pack_linear(h, st, &w.out, L + "self_attn.out_proj", D, D));
      CKS(pack_linear(h, st, &w.ff1, L + "linear1", h->cfg.ff_size, D));
      CKS(pack_linear(h, st, &w.ff2, L + "linear2", D, h->cfg.ff_size));
      CKS(get_vec(h, L + "norm1.weight", D, &w.n1g));
      CKS(get_vec(h, L + "norm1.bias", D, &w.n1b));
      CKS(get_vec(h, L + "norm2.weight", D, &w.n2g));
      CKS(get_vec(h, L + "norm2.bias", D, &w.n2b));
    }
    for (int i = 0; i < 4; ++i) CKS(pack_linear(h, st, &h->enc_skip[i], P + "encoder.linear_blocks." + std::to_string(i), D, 2 * D));
    h->enc_ready = true;
  }
  return LADIFF_OK;
}

// ------------------------------------------------------------------------------------------------
// denoiser plan
struct DenoisePlan {
  Arena ar;
  int S = 0, B = 0, n = 0, mode = 0, T = 0, Rmax = 0;
  bool cfg = false;
  int planes = 0;  // bf16 planes written by producers: 0 fp32 mode, 1 bf16, 2 bf16x3
  // meta
  int *cnt = nullptr, *off = nullptr, *R = nullptr, *row_seq = nullptr, *row_t = nullptr, *ts = nullptr;
  float* coef = nullptr;  // [n][4]: c1, c2, c3 (variance-noise coefficient, 0 for DDIM eta = 0), guidance
  std::vector<int> cached_ts;
  std::vector<float> cached_coef;
  // scheduler-step extras shared by all chains of a ReversePlan (device scalars so that one captured graph serves every call)
  const float** noise_pp = nullptr;        // injected per-step variance noise [n, Btot, T, 256], or *noise_pp == null -> Philox
  unsigned long long* seed_p = nullptr;    // Philox seed
  int ar_mode = 0;                         // ARDIFF: only latent slot 0 is denoised
  // inputs staged in the workspace so the captured graph has fixed addresses
  float *text768 = nullptr, *lat = nullptr;
  // time side
  ActBuf sin, t1, temb, st;
  float *timekv = nullptr, *mod = nullptr;
  // text side
  ActBuf trelu, textp, tn, ca_a;
  float *textkv = nullptr, *lny = nullptr, *delta = nullptr;
  // activations
  ActBuf xin, xa, xb, x1, x3, skip[4], a, hbuf, sbuf;
  float* qkv = nullptr;
  // chained use (ReversePlan): inputs live in the parent's staging buffers, time tables are shared with chain 0
  const float* text_src = nullptr;  // parent text [2*Btot,768]; this chain covers prompts [b0, b0+B)
  int text_Btot = 0, b0 = 0, chain = 0;
  const int* cnt_src = nullptr;     // parent m[Btot] + b0
  uint64_t last_use = 0;
  cudaGraphExec_t exec = nullptr;
  int64_t graph_launches = 0;
  ~DenoisePlan() {
    if (exec) cudaGraphExecDestroy(exec);
  }
};

// LADIFF._diffusion_reverse for one (B, steps, mode, guidance): the batch is cut into independent chains of prompts
// (sequences never interact), each with its own workspace, enqueued on forked streams inside ONE captured graph so that
// the latency-bound per-chain kernel sequences overlap on the 148 SMs.
struct ReversePlan {
  Arena ar;
  int B = 0, n = 0, mode = 0;
  int* cnt = nullptr;        // m[B]
  float* text768 = nullptr;  // [2B,768]
  float* lat = nullptr;      // [B,T,256]
  std::vector<std::unique_ptr<DenoisePlan>> chains;
  std::vector<int> cached_ts;
  std::vector<float> cached_coef;
  const float** noise_pp = nullptr;
  unsigned long long* seed_p = nullptr;
  uint64_t last_use = 0;
  cudaGraphExec_t exec = nullptr;
  int64_t graph_launches = 0;
  ~ReversePlan() {
    if (exec) cudaGraphExecDestroy(exec);
  }
};

Act f32_only(float* p, int ld) { return Act{p, nullptr, ld, 0}; }

int build_denoise_plan(H* h, DenoisePlan* p, int S, int n, int mode, bool cfg, float* ext_lat = nullptr,
                       const DenoisePlan* tables = nullptr) {
  p->S = S;
  p->B = cfg ? S / 2 : 0;
  p->n = n;
  p->mode = mode;
  p->cfg = cfg;
  p->T = h->cfg.max_it;
  p->Rmax = S * p->T;
  p->planes = mode == LADIFF_MODE_FP32 ? 0 : (mode == LADIFF_MODE_BF16X3 ? 2 : 1);
  const bool tcm = mode != LADIFF_MODE_FP32, f = !tcm;
  Arena& ar = p->ar;
  const int R = p->Rmax;
  CK(ar.alloc((void**)&p->cnt, S * sizeof(int)));
  CK(ar.alloc((void**)&p->off, (S + 1) * sizeof(int)));
  CK(ar.alloc((void**)&p->R, sizeof(int)));
  CK(ar.alloc((void**)&p->row_seq, R * sizeof(int)));
  CK(ar.alloc((void**)&p->row_t, R * sizeof(int)));
  if (tables) {  // step-only tables are identical for every chain of a ReversePlan
    p->ts = tables->ts;
    p->coef = tables->coef;
    p->timekv = tables->timekv;
    p->mod = tables->mod;
  } else {
    CK(ar.alloc((void**)&p->ts, n * sizeof(int)));
    CK(ar.alloc((void**)&p->coef, 4 * n * sizeof(float)));
    CKS(alloc_act(h, ar, &p->sin, n, 768, f, tcm));
    CKS(alloc_act(h, ar, &p->t1, n, 256, f, tcm));
    CKS(alloc_act(h, ar, &p->temb, n, 256, true, tcm));
    CKS(alloc_act(h, ar, &p->st, n, 256, f, tcm));
    CK(ar.alloc((void**)&p->timekv, static_cast<size_t>(n) * NL * DC_LD * sizeof(float)));
    CK(ar.alloc((void**)&p->mod, static_cast<size_t>(n) * NL * 1024 * sizeof(float)));
  }
  if (ext_lat) {
    p->lat = ext_lat;
  } else {
    CK(ar.alloc((void**)&p->text768, static_cast<size_t>(S) * 768 * sizeof(float)));
    if (cfg) CK(ar.alloc((void**)&p->lat, static_cast<size_t>(p->B) * p->T * 256 * sizeof(float)));
  }
  CKS(alloc_act(h, ar, &p->trelu, S, 768, f, tcm));
  CKS(alloc_act(h, ar, &p->textp, S, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->tn, S, 256, f, tcm));
  CKS(alloc_act(h, ar, &p->ca_a, n * S, 256, f, tcm));
  CK(ar.alloc((void**)&p->textkv, static_cast<size_t>(S) * NL * DC_LD * sizeof(float)));
  CK(ar.alloc((void**)&p->lny, static_cast<size_t>(NL) * S * 256 * sizeof(float)));
  CK(ar.alloc((void**)&p->delta, static_cast<size_t>(NL) * n * S * 256 * sizeof(float)));
  CKS(alloc_act(h, ar, &p->xin, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->xa, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->xb, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->x1, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->x3, R, 256, true, tcm));
  for (int i = 0; i < 4; ++i) CKS(alloc_act(h, ar, &p->skip[i], R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->a, R, 256, f, tcm));
  CKS(alloc_act(h, ar, &p->hbuf, R, 1024, f, tcm));
  CKS(alloc_act(h, ar, &p->sbuf, R, 256, f, tcm));
  CK(ar.alloc((void**)&p->qkv, static_cast<size_t>(roundup(R, 128)) * DQX_LD * sizeof(float)));
  return LADIFF_OK;
}

// time tables: depend on (timesteps, weights) only -> cached across calls (SURVEY.md 8a hoist table)
int enqueue_time_tables(H* h, DenoisePlan* p, cudaStream_t st) {
  const int n = p->n, mode = p->mode, pl = p->planes;
  LAUNCHP(k_sinus_embed, cdiv(n * 768, 256), 256, 0, st, p->ts, n, p->sin.act, pl);
  LinCall c;
  c.A = &p->sin; c.W = &h->time1; c.M_max = n; c.epi = EPI_SILU; c.out = p->t1.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->t1; c.W = &h->time2; c.M_max = n; c.out = p->temb.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  LAUNCHP(k_unary, cdiv(n * 256, 256), 256, 0, st, p->temb.act.f32, 256, n, 256, (int)U_SILU, p->st.act, pl);
  c = LinCall(); c.A = &p->temb; c.W = &h->timekv_all; c.M_max = n; c.out = f32_only(p->timekv, NL * DC_LD);
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->st; c.W = &h->mod_all; c.M_max = n; c.out = f32_only(p->mod, NL * 1024);
  CKS(launch_linear(h, st, mode, c));
  return LADIFF_OK;
}

// text tables + the hoisted ca_block contribution for every (step, layer, sequence)
int enqueue_text_tables(H* h, DenoisePlan* p, cudaStream_t st) {
  const int S = p->S, n = p->n, mode = p->mode, pl = p->planes;
  if (p->text_src)
    LAUNCHP(k_text_relu, cdiv(static_cast<long>(S) * 768, 256), 256, 0, st, p->text_src, p->text_Btot, p->b0, p->B, p->trelu.act, pl);
  else
    LAUNCHP(k_unary, cdiv(static_cast<long>(S) * 768, 256), 256, 0, st, p->text768, 768, S, 768, (int)U_RELU, p->trelu.act, pl);
  LinCall c;
  c.A = &p->trelu; c.W = &h->embproj; c.M_max = S; c.out = p->textp.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->textp; c.W = &h->textkv_all; c.M_max = S; c.out = f32_only(p->textkv, NL * DC_LD);
  CKS(launch_linear(h, st, mode, c));
  for (int l = 0; l < NL; ++l) {
    const DenLayerW& w = h->den[l];
    LAUNCHP(k_layernorm256, cdiv(static_cast<long>(S) * 32, 256), 256, 0, st, p->textp.act.f32, 256, S, (const int*)nullptr,
           w.ca_tn_g, w.ca_tn_b, p->tn.act, pl);
    float* lny = p->lny + static_cast<size_t>(l) * S * 256;
    c = LinCall(); c.A = &p->tn; c.W = &w.ca_value; c.M_max = S; c.epi = EPI_LN; c.ln_g = w.ca_sn_g; c.ln_b = w.ca_sn_b;
    c.out = f32_only(lny, 256);
    CKS(launch_linear(h, st, mode, c));
    LAUNCHP(k_ca_prologue, cdiv(static_cast<long>(n) * S * 256, 256), 256, 0, st, lny, 256, p->mod + l * 1024, NL * 1024, n, S,
           p->ca_a.act, pl);
    c = LinCall(); c.A = &p->ca_a; c.W = &w.ca_out; c.M_max = n * S;
    c.out = f32_only(p->delta + static_cast<size_t>(l) * n * S * 256, 256);
    CKS(launch_linear(h, st, mode, c));
  }
  return LADIFF_OK;
}

// One denoiser layer after its in-projection (p->qkv rows hold q | k | 4 x v' [| X] of the layer input): fused sa_block
// attention + out-proj + residual + LN (k_attn_ln) -> ReLU FFN (+LN, + hoisted ca_block delta) -> GELU FFN -> Stylization
// prologue.  Leaves x3 (p->x3) and s (p->sbuf); the layer output Y = x3 + s P^T + p is never materialised except for the
// last layer (it is folded into the next layer's in-projection, see fold_denoiser_transitions).
int enqueue_den_layer(H* h, DenoisePlan* p, cudaStream_t st, int l, int step, const float* res, int ld_res, const Act& xcopy) {
  const DenLayerW& w = h->den[l];
  const int mode = p->mode, pl = p->planes, R = p->Rmax, S = p->S, n = p->n;
  LinCall c;
  unsigned long long* tr = nullptr;
  const bool fused_ffn = mode != LADIFF_MODE_FP32 && ffn_fused_enabled() && ffn_swap_rt(R) > 0;
  const bool fuse_att = fused_ffn && ffn_att_fusable(R, p->T);
  if (!fuse_att && h->trace && h->trace_n < h->trace_cap) {
    tr = h->trace + 8ull * h->trace_n;
    if (static_cast<int>(h->trace_names.size()) <= h->trace_n) h->trace_names.resize(h->trace_n + 1);
    h->trace_names[h->trace_n++] = "attn_ln M" + std::to_string(R) + " ";
  }
  if (fuse_att) {
    // attention + out-proj + residual + LN run as the prologue of the FFN kernel (one launch per layer after the in-projection)
  } else if (p->T <= 5 && !getenv("LADIFF_ATTN_LN_MAXT8")) {   // (the env switch runs the 8-latent instantiation on <= 5 latents: test hook)
    if (S > 148 && S <= 296 && getenv("LADIFF_ATTN_2SEQ")) {   // two sequences per CTA: one wave of <= 148 CTAs
      LAUNCHP((k_attn_ln<5, 2>), (S + 1) / 2, 512, 0, st, p->qkv, p->off, S, p->textkv + l * DC_LD, NL * DC_LD,
           p->timekv + static_cast<size_t>(step) * NL * DC_LD + l * DC_LD, res, ld_res, w.out_bias, w.n1g, w.n1b, p->x1.act, xcopy, pl, tr);
    } else if (getenv("LADIFF_ATTN_LN_128")) {   // opt-in: measured slower (20.87 vs 19.55 ms, profiles/r02_attn_ln_occupancy.txt)
      // 128 threads per sequence: two CTAs fit next to the in-projection CTA of their SM, so every sequence is resident before it ends
      LAUNCHP((k_attn_ln2<5>), S, 128, 0, st, p->qkv, p->off, S, p->textkv + l * DC_LD, NL * DC_LD,
           p->timekv + static_cast<size_t>(step) * NL * DC_LD + l * DC_LD, res, ld_res, w.out_bias, w.n1g, w.n1b, p->x1.act, xcopy, pl, tr);
    } else {
      LAUNCHP((k_attn_ln<5, 1>), S, 256, 0, st, p->qkv, p->off, S, p->textkv + l * DC_LD, NL * DC_LD,
           p->timekv + static_cast<size_t>(step) * NL * DC_LD + l * DC_LD, res, ld_res, w.out_bias, w.n1g, w.n1b, p->x1.act, xcopy, pl, tr);
    }
  } else
    LAUNCHP((k_attn_ln<8, 1>), S, 256, 0, st, p->qkv, p->off, S, p->textkv + l * DC_LD, NL * DC_LD,
           p->timekv + static_cast<size_t>(step) * NL * DC_LD + l * DC_LD, res, ld_res, w.out_bias, w.n1g, w.n1b, p->x1.act, xcopy, pl, tr);
  if (fused_ffn) {
    // both feed-forward pairs of the layer in ONE cluster kernel: x1 -> x3 (fp32 + planes) -> s (planes); h never leaves the SM
    FfnCall f;
    f.X = &p->x1; f.M_max = R; f.M_dev = p->R; f.npairs = 2; f.out_planes = pl;
    f.res = p->x1.act.f32;
    f.addv = p->delta + (static_cast<size_t>(l) * n + step) * S * 256; f.add_idx = p->row_seq; f.ld_add = 256;
    f.pair[0].W1 = &w.ff1; f.pair[0].W2 = &w.ff2; f.pair[0].act = EPI_RELU; f.pair[0].kind = EPI_LN;
    f.pair[0].ln_g = w.n2g; f.pair[0].ln_b = w.n2b; f.pair[0].out = p->x3.act;
    f.pair[1].W1 = &w.gff1; f.pair[1].W2 = &w.gff2; f.pair[1].act = EPI_GELU; f.pair[1].kind = EPI_LN_MOD_SILU;
    f.pair[1].ln_g = w.ffn_sn_g; f.pair[1].ln_b = w.ffn_sn_b;
    f.pair[1].mod = p->mod + static_cast<size_t>(step) * NL * 1024 + l * 1024 + 512; f.pair[1].out = p->sbuf.act;
    if (fuse_att) {
      f.att = true;
      f.att_qkvx = p->qkv; f.att_off = p->off; f.att_row_seq = p->row_seq;
      f.att_textkv = p->textkv + l * DC_LD; f.att_ld_textkv = NL * DC_LD;
      f.att_timekv = p->timekv + static_cast<size_t>(step) * NL * DC_LD + l * DC_LD;
      f.att_res = res; f.att_ld_res = ld_res;
      f.att_bo = w.out_bias; f.att_g = w.n1g; f.att_b = w.n1b;
      f.att_x1 = p->x1.act; f.att_xcopy = xcopy;
    }
    return launch_ffn_swap(h, st, mode, f);
  }
  if (ffn_tile_enabled(mode) && (R + 127) / 128 >= 120 && !getenv("LADIFF_NO_FFN_TILE_DEN")) {
    // enough 128-row tiles to fill the SMs (>= 15 360 latent rows per chain): each feed-forward pair as one persistent tile kernel.
    // Below that the separate linears win: 20 tiles (B = 256 per chain) would occupy 20 SMs for ~30 us (measured at B = 1024 in four
    // chains: reverse loop 154 ms with tile kernels vs 115 ms with the separate linears, profiles/r02_large_batch.txt)
    FtCall f;
    f.X = &p->x1; f.W1 = &w.ff1; f.W2 = &w.ff2; f.M_max = R; f.M_dev = p->R; f.act = EPI_RELU; f.kind = EPI_LN;
    f.res = p->x1.act.f32; f.ln_g = w.n2g; f.ln_b = w.n2b;
    f.addv = p->delta + (static_cast<size_t>(l) * n + step) * S * 256; f.add_idx = p->row_seq; f.ld_add = 256;
    f.out = p->x3.act; f.out_planes = pl;
    CKS(launch_ffn_tile(h, st, mode, f));
    f = FtCall();
    f.X = &p->x3; f.W1 = &w.gff1; f.W2 = &w.gff2; f.M_max = R; f.M_dev = p->R; f.act = EPI_GELU; f.kind = EPI_LN_MOD_SILU;
    f.ln_g = w.ffn_sn_g; f.ln_b = w.ffn_sn_b; f.mod = p->mod + static_cast<size_t>(step) * NL * 1024 + l * 1024 + 512;
    f.out = p->sbuf.act; f.out_planes = pl;
    return launch_ffn_tile(h, st, mode, f);
  }
  c = LinCall(); c.A = &p->x1; c.W = &w.ff1; c.M_max = R; c.M_dev = p->R; c.epi = EPI_RELU; c.out = p->hbuf.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->hbuf; c.W = &w.ff2; c.M_max = R; c.M_dev = p->R; c.epi = EPI_LN; c.res = p->x1.act.f32;
  c.ln_g = w.n2g; c.ln_b = w.n2b;
  c.addv = p->delta + (static_cast<size_t>(l) * n + step) * S * 256; c.add_idx = p->row_seq; c.ld_add = 256;
  c.out = p->x3.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->x3; c.W = &w.gff1; c.M_max = R; c.M_dev = p->R; c.epi = EPI_GELU; c.out = p->hbuf.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->hbuf; c.W = &w.gff2; c.M_max = R; c.M_dev = p->R; c.epi = EPI_LN_MOD_SILU;
  c.ln_g = w.ffn_sn_g; c.ln_b = w.ffn_sn_b; c.mod = p->mod + static_cast<size_t>(step) * NL * 1024 + l * 1024 + 512;
  c.out = p->sbuf.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  return LADIFF_OK;
}

// SkipTransformerEncoder wiring (operator/cross_attention.py:69-85); the last layer's tokens end up in p->xa.
// Six dependent launches per layer: in-projection (folded with the previous layer's output projection, residual and skip
// merge) -> attention+LN -> FFN1 -> FFN2+LN -> FFN1' -> FFN2'+Stylization.
int enqueue_den_tokens(H* h, DenoisePlan* p, cudaStream_t st, int step) {
  const int mode = p->mode, R = p->Rmax;
  const Act none{nullptr, nullptr, 0, 0};
  LinCall c;
  c.A = &p->xin; c.W = &h->den[0].qkv; c.M_max = R; c.M_dev = p->R; c.out = f32_only(p->qkv, DQX_LD);
  CKS(launch_linear(h, st, mode, c));
  for (int l = 0; l < NL; ++l) {
    // layer input X_l: the packed latents for l = 0, else the X columns of the folded in-projection; X_1..X_4 (= Y_0..Y_3)
    // are also copied out (planes) for the U-Net skip connections of layers 8..5
    const float* res = l == 0 ? p->xin.act.f32 : p->qkv + DQ_LD;
    Act xc = (l >= 1 && l <= 4) ? p->skip[l - 1].act : none;
    if (mode != LADIFF_MODE_FP32 && xc.pl) xc.f32 = nullptr;   // tensor-core modes read the skip source as operand planes only
    CKS(enqueue_den_layer(h, p, st, l, step, res, l == 0 ? 256 : DQX_LD, xc));
    const DenLayerW& w = h->den[l];
    if (l == NL - 1) {  // last layer: Y_8 itself (input of encoder.norm in k_cfg_ddim / k_final_ln_out)
      c = LinCall(); c.A = &p->sbuf; c.W = &w.ffn_out; c.M_max = R; c.M_dev = p->R; c.epi = EPI_RES; c.res = p->x3.act.f32;
      c.out = f32_only(p->xa.act.f32, 256);
      CKS(launch_linear(h, st, mode, c));
      break;
    }
    c = LinCall(); c.A = &p->x3; c.A2 = &p->sbuf; if (l >= 4) c.A3 = &p->skip[7 - l];
    c.W = &h->den_qkv_fold[l + 1]; c.M_max = R; c.M_dev = p->R; c.out = f32_only(p->qkv, DQX_LD);
    CKS(launch_linear(h, st, mode, c));
  }
  return LADIFF_OK;
}

int enqueue_meta(H* h, cudaStream_t st, const int* cnt_host, int S, int* cnt, int* off, int* R, int* row_seq, int* row_t,
                 int* row_dst, int dst_stride) {
  CK(cudaMemcpyAsync(cnt, cnt_host, S * sizeof(int), cudaMemcpyHostToDevice, st));
  LAUNCH(k_scan_counts, 1, 1024, 0, st, cnt, S, S, off, R);
  LAUNCH(k_fill_rows, cdiv(static_cast<long>(S) * 32, 256), 256, 0, st, off, S, row_seq, row_t, row_dst, dst_stride);
  return LADIFF_OK;
}

int enqueue_reverse_body(H* h, DenoisePlan* p, cudaStream_t st) {
  CKS(enqueue_text_tables(h, p, st));
  LAUNCHP(k_pack_x, cdiv(static_cast<long>(p->Rmax) * 256, 256), 256, 0, st, p->lat, p->B, p->T, h->den_pe, p->row_seq, p->row_t, p->R,
         p->xin.act, p->planes);
  for (int step = 0; step < p->n; ++step) {
    CKS(enqueue_den_tokens(h, p, st, step));
    LAUNCHP(k_cfg_ddim, cdiv(static_cast<long>(p->B) * p->T * 32, 256), 256, 0, st, p->xa.act.f32, p->off, p->B, p->T, h->den_fg, h->den_fb,
           p->coef + 4 * step, p->lat, h->den_pe, p->xin.act, p->planes, p->ar_mode, step, p->noise_pp, p->seed_p,
           static_cast<long>(p->b0) * p->T * 256, static_cast<long>(p->text_Btot) * p->T * 256);
  }
  return LADIFF_OK;
}

// all chains of a ReversePlan: chain 0 on `st`, the others on forked side streams joined back into `st`
int enqueue_reverse_chains(H* h, ReversePlan* rp, cudaStream_t st) {
  const int nch = static_cast<int>(rp->chains.size());
  if (nch > 1) {
    if (!h->ev_fork) CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CK(cudaEventRecord(h->ev_fork, st));
  }
  for (int c = 0; c < nch; ++c) {
    DenoisePlan* p = rp->chains[c].get();
    cudaStream_t sc = st;
    if (c > 0) {
      if (!h->side[c]) CK(cudaStreamCreateWithFlags(&h->side[c], cudaStreamNonBlocking));
      if (!h->ev_join[c]) CK(cudaEventCreateWithFlags(&h->ev_join[c], cudaEventDisableTiming));
      sc = h->side[c];
      CK(cudaStreamWaitEvent(sc, h->ev_fork, 0));
    }
    LAUNCH(k_scan_counts, 1, 1024, 0, sc, p->cnt_src, p->S, p->B, p->off, p->R);
    LAUNCH(k_fill_rows, cdiv(static_cast<long>(p->S) * 32, 256), 256, 0, sc, p->off, p->S, p->row_seq, p->row_t, (int*)nullptr, 0);
    CKS(enqueue_reverse_body(h, p, sc));
    if (c > 0) {
      CK(cudaEventRecord(h->ev_join[c], sc));
      CK(cudaStreamWaitEvent(st, h->ev_join[c], 0));
    }
  }
  return LADIFF_OK;
}

int pick_chains(int B) {
  if (const char* e = getenv("LADIFF_CHAINS")) {
    int v = atoi(e);
    if (v >= 1) return v > MAX_CHAINS ? MAX_CHAINS : (v > B ? B : v);
  }
  // Measured on B200 (profiles/r01b_chains.txt): at B = 128 every link is latency-bound and a chain of 32 prompts takes as
  // long as one of 128, so splitting gains nothing; at B = 1024 four chains of 256 prompts overlap one chain's epilogues /
  // launch gaps with another's mainloops (reverse 107 -> 84 ms).
  // Round 2 (same-box sweeps, profiles/r02_large_batch.txt): above 177 prompts one chain no longer fits the co-resident cluster
  // feed-forward kernel (1776 latent rows), and two chains that do (B = 256: 2 x 128) beat one chain on the separate linears
  // (reverse 36.3 -> 31.7 ms in the x3 mode, 26.2 -> 24.3 ms in bf16); B = 384: 45.9 -> 44.9 ms; from B = 512 on, chains of
  // 256 stay best (B = 512: 54.6 ms with 2 chains against 57.3 / 59.2 with 3 / 4; B = 768: 78.0 with 3; B = 1024: 110 with 4).
  // Between 129 and 177 prompts one chain still fits the cluster kernel but no longer one wave of the in-projection / attention
  // grids: B = 136 / 144: 20.9 / 21.9 ms with one chain against 22.9 / 24.1 with two; B = 160 / 177: 31.6 / 32.0 against 25.3 / 26.1.
  int n = B / 256;
  if (n < 2 && B >= 150) n = 2;
  return n < 1 ? 1 : (n > 4 ? 4 : n);
}

int build_reverse_plan(H* h, ReversePlan* rp, int B, int n, int mode, int ar_flag) {
  rp->B = B;
  rp->n = n;
  rp->mode = mode;
  const int T = h->cfg.max_it;
  CK(rp->ar.alloc((void**)&rp->cnt, B * sizeof(int)));
  CK(rp->ar.alloc((void**)&rp->noise_pp, sizeof(float*)));
  CK(rp->ar.alloc((void**)&rp->seed_p, sizeof(unsigned long long)));
  CK(cudaMemset(rp->noise_pp, 0, sizeof(float*)));
  CK(cudaMemset(rp->seed_p, 0, sizeof(unsigned long long)));
  CK(rp->ar.alloc((void**)&rp->text768, static_cast<size_t>(2 * B) * 768 * sizeof(float)));
  CK(rp->ar.alloc((void**)&rp->lat, static_cast<size_t>(B) * T * 256 * sizeof(float)));
  const int nch = pick_chains(B);
  for (int c = 0; c < nch; ++c) {
    const int b0 = static_cast<int>(static_cast<long>(B) * c / nch), b1 = static_cast<int>(static_cast<long>(B) * (c + 1) / nch);
    std::unique_ptr<DenoisePlan> p(new DenoisePlan());
    CKS(build_denoise_plan(h, p.get(), 2 * (b1 - b0), n, mode, true, rp->lat + static_cast<size_t>(b0) * T * 256,
                           c ? rp->chains[0].get() : nullptr));
    p->text_src = rp->text768;
    p->text_Btot = B;
    p->b0 = b0;
    p->cnt_src = rp->cnt + b0;
    p->chain = c;
    p->noise_pp = rp->noise_pp;
    p->seed_p = rp->seed_p;
    p->ar_mode = ar_flag;
    rp->chains.push_back(std::move(p));
  }
  return LADIFF_OK;
}

// ------------------------------------------------------------------------------------------------
// decoder plan
struct DecodePlan {
  Arena ar;
  int B = 0, mode = 0, T = 0, Lmax = 0, Rmax = 0, Mmax = 0, planes = 0;
  int *cnt = nullptr, *foff = nullptr, *Rf = nullptr, *frow_seq = nullptr, *frow_t = nullptr, *frow_dst = nullptr;
  int *mcnt = nullptr, *moff = nullptr, *Rm = nullptr, *mrow_seq = nullptr, *mrow_t = nullptr;
  float* z = nullptr;  // staged [T,B,256]
  ActBuf zrows, x0, xa, xb, x1, x2, skip[4], a, hbuf, xn, qkvp;   // qkvp: q | k | v operand planes of the self-attention (tensor-core modes)
  float *qkv = nullptr, *memx = nullptr;   // memx: folded cross-attention table [memory rows, 9 x CX_LD]
  uint64_t last_use = 0;
  cudaGraphExec_t exec = nullptr;
  int64_t graph_launches = 0;
  ~DecodePlan() {
    if (exec) cudaGraphExecDestroy(exec);
  }
};

int build_decode_plan(H* h, DecodePlan* p, int B, int mode) {
  p->B = B;
  p->mode = mode;
  p->T = h->cfg.max_it;
  p->Lmax = h->cfg.max_frames;
  p->Rmax = B * p->Lmax;
  p->Mmax = B * p->T;
  p->planes = mode == LADIFF_MODE_FP32 ? 0 : (mode == LADIFF_MODE_BF16X3 ? 2 : 1);
  const bool tcm = mode != LADIFF_MODE_FP32, f = !tcm;
  Arena& ar = p->ar;
  const int R = p->Rmax, M = p->Mmax;
  CK(ar.alloc((void**)&p->cnt, B * sizeof(int)));
  CK(ar.alloc((void**)&p->foff, (B + 1) * sizeof(int)));
  CK(ar.alloc((void**)&p->Rf, sizeof(int)));
  CK(ar.alloc((void**)&p->frow_seq, R * sizeof(int)));
  CK(ar.alloc((void**)&p->frow_t, R * sizeof(int)));
  CK(ar.alloc((void**)&p->frow_dst, R * sizeof(int)));
  CK(ar.alloc((void**)&p->mcnt, B * sizeof(int)));
  CK(ar.alloc((void**)&p->moff, (B + 1) * sizeof(int)));
  CK(ar.alloc((void**)&p->Rm, sizeof(int)));
  CK(ar.alloc((void**)&p->mrow_seq, M * sizeof(int)));
  CK(ar.alloc((void**)&p->mrow_t, M * sizeof(int)));
  CK(ar.alloc((void**)&p->z, static_cast<size_t>(M) * 256 * sizeof(float)));
  CKS(alloc_act(h, ar, &p->zrows, M, 256, f, tcm));
  CKS(alloc_act(h, ar, &p->x0, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->xa, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->xb, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->x1, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->x2, R, 256, true, tcm));
  for (int i = 0; i < 4; ++i) CKS(alloc_act(h, ar, &p->skip[i], R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->a, R, 256, f, tcm));
  CKS(alloc_act(h, ar, &p->hbuf, R, 1024, f, tcm));
  CKS(alloc_act(h, ar, &p->xn, R, 256, f, tcm));
  CK(ar.alloc((void**)&p->qkv, static_cast<size_t>(roundup(R, 128)) * 768 * sizeof(float)));
  if (tcm) CKS(alloc_act(h, ar, &p->qkvp, R, 768, false, true));
  CK(ar.alloc((void**)&p->memx, static_cast<size_t>(roundup(M, 128)) * NL * CX_LD * sizeof(float)));
  return LADIFF_OK;
}

// ragged self-attention of B sequences of at most Lmax rows (qkv [rows, 768], off[B + 1]) -> out [rows, 256]
// tensor-core modes: the in-projection writes q | k | v as 16-bit operand planes (qkvp) and the attention kernel pulls them by TMA
bool attn_planes_input(int mode) {
  return mode != LADIFF_MODE_FP32 && !getenv("LADIFF_ATTN_SIMT") && !getenv("LADIFF_ATTN_MMA_SYNC") && !getenv("LADIFF_ATTN_F32_STAGE");
}
int enqueue_self_attention(H* h, cudaStream_t st, int mode, int pl, int B, int Lmax, const float* qkv, const ActBuf* qkvp, const int* off, const Act& out) {
  if (mode == LADIFF_MODE_FP32 || getenv("LADIFF_ATTN_SIMT")) {
    dim3 grid((Lmax + SA_QB - 1) / SA_QB, 4, B);
    LAUNCHP(k_attn_self, grid, 256, sizeof(SelfAttnSmem), st, qkv, off, out, pl);
  } else if (!getenv("LADIFF_ATTN_MMA_SYNC")) {
    // tensor-core modes: (head, sequence) CTAs on tcgen05 (scores and probabilities live in tensor memory)
    dim3 grid(4, B);
    long long* dbg = nullptr;
    static int dbg_left = getenv("LADIFF_AT5_DBG") ? 2 : 0;    // profiling: clock stamps of CTA (0, 0) of the first launches (needs LADIFF_NO_GRAPH=1)
    if (dbg_left > 0) {
      CK(cudaMalloc(&dbg, 32 * sizeof(long long)));
      CK(cudaMemsetAsync(dbg, 0, 32 * sizeof(long long), st));
    }
    if (attn_planes_input(mode)) {
      if (!qkvp || !qkvp->has_map) return h->err.set(LADIFF_ERR_STATE, "self-attention: the q | k | v operand planes are missing");
      const int prow = qkvp->act.rows_alloc;
      if (mode == LADIFF_MODE_BF16X3) LAUNCHP((k_attn_self_t5<2, true>), grid, At5Cfg<2>::THREADS, At5Cfg<2>::SMEM_BYTES, st, qkvp->map, prow, qkv, off, out, pl, dbg);
      else LAUNCHP((k_attn_self_t5<1, true>), grid, At5Cfg<1>::THREADS, At5Cfg<1>::SMEM_BYTES, st, qkvp->map, prow, qkv, off, out, pl, dbg);
    } else {
      const CUtensorMap dummy{};
      if (mode == LADIFF_MODE_BF16X3) LAUNCHP((k_attn_self_t5<2, false>), grid, At5Cfg<2>::THREADS, At5Cfg<2>::SMEM_BYTES, st, dummy, 0, qkv, off, out, pl, dbg);
      else LAUNCHP((k_attn_self_t5<1, false>), grid, At5Cfg<1>::THREADS, At5Cfg<1>::SMEM_BYTES, st, dummy, 0, qkv, off, out, pl, dbg);
    }
    if (dbg) {
      long long hb[32];
      CK(cudaStreamSynchronize(st));
      CK(cudaMemcpy(hb, dbg, sizeof(hb), cudaMemcpyDeviceToHost));
      cudaFree(dbg);
      --dbg_left;
      fprintf(stderr, "  attn_t5 cta(0,0) clk:");
      for (int i = 1; i < 32; ++i)
        if (hb[i]) fprintf(stderr, " [%d]%lld", i, hb[i] - hb[0]);
      fprintf(stderr, "\n");
    }
  } else {
    // the round-1 kernel (mma.sync fragments, scores in registers), kept for same-box A/B: LADIFF_ATTN_MMA_SYNC=1
    dim3 grid(4, B);
    const bool big = Lmax > 26 * 8;
    if (mode == LADIFF_MODE_BF16X3) {
      if (big) LAUNCHP((k_attn_self_tc<32, 2>), grid, 256, sat_smem_bytes<32>(2), st, qkv, off, out, pl);
      else LAUNCHP((k_attn_self_tc<26, 2>), grid, 256, sat_smem_bytes<26>(2), st, qkv, off, out, pl);
    } else {
      if (big) LAUNCHP((k_attn_self_tc<32, 1>), grid, 256, sat_smem_bytes<32>(1), st, qkv, off, out, pl);
      else LAUNCHP((k_attn_self_tc<26, 1>), grid, 256, sat_smem_bytes<26>(1), st, qkv, off, out, pl);
    }
  }
  return LADIFF_OK;
}

int enqueue_dec_layer(H* h, DecodePlan* p, cudaStream_t st, int l, const ActBuf& in, const ActBuf& out) {
  const DecLayerW& w = h->dec[l];
  const int mode = p->mode, pl = p->planes, R = p->Rmax;
  LinCall c;
  c.A = &in; c.W = &w.qkv; c.M_max = R; c.M_dev = p->Rf; c.out = f32_only(p->qkv, 768);
  if (attn_planes_input(mode)) { c.out = Act{nullptr, p->qkvp.act.pl, 768, p->qkvp.act.rows_alloc}; c.out_planes = pl; }
  CKS(launch_linear(h, st, mode, c));
  CKS(enqueue_self_attention(h, st, mode, pl, p->B, p->Lmax, p->qkv, &p->qkvp, p->foff, p->a.act));
  c = LinCall(); c.A = &p->a; c.W = &w.out; c.M_max = R; c.M_dev = p->Rf; c.epi = EPI_LN; c.res = in.act.f32;
  c.ln_g = w.n1g; c.ln_b = w.n1b; c.out = p->x1.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  // cross-attention to the <= 5 latents + residual + norm2: one fp32 kernel on the folded table (no q / out-projection GEMMs)
  {
    dim3 grid((p->Lmax + 31) / 32, p->B);
    if (p->T <= 5)
      LAUNCHP(k_cross_ln<5>, grid, 256, 5 * CX_LD * sizeof(float), st, p->x1.act.f32, p->memx, NL * CX_LD, l * CX_LD, p->foff, p->moff, w.out2_bias,
              w.n2g, w.n2b, p->x2.act, pl);
    else
      LAUNCHP(k_cross_ln<8>, grid, 256, 8 * CX_LD * sizeof(float), st, p->x1.act.f32, p->memx, NL * CX_LD, l * CX_LD, p->foff, p->moff, w.out2_bias,
              w.n2g, w.n2b, p->x2.act, pl);
  }
  if (ffn_tile_enabled(mode)) {   // linear1 -> GELU -> linear2 -> + residual -> norm3, the 1024-wide hidden stays in tensor memory
    FtCall f;
    f.X = &p->x2; f.W1 = &w.ff1; f.W2 = &w.ff2; f.M_max = R; f.M_dev = p->Rf; f.act = EPI_GELU; f.kind = EPI_LN;
    f.res = p->x2.act.f32; f.ln_g = w.n3g; f.ln_b = w.n3b; f.out = out.act; f.out_planes = pl;
    return launch_ffn_tile(h, st, mode, f);
  }
  c = LinCall(); c.A = &p->x2; c.W = &w.ff1; c.M_max = R; c.M_dev = p->Rf; c.epi = EPI_GELU; c.out = p->hbuf.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->hbuf; c.W = &w.ff2; c.M_max = R; c.M_dev = p->Rf; c.epi = EPI_LN; c.res = p->x2.act.f32;
  c.ln_g = w.n3g; c.ln_b = w.n3b; c.out = out.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  return LADIFF_OK;
}

// everything of vae.decode between the staged latent z and the final LayerNorm'd tokens (p->xn)
int enqueue_decode_body(H* h, DecodePlan* p, cudaStream_t st) {
  const int pl = p->planes, mode = p->mode;
  LAUNCHP(k_gather_z, cdiv(static_cast<long>(p->Mmax) * 256, 256), 256, 0, st, p->z, p->moff, p->B, p->T, p->zrows.act, pl);
  LinCall c;
  c.A = &p->zrows; c.W = &h->memx_all; c.M_max = p->Mmax; c.M_dev = p->Rm; c.out = f32_only(p->memx, NL * CX_LD);
  CKS(launch_linear(h, st, mode, c));
  LAUNCHP(k_dec_init, cdiv(static_cast<long>(p->Rmax) * 32, 256), 256, 0, st, h->dec_pe, p->frow_t, p->Rf, p->x0.act, pl);
  const ActBuf* x = &p->x0;
  for (int i = 0; i < 4; ++i) {
    CKS(enqueue_dec_layer(h, p, st, i, *x, p->skip[i]));
    x = &p->skip[i];
  }
  CKS(enqueue_dec_layer(h, p, st, 4, *x, p->xa));
  for (int i = 0; i < 4; ++i) {
    c = LinCall();
    c.A = &p->xa; c.A2 = &p->skip[3 - i]; c.W = &h->dec_skip[i]; c.M_max = p->Rmax; c.M_dev = p->Rf;
    c.out = p->xb.act; c.out_planes = pl;
    CKS(launch_linear(h, st, mode, c));
    CKS(enqueue_dec_layer(h, p, st, 5 + i, p->xb, p->xa));
  }
  LAUNCHP(k_layernorm256, cdiv(static_cast<long>(p->Rmax) * 32, 256), 256, 0, st, p->xa.act.f32, 256, p->Rmax, p->Rf, h->dec_fg, h->dec_fb,
         p->xn.act, pl);
  return LADIFF_OK;
}

// ------------------------------------------------------------------------------------------------
// encoder plan (LADiffVae.encode): token rows = sum_b (2 m_b + L_b)
struct EncodePlan {
  Arena ar;
  int B = 0, mode = 0, T = 0, Lmax = 0, Rmax = 0, Fmax = 0, planes = 0;
  int *cnt = nullptr, *off = nullptr, *R = nullptr, *row_seq = nullptr, *row_t = nullptr;
  int *fcnt = nullptr, *foff = nullptr, *Rf = nullptr, *frow_seq = nullptr, *frow_t = nullptr, *mcnt = nullptr;
  ActBuf feats, x0, xa, xb, x1, skip[4], a, hbuf, qkvp;
  float *emb = nullptr, *qkv = nullptr, *xn = nullptr;
  uint64_t last_use = 0;
};

int build_encode_plan(H* h, EncodePlan* p, int B, int mode) {
  p->B = B;
  p->mode = mode;
  p->T = h->cfg.max_it;
  p->Lmax = h->cfg.max_frames + 2 * p->T;
  p->Fmax = B * h->cfg.max_frames;
  p->Rmax = B * p->Lmax;
  p->planes = mode == LADIFF_MODE_FP32 ? 0 : (mode == LADIFF_MODE_BF16X3 ? 2 : 1);
  const bool tcm = mode != LADIFF_MODE_FP32, f = !tcm;
  Arena& ar = p->ar;
  const int R = p->Rmax, F = p->Fmax;
  CK(ar.alloc((void**)&p->cnt, B * sizeof(int)));
  CK(ar.alloc((void**)&p->off, (B + 1) * sizeof(int)));
  CK(ar.alloc((void**)&p->R, sizeof(int)));
  CK(ar.alloc((void**)&p->row_seq, R * sizeof(int)));
  CK(ar.alloc((void**)&p->row_t, R * sizeof(int)));
  CK(ar.alloc((void**)&p->fcnt, B * sizeof(int)));
  CK(ar.alloc((void**)&p->foff, (B + 1) * sizeof(int)));
  CK(ar.alloc((void**)&p->Rf, sizeof(int)));
  CK(ar.alloc((void**)&p->frow_seq, F * sizeof(int)));
  CK(ar.alloc((void**)&p->frow_t, F * sizeof(int)));
  CK(ar.alloc((void**)&p->mcnt, B * sizeof(int)));
  CKS(alloc_act(h, ar, &p->feats, F, h->skel_kp, f, tcm));
  CK(ar.alloc((void**)&p->emb, static_cast<size_t>(roundup(F, 128)) * 256 * sizeof(float)));
  CKS(alloc_act(h, ar, &p->x0, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->xa, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->xb, R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->x1, R, 256, true, tcm));
  for (int i = 0; i < 4; ++i) CKS(alloc_act(h, ar, &p->skip[i], R, 256, true, tcm));
  CKS(alloc_act(h, ar, &p->a, R, 256, f, tcm));
  CKS(alloc_act(h, ar, &p->hbuf, R, 1024, f, tcm));
  CK(ar.alloc((void**)&p->qkv, static_cast<size_t>(roundup(R, 128)) * 768 * sizeof(float)));
  if (tcm) CKS(alloc_act(h, ar, &p->qkvp, R, 768, false, true));
  CK(ar.alloc((void**)&p->xn, static_cast<size_t>(roundup(R, 128)) * 256 * sizeof(float)));
  return LADIFF_OK;
}

int enqueue_self_attention(H* h, cudaStream_t st, int mode, int planes, int B, int Lmax, const float* qkv, const ActBuf* qkvp, const int* off, const Act& out);
bool attn_planes_input(int mode);

// TransformerEncoderLayer.forward_post (operator/cross_attention.py:293-307, gelu): self-attention -> +res -> LN -> FFN -> +res -> LN
int enqueue_enc_layer(H* h, EncodePlan* p, cudaStream_t st, int l, const ActBuf& in, const ActBuf& out) {
  const EncLayerW& w = h->enc[l];
  const int mode = p->mode, pl = p->planes, R = p->Rmax;
  LinCall c;
  c.A = &in; c.W = &w.qkv; c.M_max = R; c.M_dev = p->R; c.out = f32_only(p->qkv, 768);
  if (attn_planes_input(mode)) { c.out = Act{nullptr, p->qkvp.act.pl, 768, p->qkvp.act.rows_alloc}; c.out_planes = pl; }
  CKS(launch_linear(h, st, mode, c));
  CKS(enqueue_self_attention(h, st, mode, pl, p->B, p->Lmax, p->qkv, &p->qkvp, p->off, p->a.act));
  c = LinCall(); c.A = &p->a; c.W = &w.out; c.M_max = R; c.M_dev = p->R; c.epi = EPI_LN; c.res = in.act.f32;
  c.ln_g = w.n1g; c.ln_b = w.n1b; c.out = p->x1.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  if (ffn_tile_enabled(mode)) {
    FtCall f;
    f.X = &p->x1; f.W1 = &w.ff1; f.W2 = &w.ff2; f.M_max = R; f.M_dev = p->R; f.act = EPI_GELU; f.kind = EPI_LN;
    f.res = p->x1.act.f32; f.ln_g = w.n2g; f.ln_b = w.n2b; f.out = out.act; f.out_planes = pl;
    return launch_ffn_tile(h, st, mode, f);
  }
  c = LinCall(); c.A = &p->x1; c.W = &w.ff1; c.M_max = R; c.M_dev = p->R; c.epi = EPI_GELU; c.out = p->hbuf.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  c = LinCall(); c.A = &p->hbuf; c.W = &w.ff2; c.M_max = R; c.M_dev = p->R; c.epi = EPI_LN; c.res = p->x1.act.f32;
  c.ln_g = w.n2g; c.ln_b = w.n2b; c.out = out.act; c.out_planes = pl;
  CKS(launch_linear(h, st, mode, c));
  return LADIFF_OK;
}

// SkipTransformerEncoder.forward, non-MD branch (operator/cross_attention.py:48-67) over the ragged token rows
int enqueue_encode_body(H* h, EncodePlan* p, cudaStream_t st, const float* feats_dev, int max_len) {
  const int pl = p->planes, mode = p->mode, Kp = h->skel_kp;
  LAUNCHP(k_enc_pack_feats, cdiv(static_cast<long>(p->Fmax) * Kp, 256), 256, 0, st, feats_dev, max_len, h->cfg.nfeats, Kp, p->frow_seq,
          p->frow_t, p->Rf, p->feats.act, pl);
  LinCall c;
  c.A = &p->feats; c.W = &h->skel; c.M_max = p->Fmax; c.M_dev = p->Rf; c.out = f32_only(p->emb, 256);
  CKS(launch_linear(h, st, mode, c));
  LAUNCHP(k_enc_init, cdiv(static_cast<long>(p->Rmax) * 256, 256), 256, 0, st, h->enc_gmt, p->emb, h->enc_pe, p->T, p->row_seq, p->row_t,
          p->mcnt, p->foff, p->R, p->x0.act, pl);
  const ActBuf* x = &p->x0;
  for (int i = 0; i < 4; ++i) {
    CKS(enqueue_enc_layer(h, p, st, i, *x, p->skip[i]));
    x = &p->skip[i];
  }
  CKS(enqueue_enc_layer(h, p, st, 4, *x, p->xa));
  for (int i = 0; i < 4; ++i) {
    c = LinCall();
    c.A = &p->xa; c.A2 = &p->skip[3 - i]; c.W = &h->enc_skip[i]; c.M_max = p->Rmax; c.M_dev = p->R;
    c.out = p->xb.act; c.out_planes = pl;
    CKS(launch_linear(h, st, mode, c));
    CKS(enqueue_enc_layer(h, p, st, 5 + i, p->xb, p->xa));
  }
  LAUNCHP(k_layernorm256, cdiv(static_cast<long>(p->Rmax) * 32, 256), 256, 0, st, p->xa.act.f32, 256, p->Rmax, p->R, h->enc_fg, h->enc_fb,
          f32_only(p->xn, 256), 0);
  return LADIFF_OK;
}

// ------------------------------------------------------------------------------------------------
// graph capture helper: records `body` once on the handle's private stream, replays on the caller's stream
template <typename Body>
int run_graphed(H* h, cudaStream_t st, cudaGraphExec_t* exec, int64_t* graph_launches, Body body) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  CK(cudaStreamIsCapturing(st, &cs));
  if (!h->cfg.use_cuda_graph || cs != cudaStreamCaptureStatusNone) return body(st);  // plain launches (also inside a caller's capture)
  if (!*exec) {
    if (!h->cap_stream) CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    CK(cudaStreamSynchronize(st));  // plan tables / workspace initialisation enqueued so far
    const int64_t before = h->launches;
    CK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int s = body(h->cap_stream);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &g);
    if (s != LADIFF_OK) {
      if (g) cudaGraphDestroy(g);
      return s;
    }
    if (e != cudaSuccess) return h->err.set(LADIFF_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    *graph_launches = h->launches - before;
    h->launches = before;
    e = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return h->err.set(LADIFF_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  }
  CK(cudaGraphLaunch(*exec, st));
  h->launches += *graph_launches;
  return LADIFF_OK;
}

template <int BN, int NS, int EPI>
cudaError_t set_tc_attr_e() {
  return cudaFuncSetAttribute(k_linear_tc<BN, NS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN, NS>::SMEM_BYTES);
}
template <int BN, int NS>
cudaError_t set_tc_attr() {
  cudaError_t e = set_tc_attr_e<BN, NS, EPI_BIAS>();
  if (e == cudaSuccess) e = set_tc_attr_e<BN, NS, EPI_RELU>();
  if (e == cudaSuccess) e = set_tc_attr_e<BN, NS, EPI_GELU>();
  if (e == cudaSuccess) e = set_tc_attr_e<BN, NS, EPI_RES>();
  if (e == cudaSuccess) e = set_tc_attr_e<BN, NS, EPI_SILU>();
  if (e == cudaSuccess && BN == 256) e = set_tc_attr_e<256, NS, EPI_LN>();
  if (e == cudaSuccess && BN == 256) e = set_tc_attr_e<256, NS, EPI_LN_MOD_SILU>();
  if (e == cudaSuccess && BN == 256)
    e = cudaFuncSetAttribute(k_linear_tc_ln<LN_CL, NS, EPI_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256 / LN_CL, NS>::SMEM_BYTES);
  if (e == cudaSuccess && BN == 256)
    e = cudaFuncSetAttribute(k_linear_tc_ln<LN_CL, NS, EPI_LN_MOD_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256 / LN_CL, NS>::SMEM_BYTES);
  return e;
}

// Plan caches are bounded: a plan owns its workspace and CUDA graph (the hoisted delta table alone is ~118 MB at B = 128 x 50
// steps), so a sweep over batch sizes / step counts must not grow device memory without limit.  Least recently used first.
constexpr size_t MAX_PLANS = 6;
template <class P>
void evict_lru(std::map<std::string, std::unique_ptr<P>>& m, const std::string& keep) {
  while (m.size() > MAX_PLANS) {
    auto victim = m.end();
    for (auto it = m.begin(); it != m.end(); ++it)
      if (it->first != keep && it->second && (victim == m.end() || it->second->last_use < victim->second->last_use)) victim = it;
    if (victim == m.end()) break;
    cudaDeviceSynchronize();   // the victim's graph / workspace may still be in flight on some stream
    m.erase(victim);
  }
}

int check_mode(H* h, int mode) {
  if (mode < 0 || mode > 2) return h->err.set(LADIFF_ERR_INVALID, "unknown mode %d", mode);
  return LADIFF_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int ladiff_abi_version(void) { return LADIFF_ABI_VERSION; }

const char* ladiff_last_error(const ladiff_handle* h) { return h ? h->err.msg.c_str() : g_create_error.c_str(); }

int64_t ladiff_last_launch_count(const ladiff_handle* h) { return h ? h->launches : 0; }

int ladiff_create(const ladiff_config* cfg, ladiff_handle** out) {
  if (!cfg || !out) {
    g_create_error = "null argument";
    return LADIFF_ERR_INVALID;
  }
  if (cfg->num_layers != 9 || cfg->latent_dim != 256 || cfg->num_heads != 4 || cfg->ff_size != 1024 || cfg->text_dim != 768 ||
      cfg->max_it < 1 || cfg->max_it > 8 || cfg->frame_per_latent < 1 || cfg->max_frames < 1 || cfg->max_frames > SA_MAXL ||
      cfg->nfeats < 4 || cfg->nfeats > 1024) {
    g_create_error = "unsupported configuration (kernels are specialised for 9 layers, width 256, 4 heads, ff 1024, text 768, "
                     "max_it <= 8, max_frames <= 256)";
    return LADIFF_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return LADIFF_ERR_CUDA;
  }
  std::unique_ptr<ladiff_handle> h(new ladiff_handle());
  h->cfg = *cfg;
  g_use_pdl = getenv("LADIFF_NO_PDL") == nullptr;
  cudaGetDevice(&h->device);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, h->device);
  if (e != cudaSuccess || prop.major != 10) {
    g_create_error = "device is not sm_100 (B200): the kernels are built for sm_100a only";
    return LADIFF_ERR_CUDA;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  if (e != cudaSuccess || !fn) {
    g_create_error = "cuTensorMapEncodeTiled not available from the driver";
    return LADIFF_ERR_CUDA;
  }
  h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  e = cudaFuncSetAttribute(k_attn_self, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelfAttnSmem));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_tc<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sat_smem_bytes<32>(2));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_tc<26, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sat_smem_bytes<26>(2));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_tc<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sat_smem_bytes<32>(1));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_tc<26, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sat_smem_bytes<26>(1));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cross_ln<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * CX_LD * (int)sizeof(float));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_cross_ln<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * CX_LD * (int)sizeof(float));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ffn_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FtCfg<2>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ffn_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FtCfg<1>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_t5<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, At5Cfg<2>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_t5<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, At5Cfg<2>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_t5<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, At5Cfg<1>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_attn_self_t5<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, At5Cfg<1>::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ffn_swap<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SwapCfg<2>::smem_bytes(48));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ffn_swap<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SwapCfg<1>::smem_bytes(48));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ffn_swap<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SwapCfg<2>::smem_bytes(48));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ffn_swap<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SwapCfg<1>::smem_bytes(48));
  if (e == cudaSuccess) e = set_tc_attr<256, 2>();
  if (e == cudaSuccess) e = set_tc_attr<256, 1>();
  if (e == cudaSuccess) e = set_tc_attr<128, 2>();
  if (e == cudaSuccess) e = set_tc_attr<128, 1>();
  if (e == cudaSuccess) e = set_tc_attr<64, 2>();
  if (e == cudaSuccess) e = set_tc_attr<64, 1>();
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
    return LADIFF_ERR_CUDA;
  }
  if (getenv("LADIFF_TRACE")) {
    h->trace_cap = 16384;
    if (cudaMalloc(&h->trace, 8ull * h->trace_cap * sizeof(unsigned long long)) != cudaSuccess) h->trace = nullptr;
    if (h->trace) cudaMemset(h->trace, 0xFF, 8ull * h->trace_cap * sizeof(unsigned long long));
  }
  *out = h.release();
  return LADIFF_OK;
}

int ladiff_trace_read(ladiff_handle* h, uint64_t* out_host, int32_t max_launches, char* names_host, int32_t name_stride) {
  if (!h || !h->trace || !out_host) return 0;
  cudaDeviceSynchronize();
  int n = static_cast<int>(h->trace_names.size());
  if (n > max_launches) n = max_launches;
  cudaMemcpy(out_host, h->trace, 8ull * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; ++i) {
    for (int s = 1; s < 8; ++s) out_host[8 * i + s] = ~out_host[8 * i + s];
    if (names_host && name_stride > 0) snprintf(names_host + static_cast<size_t>(i) * name_stride, name_stride, "%s", h->trace_names[i].c_str());
  }
  cudaMemset(h->trace, 0xFF, 8ull * h->trace_cap * sizeof(unsigned long long));
  return n;
}

void ladiff_destroy(ladiff_handle* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  h->den_plans.clear();
  h->rev_plans.clear();
  h->dec_plans.clear();
  h->enc_plans.clear();
  for (auto& kv : h->raw) cudaFree(kv.second.dev);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (int c = 0; c < MAX_CHAINS; ++c) {
    if (h->aux[c]) cudaStreamDestroy(h->aux[c]);
    if (h->ev_aux_fork[c]) cudaEventDestroy(h->ev_aux_fork[c]);
    if (h->ev_aux_join[c]) cudaEventDestroy(h->ev_aux_join[c]);
    if (h->side[c]) cudaStreamDestroy(h->side[c]);
    if (h->ev_join[c]) cudaEventDestroy(h->ev_join[c]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  delete h;
}

int ladiff_set_weight(ladiff_handle* h, const char* name, const float* data_dev, const int64_t* shape, int32_t ndim, void* stream) {
  if (!h || !name || !data_dev || !shape || ndim < 1 || ndim > 4) return h ? h->err.set(LADIFF_ERR_INVALID, "bad argument") : LADIFF_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Raw r;
  r.shape.assign(shape, shape + ndim);
  auto it = h->raw.find(name);
  if (it != h->raw.end()) {
    cudaFree(it->second.dev);
    h->raw.erase(it);
  }
  CK(cudaMalloc(&r.dev, r.numel() * sizeof(float)));
  CK(cudaMemcpyAsync(r.dev, data_dev, r.numel() * sizeof(float), cudaMemcpyDeviceToDevice, st));
  h->raw[name] = r;
  if (strncmp(name, "denoiser.", 9) == 0) h->den_ready = false;
  else if (strncmp(name, "vae.", 4) == 0) h->dec_ready = false;
  else h->den_ready = h->dec_ready = false;
  return LADIFF_OK;
}

int ladiff_finalize_weights(ladiff_handle* h, int32_t which, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(cudaDeviceSynchronize());  // plans may still be in flight
  if (which & 1) {
    h->den_plans.clear();
    h->rev_plans.clear();
    CKS(finalize_denoiser(h, st));
  }
  if (which & 2) {
    h->dec_plans.clear();
    h->enc_plans.clear();
    CKS(finalize_decoder(h, st));
  }
  CK(cudaStreamSynchronize(st));
  return LADIFF_OK;
}

int ladiff_diffusion_reverse(ladiff_handle* h, const float* text_emb_dev, const int32_t* lengths_host, int32_t B,
                             const float* noise_dev, int32_t n_steps, const int32_t* timesteps_host, const float* c1_host,
                             const float* c2_host, float guidance_scale, int32_t mode, float* z_out_dev, void* stream) {
  return ladiff_diffusion_reverse_ex(h, text_emb_dev, lengths_host, nullptr, B, noise_dev, n_steps, timesteps_host, c1_host, c2_host,
                                     nullptr, nullptr, 0ull, 0, guidance_scale, mode, z_out_dev, stream);
}

int ladiff_diffusion_reverse_ex(ladiff_handle* h, const float* text_emb_dev, const int32_t* lengths_host,
                                const int32_t* rows_host, int32_t B, const float* noise_dev, int32_t n_steps,
                                const int32_t* timesteps_host, const float* c1_host, const float* c2_host, const float* c3_host,
                                const float* step_noise_dev, uint64_t seed, int32_t flags, float guidance_scale, int32_t mode,
                                float* z_out_dev, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  h->trace_n = 0;
  if (!h->den_ready) return h->err.set(LADIFF_ERR_STATE, "denoiser weights not finalised");
  CKS(check_mode(h, mode));
  if (!text_emb_dev || (!lengths_host && !rows_host) || !noise_dev || !timesteps_host || !c1_host || !c2_host || !z_out_dev || B < 1 ||
      n_steps < 1 || (flags & ~1))
    return h->err.set(LADIFF_ERR_INVALID, "ladiff_diffusion_reverse: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = h->cfg.max_it;
  const int ar_flag = flags & LADIFF_REVERSE_AR;
  char key[96];
  snprintf(key, sizeof(key), "cfg:%d:%d:%d:%d", B, n_steps, mode, ar_flag);   // guidance / coefficients / seed live in device tables
  auto& slot = h->rev_plans[key];
  if (!slot) {
    slot.reset(new ReversePlan());
    int s = build_reverse_plan(h, slot.get(), B, n_steps, mode, ar_flag);
    if (s != LADIFF_OK) {
      h->rev_plans.erase(key);
      return s;
    }
    evict_lru(h->rev_plans, key);
  }
  ReversePlan* rp = h->rev_plans[key].get();
  rp->last_use = ++h->use_clock;
  // ---- per-call inputs staged at fixed addresses (the captured graph reads them); the ragged row layout is derived on
  // the device inside the graph
  std::vector<int> cnt(B);
  for (int b = 0; b < B; ++b) {
    int m;
    if (rows_host) {   // explicit latent rows per sequence (ARDIFF: 1 + number of context latents, no length mask)
      m = rows_host[b];
      if (m < 1 || m > T) return h->err.set(LADIFF_ERR_INVALID, "rows[%d] = %d outside [1, %d]", b, m, T);
    } else {
      if (lengths_host[b] < 1) return h->err.set(LADIFF_ERR_INVALID, "lengths[%d] = %d", b, lengths_host[b]);
      m = (lengths_host[b] + h->cfg.frame_per_latent - 1) / h->cfg.frame_per_latent;
    }
    cnt[b] = m < T ? m : T;
  }
  CK(cudaMemcpyAsync(rp->cnt, cnt.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(rp->text768, text_emb_dev, static_cast<size_t>(2 * B) * 768 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(rp->lat, noise_dev, static_cast<size_t>(B) * T * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(rp->noise_pp, &step_noise_dev, sizeof(float*), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(rp->seed_p, &seed, sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  // ---- schedule-dependent tables (cached; shared by all chains)
  std::vector<int> ts(timesteps_host, timesteps_host + n_steps);
  std::vector<float> coef(4 * n_steps);
  for (int i = 0; i < n_steps; ++i) {
    coef[4 * i] = c1_host[i];
    coef[4 * i + 1] = c2_host[i];
    coef[4 * i + 2] = c3_host ? c3_host[i] : 0.f;
    coef[4 * i + 3] = guidance_scale;
  }
  DenoisePlan* p0 = rp->chains[0].get();
  if (coef != rp->cached_coef) {   // stream-ordered: earlier replays on this stream have read the old values by then
    CK(cudaMemcpyAsync(p0->coef, coef.data(), coef.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    rp->cached_coef = coef;
  }
  if (ts != rp->cached_ts) {
    CK(cudaDeviceSynchronize());  // previous users of the time tables (any stream)
    CK(cudaMemcpyAsync(p0->ts, ts.data(), n_steps * sizeof(int), cudaMemcpyHostToDevice, st));
    CKS(enqueue_time_tables(h, p0, st));
    CK(cudaStreamSynchronize(st));
    rp->cached_ts = ts;
  }
  CKS(run_graphed(h, st, &rp->exec, &rp->graph_launches, [&](cudaStream_t s) { return enqueue_reverse_chains(h, rp, s); }));
  LAUNCH(k_z_out, cdiv(static_cast<long>(B) * T * 256, 256), 256, 0, st, rp->lat, rp->cnt, B, T, z_out_dev);
  return LADIFF_OK;
}

int ladiff_denoiser_forward(ladiff_handle* h, const float* sample_dev, int32_t timestep, const float* text_emb_dev,
                            const int32_t* max_iter_elements_host, int32_t S, int32_t mode, float* out_dev, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  if (!h->den_ready) return h->err.set(LADIFF_ERR_STATE, "denoiser weights not finalised");
  CKS(check_mode(h, mode));
  if (!sample_dev || !text_emb_dev || !max_iter_elements_host || !out_dev || S < 1)
    return h->err.set(LADIFF_ERR_INVALID, "ladiff_denoiser_forward: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = h->cfg.max_it;
  char key[96];
  snprintf(key, sizeof(key), "fwd:%d:%d", S, mode);
  auto& slot = h->den_plans[key];
  if (!slot) {
    slot.reset(new DenoisePlan());
    int s = build_denoise_plan(h, slot.get(), S, 1, mode, false);
    if (s != LADIFF_OK) {
      h->den_plans.erase(key);
      return s;
    }
    evict_lru(h->den_plans, key);
  }
  DenoisePlan* p = h->den_plans[key].get();
  p->last_use = ++h->use_clock;
  std::vector<int> cnt(S);
  for (int s = 0; s < S; ++s) {
    if (max_iter_elements_host[s] < 1) return h->err.set(LADIFF_ERR_INVALID, "max_iter_elements[%d] = %d", s, max_iter_elements_host[s]);
    cnt[s] = max_iter_elements_host[s] < T ? max_iter_elements_host[s] : T;
  }
  CKS(enqueue_meta(h, st, cnt.data(), S, p->cnt, p->off, p->R, p->row_seq, p->row_t, nullptr, 0));
  CK(cudaMemcpyAsync(p->text768, text_emb_dev, static_cast<size_t>(S) * 768 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(p->ts, &timestep, sizeof(int), cudaMemcpyHostToDevice, st));
  CKS(enqueue_time_tables(h, p, st));
  CKS(enqueue_text_tables(h, p, st));
  LAUNCHP(k_pack_x, cdiv(static_cast<long>(p->Rmax) * 256, 256), 256, 0, st, sample_dev, S, T, h->den_pe, p->row_seq, p->row_t, p->R,
         p->xin.act, p->planes);
  CKS(enqueue_den_tokens(h, p, st, 0));
  LAUNCH(k_final_ln_out, cdiv(static_cast<long>(S) * T * 32, 256), 256, 0, st, p->xa.act.f32, p->off, S, T, h->den_fg, h->den_fb, out_dev);
  return LADIFF_OK;
}

int ladiff_cfg_ddim_step(ladiff_handle* h, const float* noise_pred_dev, float* latents_dev, int32_t B, float guidance_scale,
                         float c1, float c2, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  if (!noise_pred_dev || !latents_dev || B < 1) return h->err.set(LADIFF_ERR_INVALID, "ladiff_cfg_ddim_step: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long n_half = static_cast<long>(B) * h->cfg.max_it * 256;
  LAUNCH(k_cfg_ddim_dense, cdiv(n_half / 4, 256), 256, 0, st, noise_pred_dev, latents_dev, n_half, guidance_scale, c1, c2);
  return LADIFF_OK;
}

int ladiff_vae_decode(ladiff_handle* h, const float* z_dev, const int32_t* lengths_host, int32_t B, int32_t max_len,
                      int32_t mode, float* out_dev, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  if (!h->dec_ready) return h->err.set(LADIFF_ERR_STATE, "vae decoder weights not finalised");
  CKS(check_mode(h, mode));
  if (!z_dev || !lengths_host || !out_dev || B < 1) return h->err.set(LADIFF_ERR_INVALID, "ladiff_vae_decode: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int T = h->cfg.max_it, nf = h->cfg.nfeats;
  std::vector<int> cnt(B), mcnt(B);
  for (int b = 0; b < B; ++b) {
    const int L = lengths_host[b];
    if (L < 1 || L > h->cfg.max_frames || L > max_len)
      return h->err.set(LADIFF_ERR_INVALID, "lengths[%d] = %d outside [1, min(max_frames=%d, max_len=%d)]", b, L, h->cfg.max_frames, max_len);
    cnt[b] = L;
    int m = (L + h->cfg.frame_per_latent - 1) / h->cfg.frame_per_latent;
    mcnt[b] = m < T ? m : T;
  }
  char key[96];
  snprintf(key, sizeof(key), "dec:%d:%d", B, mode);
  auto& slot = h->dec_plans[key];
  if (!slot) {
    slot.reset(new DecodePlan());
    int s = build_decode_plan(h, slot.get(), B, mode);
    if (s != LADIFF_OK) {
      h->dec_plans.erase(key);
      return s;
    }
    evict_lru(h->dec_plans, key);
  }
  DecodePlan* p = h->dec_plans[key].get();
  p->last_use = ++h->use_clock;
  CKS(enqueue_meta(h, st, cnt.data(), B, p->cnt, p->foff, p->Rf, p->frow_seq, p->frow_t, p->frow_dst, max_len));
  CKS(enqueue_meta(h, st, mcnt.data(), B, p->mcnt, p->moff, p->Rm, p->mrow_seq, p->mrow_t, nullptr, 0));
  CK(cudaMemcpyAsync(p->z, z_dev, static_cast<size_t>(T) * B * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemsetAsync(out_dev, 0, static_cast<size_t>(B) * max_len * nf * sizeof(float), st));  // padded frames: exact zeros
  CKS(run_graphed(h, st, &p->exec, &p->graph_launches, [&](cudaStream_t s) { return enqueue_decode_body(h, p, s); }));
  // final_layer (256 -> nfeats) scattering rows straight into [B, max_len, nfeats]; outside the graph: per-call pointer
  LinCall c;
  c.A = &p->xn; c.W = &h->dec_final; c.M_max = p->Rmax; c.M_dev = p->Rf; c.row_map = p->frow_dst;
  c.out = f32_only(out_dev, nf);
  CKS(launch_linear(h, st, mode, c));
  return LADIFF_OK;
}

int ladiff_vae_encode(ladiff_handle* h, const float* feats_dev, const int32_t* lengths_host, int32_t B, int32_t max_len,
                      const float* eps_dev, int32_t mode, float* latent_dev, float* mu_dev, float* std_dev, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  if (!h->dec_ready || !h->enc_ready) return h->err.set(LADIFF_ERR_STATE, "vae encoder weights not finalised (vae.encoder.* / vae.skel_embedding.* missing?)");
  CKS(check_mode(h, mode));
  if (!feats_dev || !lengths_host || B < 1 || max_len < 1 || (!latent_dev && !mu_dev && !std_dev))
    return h->err.set(LADIFF_ERR_INVALID, "ladiff_vae_encode: bad argument");
  const int T = h->cfg.max_it;
  if (h->cfg.max_frames + 2 * T > SA_MAXL) return h->err.set(LADIFF_ERR_INVALID, "encode: max_frames + 2 max_it must be <= %d", SA_MAXL);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<int> cnt(B), fcnt(B), mcnt(B);
  for (int b = 0; b < B; ++b) {
    const int L = lengths_host[b];
    if (L < 1 || L > h->cfg.max_frames || L > max_len)
      return h->err.set(LADIFF_ERR_INVALID, "lengths[%d] = %d outside [1, min(max_frames=%d, max_len=%d)]", b, L, h->cfg.max_frames, max_len);
    int m = (L + h->cfg.frame_per_latent - 1) / h->cfg.frame_per_latent;
    m = m < T ? m : T;
    mcnt[b] = m;
    fcnt[b] = L;
    cnt[b] = 2 * m + L;
  }
  char key[96];
  snprintf(key, sizeof(key), "enc:%d:%d", B, mode);
  auto& slot = h->enc_plans[key];
  if (!slot) {
    slot.reset(new EncodePlan());
    int s = build_encode_plan(h, slot.get(), B, mode);
    if (s != LADIFF_OK) {
      h->enc_plans.erase(key);
      return s;
    }
    evict_lru(h->enc_plans, key);
  }
  EncodePlan* p = h->enc_plans[key].get();
  p->last_use = ++h->use_clock;
  CKS(enqueue_meta(h, st, cnt.data(), B, p->cnt, p->off, p->R, p->row_seq, p->row_t, nullptr, 0));
  CKS(enqueue_meta(h, st, fcnt.data(), B, p->fcnt, p->foff, p->Rf, p->frow_seq, p->frow_t, nullptr, 0));
  CK(cudaMemcpyAsync(p->mcnt, mcnt.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  g_skip_pdl_once = true;   // first kernel after plain copies
  CKS(enqueue_encode_body(h, p, st, feats_dev, max_len));
  LAUNCH(k_enc_out, cdiv(static_cast<long>(T) * B * 256, 256), 256, 0, st, p->xn, p->off, p->mcnt, B, T, eps_dev, latent_dev, mu_dev, std_dev);
  return LADIFF_OK;
}

int ladiff_feats2joints(ladiff_handle* h, const float* feats_dev, const float* mean_dev, const float* std_dev, int32_t B,
                        int32_t max_len, int32_t njoints, float* joints_dev, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  if (!feats_dev || !mean_dev || !std_dev || !joints_dev || B < 1 || max_len < 1 || max_len > SA_MAXL || njoints < 2 ||
      (njoints - 1) * 3 + 4 > h->cfg.nfeats)
    return h->err.set(LADIFF_ERR_INVALID, "ladiff_feats2joints: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LAUNCH(k_feats2joints, B, 256, 0, st, feats_dev, mean_dev, std_dev, max_len, h->cfg.nfeats, njoints, joints_dev);
  return LADIFF_OK;
}

int ladiff_linear_test(ladiff_handle* h, const float* A_dev, const float* W_dev, const float* bias_dev, const float* res_dev,
                       const float* ln_g_dev, const float* ln_b_dev, const float* mod_dev, int32_t M, int32_t N, int32_t K,
                       int32_t epilogue, int32_t mode, float* out_dev, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  CKS(check_mode(h, mode));
  if (!A_dev || !W_dev || !out_dev || M < 1 || N < 1 || K < 64 || K % 64) return h->err.set(LADIFF_ERR_INVALID, "ladiff_linear_test: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena ar;
  Weight w;
  CKS(pack_weight(h, ar, st, &w, W_dev, K, N, K, bias_dev));
  const int planes = mode == LADIFF_MODE_FP32 ? 0 : (mode == LADIFF_MODE_BF16X3 ? 2 : 1);
  ActBuf a;
  CKS(alloc_act(h, ar, &a, M, K, true, planes > 0));
  LAUNCHP(k_unary, cdiv(static_cast<long>(M) * K, 256), 256, 0, st, A_dev, K, M, K, (int)U_COPY, a.act, planes);
  LinCall c;
  c.A = &a; c.W = &w; c.M_max = M; c.epi = epilogue; c.res = res_dev; c.ldres = N; c.ln_g = ln_g_dev; c.ln_b = ln_b_dev; c.mod = mod_dev;
  c.out = f32_only(out_dev, N);
  CKS(launch_linear(h, st, mode, c));
  CK(cudaStreamSynchronize(st));
  return LADIFF_OK;
}

// Test / measurement hook: the two feed-forward pairs of denoiser layer `layer` on M rows of x, either as ONE cluster
// kernel (fused = 1) or as the four separate fused linears (fused = 0).  iters > 0 additionally times back-to-back launches.
int ladiff_ffn_test(ladiff_handle* h, const float* x_dev, int32_t M, int32_t layer, const float* mod_dev, int32_t mode,
                    int32_t fused, int32_t iters, float* x3_out_dev, float* s_out_dev, float* ms_per_call_host, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  CKS(check_mode(h, mode));
  if (!h->den_ready) return h->err.set(LADIFF_ERR_STATE, "denoiser weights not finalised");
  if (!x_dev || !mod_dev || !x3_out_dev || !s_out_dev || M < 1 || layer < 0 || layer >= NL)
    return h->err.set(LADIFF_ERR_INVALID, "ladiff_ffn_test: bad argument");
  if (fused && mode == LADIFF_MODE_FP32) return h->err.set(LADIFF_ERR_INVALID, "ladiff_ffn_test: the fused kernel is a tensor-core path");
  const bool tile_path = fused && ffn_swap_rt(M) == 0;   // above the co-resident capacity the plans run the persistent tile kernel per pair
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DenLayerW& w = h->den[layer];
  const int planes = mode == LADIFF_MODE_FP32 ? 0 : (mode == LADIFF_MODE_BF16X3 ? 2 : 1);
  const bool tcm = planes > 0;
  Arena ar;
  ActBuf x, x3, sb, hb;
  CKS(alloc_act(h, ar, &x, M, 256, true, tcm));
  CKS(alloc_act(h, ar, &x3, M, 256, true, tcm));
  CKS(alloc_act(h, ar, &sb, M, 256, true, tcm));
  CKS(alloc_act(h, ar, &hb, M, 1024, !tcm, tcm));
  LAUNCHP(k_unary, cdiv(static_cast<long>(M) * 256, 256), 256, 0, st, x_dev, 256, M, 256, (int)U_COPY, x.act, planes);
  auto run = [&]() -> int {
    if (tile_path) {
      FtCall f;
      f.X = &x; f.W1 = &w.ff1; f.W2 = &w.ff2; f.M_max = M; f.act = EPI_RELU; f.kind = EPI_LN; f.res = x.act.f32;
      f.ln_g = w.n2g; f.ln_b = w.n2b; f.out = x3.act; f.out_planes = planes;
      CKS(launch_ffn_tile(h, st, mode, f));
      f = FtCall();
      f.X = &x3; f.W1 = &w.gff1; f.W2 = &w.gff2; f.M_max = M; f.act = EPI_GELU; f.kind = EPI_LN_MOD_SILU;
      f.ln_g = w.ffn_sn_g; f.ln_b = w.ffn_sn_b; f.mod = mod_dev; f.out = sb.act; f.out_planes = planes;
      return launch_ffn_tile(h, st, mode, f);
    }
    if (fused) {
      FfnCall f;
      f.X = &x; f.M_max = M; f.npairs = 2; f.out_planes = planes; f.res = x.act.f32;
      f.pair[0].W1 = &w.ff1; f.pair[0].W2 = &w.ff2; f.pair[0].act = EPI_RELU; f.pair[0].kind = EPI_LN;
      f.pair[0].ln_g = w.n2g; f.pair[0].ln_b = w.n2b; f.pair[0].out = x3.act;
      f.pair[1].W1 = &w.gff1; f.pair[1].W2 = &w.gff2; f.pair[1].act = EPI_GELU; f.pair[1].kind = EPI_LN_MOD_SILU;
      f.pair[1].ln_g = w.ffn_sn_g; f.pair[1].ln_b = w.ffn_sn_b; f.pair[1].mod = mod_dev; f.pair[1].out = sb.act;
      f.dbg = g_ffn_dbg;
      return launch_ffn_swap(h, st, mode, f);
    }
    LinCall c;
    c.A = &x; c.W = &w.ff1; c.M_max = M; c.epi = EPI_RELU; c.out = hb.act; c.out_planes = planes;
    CKS(launch_linear(h, st, mode, c));
    c = LinCall(); c.A = &hb; c.W = &w.ff2; c.M_max = M; c.epi = EPI_LN; c.res = x.act.f32; c.ln_g = w.n2g; c.ln_b = w.n2b;
    c.out = x3.act; c.out_planes = planes;
    CKS(launch_linear(h, st, mode, c));
    c = LinCall(); c.A = &x3; c.W = &w.gff1; c.M_max = M; c.epi = EPI_GELU; c.out = hb.act; c.out_planes = planes;
    CKS(launch_linear(h, st, mode, c));
    c = LinCall(); c.A = &hb; c.W = &w.gff2; c.M_max = M; c.epi = EPI_LN_MOD_SILU; c.ln_g = w.ffn_sn_g; c.ln_b = w.ffn_sn_b; c.mod = mod_dev;
    c.out = sb.act; c.out_planes = planes;
    return launch_linear(h, st, mode, c);
  };
  CKS(run());
  CK(cudaMemcpyAsync(x3_out_dev, x3.act.f32, static_cast<size_t>(M) * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(s_out_dev, sb.act.f32, static_cast<size_t>(M) * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (iters > 0 && ms_per_call_host) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) CKS(run());
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_call_host = ms / iters;
  }
  CK(cudaStreamSynchronize(st));
  if (fused && !tile_path && getenv("LADIFF_DBG_STAMPS")) {
    long long* dbg = nullptr;
    const int rt = ffn_swap_rt(M);
    const int ncta = ((M + rt - 1) / rt) * 4;
    const int dstride = 128;
    CK(ar.alloc((void**)&dbg, ncta * dstride * sizeof(long long)));
    CK(cudaMemsetAsync(dbg, 0, ncta * dstride * sizeof(long long), st));
    g_ffn_dbg = dbg;
    CKS(run());
    g_ffn_dbg = nullptr;
    CK(cudaStreamSynchronize(st));
    std::vector<long long> hb(ncta * dstride);
    CK(cudaMemcpy(hb.data(), dbg, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    // clock64 stamps relative to the CTA start; slot meaning: see the FSTAMP / SSTAMP sites of the kernel (+20 per pair)
    for (int cta : {0, 3, ncta - 1}) {
      const long long* d = hb.data() + cta * dstride;
      fprintf(stderr, "  swap cta %3d:", cta);
      for (int i = 1; i < dstride; ++i)
        if (d[i]) fprintf(stderr, " [%d]%lld", i, d[i] - d[0]);
      fprintf(stderr, "\n");
    }
  }
  return LADIFF_OK;
}

// Times `iters` back-to-back launches of one fused linear (same kernels as the plans) with CUDA events on `stream`.
int ladiff_linear_bench(ladiff_handle* h, int32_t M, int32_t N, int32_t K, int32_t epilogue, int32_t mode, int32_t iters,
                        float* ms_per_launch_host, void* stream) {
  if (!h) return LADIFF_ERR_INVALID;
  h->launches = 0;
  CKS(check_mode(h, mode));
  if (!ms_per_launch_host || M < 1 || N < 1 || K < 64 || K % 64 || iters < 1) return h->err.set(LADIFF_ERR_INVALID, "ladiff_linear_bench: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena ar;
  float *Wsrc = nullptr, *Asrc = nullptr, *vec = nullptr, *res = nullptr;
  CK(ar.alloc((void**)&Wsrc, static_cast<size_t>(N) * K * sizeof(float)));
  CK(ar.alloc((void**)&Asrc, static_cast<size_t>(M) * K * sizeof(float)));
  CK(ar.alloc((void**)&vec, 1024 * sizeof(float)));
  CK(ar.alloc((void**)&res, static_cast<size_t>(M) * N * sizeof(float)));
  LAUNCH(k_fill_pseudo, cdiv(static_cast<long>(N) * K, 256), 256, 0, st, Wsrc, static_cast<long>(N) * K, 0.05f, 1u);
  LAUNCH(k_fill_pseudo, cdiv(static_cast<long>(M) * K, 256), 256, 0, st, Asrc, static_cast<long>(M) * K, 1.0f, 2u);
  LAUNCH(k_fill_pseudo, 4, 256, 0, st, vec, 1024L, 0.1f, 3u);
  LAUNCH(k_fill_pseudo, cdiv(static_cast<long>(M) * N, 256), 256, 0, st, res, static_cast<long>(M) * N, 1.0f, 4u);
  Weight w;
  CKS(pack_weight(h, ar, st, &w, Wsrc, K, N, K, vec));
  const int planes = mode == LADIFF_MODE_FP32 ? 0 : (mode == LADIFF_MODE_BF16X3 ? 2 : 1);
  ActBuf a, o;
  CKS(alloc_act(h, ar, &a, M, K, true, planes > 0));
  CKS(alloc_act(h, ar, &o, M, N, true, planes > 0));
  LAUNCHP(k_unary, cdiv(static_cast<long>(M) * K, 256), 256, 0, st, Asrc, K, M, K, (int)U_COPY, a.act, planes);
  LinCall c;
  c.A = &a; c.W = &w; c.M_max = M; c.epi = epilogue; c.res = res; c.ldres = N; c.ln_g = vec; c.ln_b = vec + 256; c.mod = vec + 512;
  c.out = o.act; c.out_planes = planes;
  for (int i = 0; i < 3; ++i) CKS(launch_linear(h, st, mode, c));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < iters; ++i) CKS(launch_linear(h, st, mode, c));
  CK(cudaEventRecord(e1, st));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch_host = ms / iters;
  if (getenv("LADIFF_DBG_STAMPS") && mode != LADIFF_MODE_FP32) {
    long long* dbg = nullptr;
    const int ncta = 4096;
    CK(ar.alloc((void**)&dbg, ncta * 16 * sizeof(long long)));
    CK(cudaMemsetAsync(dbg, 0, ncta * 16 * sizeof(long long), st));
    c.dbg = dbg;
    CKS(launch_linear(h, st, mode, c));
    CK(cudaStreamSynchronize(st));
    std::vector<long long> hb(ncta * 16);
    CK(cudaMemcpy(hb.data(), dbg, hb.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    for (int cta : {0, 9}) {
      const long long* d = hb.data() + cta * 16;
      if (!d[0]) continue;
      fprintf(stderr, "  cta %3d: setup %lld | tma0-issued %lld | tma-all %lld | first-full %lld | last-full %lld | mma-done-issue %lld | accum-ready %lld | epi-done %lld | e10 %lld e11 %lld e12 %lld e13 %lld\n",
              cta, d[1] - d[0], d[2] - d[0], d[3] - d[0], d[4] - d[0], d[5] - d[0], d[6] - d[0], d[7] - d[0], d[8] - d[0],
              d[10] - d[0], d[11] - d[0], d[12] - d[0], d[13] - d[0]);
    }
  }
  return LADIFF_OK;
}

}  // extern "C"
