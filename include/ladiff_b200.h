/*
 * ladiff_b200.h -- C ABI of the B200-native LADiff sampling hot path.
 *
 * The reference (AlessioSam/LADiff) is pure Python/PyTorch and has no FFI; its
 * plugin boundary is `instantiate_from_config` (src/ladiff/config.py:26-33), i.e.
 * `target: pkg.mod.Class` strings in the YAML configs.  The Python classes in
 * `ladiff_b200/` mirror those targets and call ONLY the entry points below
 * (ctypes).  Each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross the boundary;
 *   - `*_dev` pointers are device memory owned by the caller and must stay valid
 *     until the work enqueued on `stream` has finished; `*_host` pointers are host
 *     memory, consumed before the call returns;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*);
 *     no call synchronises the device, except plan building on first use of a new
 *     (batch, steps, mode) shape, which allocates workspace and captures a CUDA graph;
 *   - one host thread per handle; handles are independent (no global state);
 *   - return value 0 = ok, negative = ladiff_status; text via ladiff_last_error().
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with LADIFF_ERR_CUDA.
 */
#ifndef LADIFF_B200_H
#define LADIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LADIFF_ABI_VERSION 1

#if defined(__GNUC__)
#define LADIFF_API __attribute__((visibility("default")))
#else
#define LADIFF_API
#endif

typedef struct ladiff_handle ladiff_handle;

typedef enum {
  LADIFF_OK = 0,
  LADIFF_ERR_INVALID = -1,   /* bad argument / unsupported configuration (Python raises ValueError) */
  LADIFF_ERR_WEIGHTS = -2,   /* missing / mis-shaped weight (Python raises KeyError / RuntimeError like load_state_dict) */
  LADIFF_ERR_CUDA = -3,      /* CUDA runtime / driver failure */
  LADIFF_ERR_STATE = -4      /* call order (weights not finalised, ...) */
} ladiff_status;

/* Arithmetic used by the GEMMs (everything else is fp32):
 *   FP32   : SIMT FFMA, exact fp32 products                     (parity path, slow)
 *   BF16X3 : tcgen05 bf16 MMA on hi/lo split operands, 3 products, fp32 TMEM accumulate
 *            (fp32-grade: decoded features within 1e-3 of the fp32 reference)
 *   BF16   : tcgen05 bf16 MMA, fp32 accumulate                  (fast path, measured tolerance in DESIGN.md) */
typedef enum { LADIFF_MODE_FP32 = 0, LADIFF_MODE_BF16X3 = 1, LADIFF_MODE_BF16 = 2 } ladiff_mode;

/* Mirrors the constructor parameters that reach the hot path:
 * configs/modules/denoiser.yaml:3-22, configs/modules/motion_vae.yaml:3-14,
 * TRAIN.ABLATION.{MAX_IT,FRAME_PER_LATENT} (configs/config_ladiff_humanml3d.yaml:58-59). */
typedef struct {
  int32_t nfeats;            /* 263 HumanML3D / 251 KIT-ML */
  int32_t num_layers;        /* 9  (must be 9)   */
  int32_t latent_dim;        /* 256 (must be 256) */
  int32_t num_heads;         /* 4  (must be 4)   */
  int32_t ff_size;           /* 1024 (must be 1024) */
  int32_t text_dim;          /* 768 (must be 768) */
  int32_t max_it;            /* 5: latent slots T (<= 8) */
  int32_t frame_per_latent;  /* 48 */
  int32_t max_frames;        /* 196 (<= 256) */
  int32_t use_cuda_graph;    /* 1: replay the captured loop; 0: launch kernel by kernel (debug / profiling) */
} ladiff_config;

LADIFF_API int  ladiff_abi_version(void);
/* Creates a handle on the current CUDA device. */
LADIFF_API int  ladiff_create(const ladiff_config* cfg, ladiff_handle** out);
LADIFF_API void ladiff_destroy(ladiff_handle* h);
/* Last error text of this handle (or of the failed ladiff_create when h == NULL). */
LADIFF_API const char* ladiff_last_error(const ladiff_handle* h);

/* Weights: replaces `load_state_dict` of LADiffDenoiser / LADiffVae (demo.py:138-159).
 * `name` is the reference state_dict key ("denoiser.encoder.input_blocks.0.sa_block.self_attn.in_proj_weight",
 * "vae.final_layer.weight", ...), data fp32 row-major on the device.  The library copies / repacks
 * (bf16 hi/lo planes, K-major, concatenated tables); the caller's tensor may be freed after the call. */
LADIFF_API int ladiff_set_weight(ladiff_handle* h, const char* name, const float* data_dev,
                      const int64_t* shape, int32_t ndim, void* stream);
/* Verifies that every key needed by `which` (1 = denoiser, 2 = vae decoder, 3 = both) is present and packs. */
LADIFF_API int ladiff_finalize_weights(ladiff_handle* h, int32_t which, void* stream);

/* LADIFF._diffusion_reverse, LAD branch (models/modeltype/ladiff.py:333-571): the whole
 * n_steps x (CFG-doubled denoiser -> CFG combine -> scheduler.step) loop.
 *   text_emb_dev  [2B,768]  rows 0..B-1 uncond, B..2B-1 cond (ladiff.py:258-264)
 *   lengths_host  [B]       frames; m_i = ceil(L_i / frame_per_latent)
 *   noise_dev     [B,T,256] initial randn (ladiff.py:380-385 draws it; here injected)
 *   timesteps_host[n_steps] scheduler.timesteps; c1/c2: per-step x' = c1*x + c2*eps
 *                           (DDIM eta=0 closed form, see ladiff_b200/scheduler.py)
 *   z_out_dev     [T,B,256] rows >= m_i exactly zero (ladiff.py:500,562-566) */
LADIFF_API int ladiff_diffusion_reverse(ladiff_handle* h, const float* text_emb_dev, const int32_t* lengths_host,
                             int32_t B, const float* noise_dev, int32_t n_steps,
                             const int32_t* timesteps_host, const float* c1_host, const float* c2_host,
                             float guidance_scale, int32_t mode, float* z_out_dev, void* stream);

/* Extended form of the loop for the reference's other scheduler / branch on this path:
 *   - DDPM sampling (configs/modules/scheduler.yaml:16-29, diffusers DDPMScheduler.step): x' = c1*x + c2*eps + c3*noise,
 *     c3_host[n_steps] = sqrt(variance_t) (NULL = all zero).  noise: step_noise_dev [n_steps,B,T,256] when injected
 *     (parity tests; the reference draws torch.randn inside scheduler.step), else an in-kernel Philox4x32-10 stream
 *     that is a pure function of (seed, step, element);
 *   - ARDIFF autoregressive branch (ladiff.py:419-467): flags & LADIFF_REVERSE_AR -> only latent slot 0 of every
 *     sequence is denoised, slots 1..rows-1 of noise_dev are fixed context latents (enclat, ladiff_denoiser.py:247-248),
 *     rows_host[B] = 1 + number of context latents (no length mask on this branch: lengths_host may be NULL).
 *   rows_host (optional) overrides m_i = ceil(L_i / frame_per_latent). */
#define LADIFF_REVERSE_AR 1
LADIFF_API int ladiff_diffusion_reverse_ex(ladiff_handle* h, const float* text_emb_dev, const int32_t* lengths_host,
                                const int32_t* rows_host, int32_t B, const float* noise_dev, int32_t n_steps,
                                const int32_t* timesteps_host, const float* c1_host, const float* c2_host,
                                const float* c3_host, const float* step_noise_dev, uint64_t seed, int32_t flags,
                                float guidance_scale, int32_t mode, float* z_out_dev, void* stream);

/* LADiffDenoiser.forward (models/architectures/ladiff_denoiser.py:153-295), one call.
 *   sample_dev [S,T,256], timestep (integer), text_emb_dev [S,768], max_iter_elements_host [S]
 *   out_dev    [S,T,256]; rows t >= max_iter_elements[s] are written as 0 (the reference leaves
 *              values there that never reach a valid row and are re-zeroed at ladiff.py:562-566). */
LADIFF_API int ladiff_denoiser_forward(ladiff_handle* h, const float* sample_dev, int32_t timestep,
                            const float* text_emb_dev, const int32_t* max_iter_elements_host,
                            int32_t S, int32_t mode, float* out_dev, void* stream);

/* CFG combine + scheduler.step (ladiff.py:487-492) as ONE elementwise kernel:
 *   eps = u + g*(c-u);  latents = c1*latents + c2*eps   on [B,T,256]; noise_pred_dev [2B,T,256]. */
LADIFF_API int ladiff_cfg_ddim_step(ladiff_handle* h, const float* noise_pred_dev, float* latents_dev, int32_t B,
                         float guidance_scale, float c1, float c2, void* stream);

/* LADiffVae.decode(z, lengths) (models/architectures/ladiff_vae.py:288-362).
 *   z_dev [T,B,256], lengths_host [B], out_dev [B,max_len,nfeats] with max_len >= max(lengths);
 *   frames >= L_i are written as exact zeros (ladiff_vae.py:358). */
LADIFF_API int ladiff_vae_decode(ladiff_handle* h, const float* z_dev, const int32_t* lengths_host, int32_t B,
                      int32_t max_len, int32_t mode, float* out_dev, void* stream);

/* LADiffVae.encode(features, lengths) (models/architectures/ladiff_vae.py:162-286; LAD branch, JOINT_DISTRO_FIX false,
 * MLP_DIST false) -- used by the evaluation glue (ladiff.py:1149 t2m_eval, stage 'vae'), "next" row of the scope table.
 *   feats_dev [B,max_len,nfeats], lengths_host [B]; token sequence per motion = m mu tokens | m logvar tokens | L frames
 *   through the non-MD SkipTransformerEncoder (operator/cross_attention.py:48-67), ragged like the decoder.
 *   eps_dev [T,B,256]: the standard-normal draw of dist.rsample() (NULL -> latent = mu)
 *   outputs (each optional, [T,B,256]): latent = mu + std*eps with rows t >= m_i exactly zero (:265-268),
 *   mu = dist.loc, std = dist.scale = exp(logvar)^0.5 for the valid rows (0 / 1 elsewhere). */
LADIFF_API int ladiff_vae_encode(ladiff_handle* h, const float* feats_dev, const int32_t* lengths_host, int32_t B,
                      int32_t max_len, const float* eps_dev, int32_t mode, float* latent_dev, float* mu_dev,
                      float* std_dev, void* stream);

/* datamodule.feats2joints (data/HumanML3D.py:44-48 -> recover_from_ric,
 * data/humanml/scripts/motion_process.py:355-381,415-430), which the reference runs on the CPU
 * after a .cpu() (ladiff.py:307).  feats_dev [B,max_len,nfeats] -> joints_dev [B,max_len,njoints,3]. */
LADIFF_API int ladiff_feats2joints(ladiff_handle* h, const float* feats_dev, const float* mean_dev, const float* std_dev,
                        int32_t B, int32_t max_len, int32_t njoints, float* joints_dev, void* stream);

/* Test hook: one fused linear  out = epilogue(A[M,K] . W[N,K]^T + bias)  through the same kernels the
 * plans use.  epilogue: 0 bias, 1 relu, 2 gelu(erf), 3 +residual, 4 LayerNorm(+residual) (N must be 256),
 * 5 SiLU(LN(.)*(1+scale)+shift) (N must be 256; mod_dev = [scale|shift], 512 floats), 6 silu.
 * All pointers fp32 device; K multiple of 64. */
LADIFF_API int ladiff_linear_test(ladiff_handle* h, const float* A_dev, const float* W_dev, const float* bias_dev,
                       const float* res_dev, const float* ln_g_dev, const float* ln_b_dev, const float* mod_dev,
                       int32_t M, int32_t N, int32_t K, int32_t epilogue, int32_t mode, float* out_dev, void* stream);

/* Measurement hook: average device time (CUDA events on `stream`) of `iters` launches of one fused linear of the given
 * shape / epilogue / mode on pseudo-random operands -- the same kernel instances the plans launch. */
LADIFF_API int ladiff_linear_bench(ladiff_handle* h, int32_t M, int32_t N, int32_t K, int32_t epilogue, int32_t mode,
                        int32_t iters, float* ms_per_launch_host, void* stream);

/* Test / measurement hook for the cluster-fused feed-forward kernel (csrc/ffn_swap.cuh): runs the two feed-forward pairs of
 * denoiser layer `layer` (reference: mdiff_transformer.py:60-62 sa_block FFN + norm2, :248-262 FFN + StylizationBlock prologue)
 * on M rows x_dev[M,256] with the finalised denoiser weights; mod_dev = [scale(256) | shift(256)].
 * fused != 0: what the plans run -- the token-group kernel k_ffn_swap for M <= 1776, the four separate fused linears above;
 * fused = 0: always the four separate fused linears.  Writes x3 and s ([M,256] fp32).  iters > 0: also returns the average
 * milliseconds of `iters` back-to-back calls (CUDA events on `stream`). */
LADIFF_API int ladiff_ffn_test(ladiff_handle* h, const float* x_dev, int32_t M, int32_t layer, const float* mod_dev, int32_t mode,
                    int32_t fused, int32_t iters, float* x3_out_dev, float* s_out_dev, float* ms_per_call_host, void* stream);

/* Profiling hook (handle created with LADIFF_TRACE=1 in the environment): per fused-linear launch of the last
 * ladiff_diffusion_reverse call, 8 x uint64 %globaltimer nanoseconds: [0] first CTA start, [1] last dependency wait returned,
 * [2] last accumulator ready, [3] last CTA done.  Synchronises the device.  Returns the number of records. */
LADIFF_API int ladiff_trace_read(ladiff_handle* h, uint64_t* out_host, int32_t max_launches, char* names_host, int32_t name_stride);

/* Number of kernel launches (graph nodes included) enqueued by the last compute call on this handle. */
LADIFF_API int64_t ladiff_last_launch_count(const ladiff_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* LADIFF_B200_H */
