/*
 * diffpir_b200.h - C-ABI of libdiffpir_b200.so: the B200 (sm_100a) DiffPIR sampling hot path.
 *
 * The reference (yuanzhi-zhu/DiffPIR) is pure Python/PyTorch and has no FFI; its "operator API" for
 * this path is a set of Python call signatures used by the four drivers.  Each
 * entry point below names the reference call it stands in for (file:line relative to the reference
 * root).  diffpir_b200/*.py binds these with ctypes and re-exposes the reference signatures.
 *
 * Conventions
 *   - plain C: pointers + sizes only; no torch / C++ types cross the boundary.
 *   - every pointer marked "dev" is a CUDA device pointer owned by the CALLER (contiguous, 16-byte aligned);
 *     the library owns only what dpir_create / dpir_fft_plan_create allocate and frees it in *_destroy.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, the library never
 *     synchronises and never changes the current device.
 *   - return value: 0 = ok, negative = error (DPIR_E_*); dpir_last_error() returns a message for the
 *     calling thread.  No CPU fallback exists: without a GPU every compute call fails with DPIR_E_CUDA.
 *   - handles are not thread-safe.
 */
#ifndef DIFFPIR_B200_H
#define DIFFPIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define DPIR_OK 0
#define DPIR_E_ARG (-1)      /* bad argument / unsupported shape */
#define DPIR_E_CUDA (-2)     /* CUDA runtime / driver error */
#define DPIR_E_CUFFT (-3)    /* cuFFT error */
#define DPIR_E_STATE (-4)    /* call order (e.g. forward before weights are complete) */
#define DPIR_E_NOMEM (-5)

#define DPIR_PRECISION_FP32 0  /* fp32 CUDA-core convolutions: parity mode (<=1e-4 rel-L2 vs reference) */
#define DPIR_PRECISION_FP16 1  /* fp16 operands, fp32 accumulate on tcgen05 tensor cores; fp32 residual stream */
#define DPIR_PRECISION_FP16S 2 /* as FP16, and the residual stream / skip tensors are stored in fp16 as well (what the
                                  reference's own use_fp16=True mode does, unet.py:620-632); GroupNorm statistics, softmax and
                                  all accumulation stay fp32/fp64: throughput mode */

#define DPIR_PRECISION_FP16X3 3 /* split-precision tensor-core mode: every convolution operand is a pair of fp16 values
                                  (hi = fp16(v), lo = fp16(v - hi)) and the tcgen05 K loop accumulates x_hi*w_hi + x_lo*w_hi +
                                  x_hi*w_lo in fp32 (products exact, dropped term ~2^-24): fp32-equivalent convolutions at 3x
                                  the MMA work; fp32 residual stream, exact SiLU, fp32 attention: the <=1e-4 rel-L2 parity
                                  mode on tensor cores */

const char *dpir_last_error(void);
int dpir_version(void);
/* number of CUDA kernels launched by this library on the calling thread since the last reset */
int64_t dpir_launch_count(void);
void dpir_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * UNet denoiser engine.  Replaces the model object built by
 *   guided_diffusion/script_util.py:74-184 create_model_and_diffusion / create_model  and
 *   guided_diffusion/unet.py:396-663 UNetModel (forward :634-663).
 * ------------------------------------------------------------------------------------------------ */
typedef struct dpir_engine dpir_engine;

typedef struct dpir_unet_config {
  int32_t image_size;          /* script_util.py:147-158 picks channel_mult from it; informational here */
  int32_t in_channels;         /* 3 */
  int32_t model_channels;      /* 128 (FFHQ) / 256 (ImageNet); must be a multiple of 128 */
  int32_t out_channels;        /* 6 (learn_sigma) */
  int32_t num_res_blocks;
  int32_t n_channel_mult;
  int32_t channel_mult[8];
  int32_t n_attention_ds;
  int32_t attention_ds[8];     /* downsample rates with attention, unet.py:501, 577 */
  int32_t num_heads;
  int32_t num_head_channels;   /* 64; -1 => use num_heads */
  int32_t num_heads_upsample;  /* -1 => num_heads */
  int32_t use_new_attention_order; /* 0: QKVAttentionLegacy unet.py:328-358, 1: QKVAttention :361-393 */
  int32_t precision;           /* DPIR_PRECISION_* */
  int32_t diffusion_steps;     /* 1000: size of the eps->x0 coefficient tables */
  int32_t channel_mult_den;    /* 0 or 1: channel_mult holds the integer multipliers; d > 1: level channels =
                                  int(channel_mult[i] * model_channels / d) - the 512x512 model's (0.5, 1, 1, 2, 2, 4, 4)
                                  (script_util.py:149-150) is passed as (1, 2, 2, 4, 4, 8, 8) with d = 2 */
} dpir_unet_config;

int dpir_create(const dpir_unet_config *cfg, dpir_engine **out);
int dpir_destroy(dpir_engine *e);

/* model.load_state_dict (main_ddpir.py:234): one call per checkpoint tensor, names exactly as in the
 * reference state_dict (time_embed.0.weight, input_blocks.3.0.in_layers.2.weight, ...).  `data` is a
 * HOST pointer to contiguous fp32 in the checkpoint layout (conv OIHW, conv1d (O,I,1), linear (O,I)).
 * The tensor is repacked into the kernel-native layout inside the device weight arena. */
int dpir_load_weight(dpir_engine *e, const char *name, const float *data, const int64_t *shape, int32_t ndim);
/* number of checkpoint tensors still missing (0 => ready) */
int dpir_missing_weights(dpir_engine *e);
/* packed device weight arena (for the one-time NCCL broadcast); after a broadcast
 * into it call dpir_mark_weights_loaded. */
int dpir_weight_arena(dpir_engine *e, void **dev_ptr, size_t *bytes);
int dpir_mark_weights_loaded(dpir_engine *e);

/* The path's ONE collective, inside the library.  The reference is a single process that loads the
 * checkpoint on its only GPU (main_ddpir.py:214-236); with one process per GPU, rank `root` repacks the checkpoint once
 * (dpir_load_weight) and dpir_bcast_weights ships the packed arena to its peers with one ncclBroadcast over NVLink - the other
 * ranks never read the checkpoint.  NCCL is bound at run time (the libnccl.so.2 the process already carries, e.g. PyTorch's, else
 * the system one); without it these calls fail with DPIR_E_STATE - there is no host-staged fallback.
 *   dpir_comm_unique_id : rank 0 fills `id` (DPIR_COMM_ID_BYTES bytes) and the host program hands it to the other ranks over any
 *                         channel it has (a file, a TCP store, MPI, torch.distributed.broadcast_object_list).
 *   dpir_comm_create    : collective - every rank calls it with the same id, its rank and the job's world size, with its GPU as the
 *                         current CUDA device (ncclCommInitRank).
 *   dpir_bcast_weights  : collective, stream-ordered on `stream`: on return the broadcast is enqueued; the non-root engines are
 *                         marked loaded (their derived FiLM table is rebuilt by the next forward, which must run on the same stream
 *                         or after a synchronisation).  Every rank must have created its engine from the same dpir_unet_config and
 *                         precision (identical arena layout); the root's checkpoint must be complete.
 */
#define DPIR_COMM_ID_BYTES 128
typedef struct dpir_comm dpir_comm;
int dpir_nccl_version(int32_t *version);                       /* e.g. 22809; DPIR_E_STATE when no NCCL library can be bound */
int dpir_comm_unique_id(void *id, size_t cap);
int dpir_comm_create(const void *id, int32_t world, int32_t rank, dpir_comm **out);
/* a host program that already owns an ncclComm_t (of the SAME libnccl the process carries) hands it over instead: `world` / `rank`
 * as it was initialised with, its GPU current; the communicator is borrowed - dpir_comm_destroy does not destroy it */
int dpir_comm_adopt(void *nccl_comm, int32_t world, int32_t rank, dpir_comm **out);
int dpir_comm_destroy(dpir_comm *c);
int dpir_bcast_weights(dpir_engine *e, dpir_comm *c, int32_t root, void *stream);

/* Optional: override the built-in tables with host-computed ones (so they are bit-identical to the reference's numpy /
 * torch evaluation): freqs[model_channels/2] of nn.py:112-114, coef[T][2] = (sqrt_recip_alphas_cumprod,
 * sqrt_recipm1_alphas_cumprod) of gaussian_diffusion.py:150-151 cast to float (:905).  Either pointer may be NULL. */
int dpir_set_tables(dpir_engine *e, const float *freqs, int32_t n_freqs, const float *coef, int32_t T);

/* UNetModel.forward (unet.py:634-663): x dev (B,3,H,W) fp32 NCHW, t dev (B,) int64 -> out dev (B,6,H,W) fp32. */
int dpir_unet_forward(dpir_engine *e, const float *x, const int64_t *t, float *out,
                      int32_t B, int32_t H, int32_t W, void *stream);
/* utils_model.model_fn(..., model_out_type='pred_xstart') = diffusion.p_sample(...)['pred_xstart']
 * (utils/utils_model.py:207-258; gaussian_diffusion.py:258-311, 328-333): UNet forward with the
 * eps->x0 conversion x0 = clamp(c1[t] x - c2[t] eps, -1, 1) fused into the output convolution. */
int dpir_unet_x0(dpir_engine *e, const float *x, const int64_t *t, float *x0,
                 int32_t B, int32_t H, int32_t W, void *stream);
/* workspace bytes currently held for activations (diagnostics) */
int64_t dpir_workspace_bytes(dpir_engine *e);
/* 1 (default) => the FiLM (scale, shift) vectors of all ResBlocks are tabulated once for every integer timestep
 * 0..diffusion_steps-1 from the loaded weights, so a forward does no embedding work; timesteps must then be integers in
 * that range (what DiffPIR passes: utils_model.py:215-217).  0 => evaluate timestep_embedding / time_embed / emb_layers
 * (nn.py:103-121, unet.py:471-475, 199-205) per call for arbitrary timesteps. */
int dpir_set_film_table(dpir_engine *e, int32_t enable);
/* 1 => replay the forward from a captured CUDA graph (per batch shape) instead of stream launches */
int dpir_set_graph_mode(dpir_engine *e, int32_t enable);

/* Per-op device timing of the forward tape with CUDA events on the launch stream (used by bench.py for the roofline
 * accounting).  Between begin and end every forward runs un-graphed with an event pair around each op; `end`
 * synchronises, sums the elapsed time per tape op over all forwards since `begin` and fills up to `cap` rows:
 * kind[i] (0 memset, 1 time-embed, 2 FiLM linears, 3 stem conv, 4 GN stats, 5 GN apply, 6 conv, 7 attention),
 * dims[6i..] = (H, W, Cin, Cout, ksize, Cin2) for convs (Cin2 = channels of a fused 1x1 skip operand), (H, W, C, resample,
 * 0, 0) for GN apply, (HW, 1, C, 0, 0, 0) for GN stats, (T, 1, C, heads, 0, 0) for attention; ms[i] = summed milliseconds.
 * Returns the number of tape ops. */
int dpir_profile_begin(dpir_engine *e);
int dpir_profile_end(dpir_engine *e, int32_t *kind, int32_t *dims, float *ms, int32_t cap);

/* ------------------------------------------------------------------------------------------------
 * FFT-domain data-fidelity solver.  Replaces utils/utils_sisr.py:
 *   p2o :22-41, upsample :44-52, pre_calculate :78-95, splits :9-19, data_solution :65-75.
 * Spectra are Hermitian half-spectra (H, W/2+1) complex64 (interleaved re,im).
 * ------------------------------------------------------------------------------------------------ */
typedef struct dpir_fft_plan dpir_fft_plan;
/* batch = number of (H,W) planes transformed per call */
int dpir_fft_plan_create(int32_t batch, int32_t H, int32_t W, dpir_fft_plan **out);
int dpir_fft_plan_destroy(dpir_fft_plan *p);
int dpir_rfft2(dpir_fft_plan *p, const float *in, float *out_spec, void *stream);
/* unnormalised inverse; destroys spec */
int dpir_irfft2(dpir_fft_plan *p, float *spec, float *out, void *stream);

/* pre_calculate (utils_sisr.py:78-95).  y dev (B,C,H/sf,W/sf); k dev (Bk,1,kh,kw) with Bk in {1,B};
 * outputs: FB (Bk,1,H,Wh) complex, F2B (Bk,1,H,Wh) real, FBFy (B,C,H,Wh) complex, Wh = W/2+1.
 * plan_k has batch Bk, plan_y batch B*C.  scratch: dev, >= max(Bk, B*C) * H * W floats. */
int dpir_fft_precalc(dpir_fft_plan *plan_k, dpir_fft_plan *plan_y, const float *y, const float *k,
                     int32_t B, int32_t C, int32_t Bk, int32_t kh, int32_t kw, int32_t sf,
                     float *FB, float *F2B, float *FBFy, float *scratch, void *stream);

/* data_solution (utils_sisr.py:65-75) evaluated in the cancellation-free form
 *   FX = Fx + conj(FB) (FBFy/conj(FB) - mean_alias(FB Fx)) / (mean_alias|FB|^2 + alpha)
 * with the affine pre/post steps of the loop (main_ddpir.py:396-400) folded in:
 *   x_in  = in_scale * x + in_bias;   Xest = data_solution(x_in, ...);
 *   out   = blend * (out_scale * Xest + out_bias) + (1 - blend) * x
 * Reference semantics (x -> Xest): in_scale=1,in_bias=0,out_scale=1,out_bias=0,blend=1.
 * Loop semantics on x0 in [-1,1]: in 0.5,0.5; out 2,-1; blend = guidance_scale.
 * spec: dev scratch, B*C*H*Wh complex.  out may alias x. */
int dpir_fft_solve(dpir_fft_plan *plan_x, const float *x, const float *FB, const float *F2B, const float *FBFy,
                   float alpha, int32_t B, int32_t C, int32_t Bk, int32_t sf,
                   float in_scale, float in_bias, float out_scale, float out_bias, float blend,
                   float *spec, float *out, void *stream);
/* the same with alpha read from DEVICE memory (alpha_dev points at one float): the reference drivers keep rho_t on the GPU
 * (`tau = rhos[t_i].float().repeat(1,1,1,1)`, main_ddpir.py:389), so the drop-in path never synchronises the host */
int dpir_fft_solve_dev(dpir_fft_plan *plan_x, const float *x, const float *FB, const float *F2B, const float *FBFy,
                       const float *alpha_dev, int32_t B, int32_t C, int32_t Bk, int32_t sf,
                       float in_scale, float in_bias, float out_scale, float out_bias, float blend,
                       float *spec, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise steps of the loop (inline in the reference drivers).
 * ------------------------------------------------------------------------------------------------ */
/* inpaint prox, main_ddpir.py:393-394 / main_ddpir_inpainting.py:267-268:
 *   x0_p = (mask*(y_scale*y + y_bias) + tau*x0)/(mask+tau); out = x0 + guidance*(x0_p - x0).  n elements. */
int dpir_inpaint_prox(const float *x0, const float *y, const float *mask, float tau, float y_scale, float y_bias,
                      float guidance, float *out, int64_t n, void *stream);
/* tau read from device memory (one float), see dpir_fft_solve_dev */
int dpir_inpaint_prox_dev(const float *x0, const float *y, const float *mask, const float *tau_dev, float y_scale, float y_bias,
                          float guidance, float *out, int64_t n, void *stream);
/* re-noise, main_ddpir.py:451-456 with precomputed scalars:
 *   eps = (x - sa_t*x0)/s1_t;  out = sa_p*x0 + c_eps*eps + c_eta*noise_eta + c_zeta*noise_zeta
 * noise_eta may be NULL when c_eta == 0. */
int dpir_renoise(const float *x, const float *x0, const float *noise_eta, const float *noise_zeta,
                 float sa_t, float s1_t, float sa_p, float c_eps, float c_eta, float c_zeta,
                 float *out, int64_t n, void *stream);
/* eps -> x0 (gaussian_diffusion.py:328-333 + clamp :305-307) on an existing model output:
 *   model_out dev (B,Cout>=3,H,W), x dev (B,3,H,W), coef dev (B,2) = (c1,c2) per sample. */
int dpir_eps_to_x0(const float *model_out, const float *x, const float *coef, float *x0,
                   int32_t B, int32_t Cout, int32_t HW, int32_t clip, void *stream);
/* diffusion.p_sample / diffusion.ddim_sample tails on the raw UNet output (gaussian_diffusion.py:258-311, 395-439, 537-585;
 * reached from utils_model.model_fn with model_out_type 'pred_x_prev' / 'pred_x_prev_and_start' or ddim_sample=True,
 * utils/utils_model.py:219-243), one elementwise pass:
 *   x0 = clamp(c1 x - c2 eps);
 *   mode 0 (p_sample, learned-range variance): sample = pm1 x0 + pm2 x + [t!=0] exp(0.5 (f max_log + (1-f) min_log)) noise,
 *                                              f = (v+1)/2, v = model_out[:,3:6]
 *   mode 1 (ddim_sample): sample = sqrt_ac_prev x0 + ce (c1 x - x0)/c2 + [t!=0] sigma noise
 * model_out dev (B,Cout,HW); x, noise (or NULL), sample (or NULL), x0 (or NULL) dev (B,3,HW); t dev (B,) int64;
 * tab dev [T][10] fp32 rows (c1, c2, pm1, pm2, min_log, max_log, sqrt_ac_prev, ce, sigma, 0) indexed by t (a t outside
 * [0,T) traps, where the reference raises IndexError). */
int dpir_sample_step(const float *model_out, const float *x, const float *noise, const int64_t *t, const float *tab, int32_t T,
                     float *sample, float *x0, int32_t B, int32_t Cout, int32_t HW, int32_t mode, int32_t clip, void *stream);
/* generic y = a*x + b (range conversions x/2+.5, 2x-1 around the operators) */
int dpir_affine(const float *x, float a, float b, float *out, int64_t n, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Separable antialiased resize.  Replaces Resizer.forward, utils/utils_resizer.py:55-74: for each
 * resized dim, out[o] = sum_t w[t][o] * in[fov[t][o]] (gather + weighted sum); weights / field-of-view
 * tables come from Resizer.contributions :104-167 (host, init-time) and are passed in.
 *   in dev (N, Hin, Win) planes; w_h/fov_h dev (taps_h, Hout); w_w/fov_w dev (taps_w, Wout);
 *   tmp dev (N, Hout, Win) floats; out dev (N, Hout, Wout).
 * ------------------------------------------------------------------------------------------------ */
int dpir_resize_separable(const float *in, int32_t N, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout,
                          const float *w_h, const int32_t *fov_h, int32_t taps_h,
                          const float *w_w, const int32_t *fov_w, int32_t taps_w,
                          float *tmp, float *out, void *stream);
/* IBP update, main_ddpir.py:403-406 (one inner iteration), x0 in [-1,1], y and down in [0,1] low-res:
 *   out = 2*((x0/2+.5) + gamma*nearest_up(y - down, sf)/(1+rho)) - 1 */
int dpir_ibp_update(const float *x0, const float *y, const float *down, float gamma, float rho,
                    int32_t N, int32_t H, int32_t W, int32_t sf, float *out, void *stream);
/* rho read from device memory (one float), see dpir_fft_solve_dev */
int dpir_ibp_update_dev(const float *x0, const float *y, const float *down, float gamma, const float *rho_dev,
                        int32_t N, int32_t H, int32_t W, int32_t sf, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Blur operator H (reference utils/utils_deblur.py:641-727 Blurkernel.forward: ReflectionPad2d(k//2) +
 * depthwise conv).  x dev (B,C,H,W), k dev (kh,kw) shared by all planes. */
int dpir_blur_reflect(const float *x, const float *k, int32_t planes, int32_t H, int32_t W, int32_t kh, int32_t kw,
                      float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Edges of the loop ("next" rows of SURVEY.md section 8(f)).
 * ------------------------------------------------------------------------------------------------ */
/* tensor2uint_batch (utils/utils_image.py:238-242): x dev (B,C,H,W) float -> out dev (B,H,W,C) uint8 =
 * round_half_even(clamp(x,0,1)*255). */
int dpir_tensor2uint(const float *x, uint8_t *out_nhwc, int32_t B, int32_t C, int32_t H, int32_t W, void *stream);
/* per-image sum of squared differences behind calculate_psnr_batch (utils_image.py:601-610):
 * sse[i] = sum over image i of ((a_scale*a+a_bias) - (b_scale*b+b_bias))^2; luma != 0 compares only
 * Y = 0.299R+0.587G+0.114B (rgb2ycbcr_batch only_y, :470-490).  a, b dev (B,C,HW) float; sse dev (B) double. */
int dpir_sqdiff_batch(const float *a, const float *b, float a_scale, float a_bias, float b_scale, float b_bias, int32_t luma,
                      int32_t B, int32_t C, int32_t HW, double *sse, void *stream);
/* measurement synthesis for deblurring (main_ddpir.py:97-114, main_ddpir_deblur.py:193-199): circular blur of the uint8
 * image exactly as scipy.ndimage.convolve(img, k[...,None], mode='wrap') computes it (fp64, same tap order, truncated to
 * uint8), then /255, *2-1, + noise, /2+.5.  img dev (B,H,W,C) uint8; k dev (B,kh,kw) double; noise dev (B,H,W,C) float or
 * NULL (AWGN drawn by the caller, std 2*sigma); blur_u8 dev (B,H,W,C) uint8 or NULL; out dev (B,C,H,W) float or NULL. */
int dpir_measure_blur_u8(const uint8_t *img_nhwc, const double *k, const float *noise_nhwc, uint8_t *blur_u8, float *out_nchw,
                         int32_t B, int32_t H, int32_t W, int32_t C, int32_t kh, int32_t kw, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Low-level entry points exposed for per-kernel parity tests and benchmarks (tests/ and bench.py).
 * NHWC activations.  act_dtype: 0 = fp32, 1 = fp16.
 * ------------------------------------------------------------------------------------------------ */
/* 3x3 (pad 1) or 1x1 stride-1 convolution on tensor cores (tcgen05): act dev fp16 (B,H,W,Cin),
 * w dev fp16 packed [tap][CoutPad][Cin], bias dev fp32 (Cout), residual dev fp32 (B,H,W,Cout) or NULL,
 * out dev fp32 (B,H,W,Cout).  Cin % 64 == 0. */
int dpir_conv_tc(const void *act, const void *w, const float *bias, const float *residual, float *out,
                 int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t CoutPad, int32_t ksize,
                 void *stream);
/* the same convolution in its fp16-stream form: residual fp16 NHWC (or NULL), output fp16 NHWC. */
int dpir_conv_tc_f16io(const void *act, const void *w, const float *bias, const void *residual_half, void *out_half, int32_t B,
                       int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t CoutPad, int32_t ksize, void *stream);
/* Fused-operand form (k_conv3x3_v3<FUSE>, "v4"): replaces GroupNorm32 + SiLU (+ FiLM) + conv3x3 of a ResBlock
 * (guided_diffusion/unet.py:199-212, 246-251; nn.py:17-19) by ONE kernel.  dpir_gn_coef folds the per-tensor quad statistics
 * q1/q2 ([B][C/4][2] sum, sumsq of the one or two concatenated sources), gamma/beta and optional per-sample FiLM rows
 * (film[b*film_stride + c] = scale, [.. + C + c] = shift) into coef dev fp32 [B][2][C] = (0.5*scale_c, 0.5*offset_c);
 * dpir_conv_gnfuse_f16io computes out = conv3x3(SiLU(scale*x + offset)) + bias (+ residual) with x = (src1 | src2) raw fp16
 * NHWC (C1, C2 multiples of 64; src2 NULL / C2 0 for one source), zero padding applied AFTER the activation, fp16 NHWC
 * residual/output.  CoutPad == 16 (the UNet's output head out.0-2, unet.py:617-621, <= 16 output channels): one source, no residual,
 * and `out_half` receives fp32 NHWC [B][H][W][Cout] (the narrow-output kernel k_conv3x3_dxn stores fp32). */
int dpir_gn_coef(const double *q1, const double *q2, int32_t C1, int32_t C2, const float *gamma, const float *beta, const float *film,
                 int32_t film_stride, int32_t B, int32_t HW, float *coef, void *stream);
int dpir_conv_gnfuse_f16io(const void *src1, int32_t C1, const void *src2, int32_t C2, const float *coef, const void *w,
                           const float *bias, const void *residual_half, void *out_half, int32_t B, int32_t H, int32_t W,
                           int32_t Cout, int32_t CoutPad, void *stream);
/* GroupNorm+SiLU inside the convolutions of eligible ResBlocks (fp16s pipeline).  1 (default; 3 is an alias): producers that
 * activate each input element once per 64-channel chunk and publish the three horizontal tap boxes from registers
 * (k_conv3x3_v3<3>, fp32 operand arithmetic = the activation values of the separate pass); 4: round-1 producers (one
 * activation per tap box, fp32 arithmetic); 2: the same with packed-half arithmetic (tanh.approx.f16x2); 0: every norm is a
 * separate k_gn_apply pass.  Per engine (drops the cached plans of `e`); with e == NULL it sets the flavour used by the
 * per-kernel test entry dpir_conv_gnfuse_f16io. */
int dpir_set_gn_fusion(dpir_engine *e, int32_t enable);
/* The attention core of AttentionBlock (guided_diffusion/unet.py:299-305 calls QKVAttentionLegacy :337-354 / QKVAttention :371-389)
 * on its own: qkv dev fp32 [B][T][3C] (NHWC output of the qkv 1x1 convolution; legacy channel order [head][q,k,v][64], new order
 * [q,k,v][head][64]); out dev [B][T][C] fp16 when out_half (tcgen05 kernel: S = QK^T and O = PV on the tensor cores with TMEM
 * accumulators, fp32 softmax) or fp32 (fp32 CUDA-core kernel).  Head dim 64, any T >= 1. */
int dpir_attention(const float *qkv, void *out, int32_t out_half, int32_t B, int32_t T, int32_t C, int32_t heads, int32_t new_order,
                   void *stream);
/* A/B switch for tests and benchmarks.  0 (default): 3x3 convs on >=16x16 maps use the persistent swap-AB halo-reuse
 * kernel (k_conv3x3_v3), everything else the one-tile-per-CTA kernel (k_conv_tc); 1: k_conv_tc everywhere;
 * 2: the persistent kernel with pixels on the M side (k_conv3x3_v2).  Affects plans created afterwards. */
int dpir_conv_tc_variant(int32_t variant);
/* same contract on CUDA cores in fp32: act fp32, w fp32 packed [tap][Cin][Cout]. */
int dpir_conv_simt(const float *act, const float *w, const float *bias, const float *residual, float *out,
                   int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t ksize, void *stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DIFFPIR_B200_H */
