// unet_kernels.cuh - launch wrappers for the non-GEMM UNet kernels and the fp32 CUDA-core convolution.
// All activations are NHWC; "virtual concat" sources (src1|src2 along C) implement th.cat of unet.py:660 without a copy.
#pragma once
#include "common.cuh"

namespace dpir {

enum Resample { RS_NONE = 0, RS_DOWN = 1, RS_UP = 2 };
enum EpiMode { EPI_NHWC = 0, EPI_NCHW = 1, EPI_X0 = 2 };

struct GnApplyArgs {
  const void *src1; int C1;           // fp32 (or fp16 when src_half) NHWC
  const void *src2; int C2;           // may be null / 0; same dtype as src1
  int src_half;                       // 1: the sources are __half
  const double *q1, *q2;              // per-tensor quad statistics [B][C/4][2] (sum, sumsq) of src1 / src2
  const float *gamma, *beta;          // [C]
  const float *film; int film_stride; // per-sample (scale[C], shift[C]) at film + idx*film_stride, or null;
  const int64_t *tsteps;              // idx = tsteps[b] (film is the per-timestep table) when non-null, else idx = b
  int silu;                           // apply SiLU after the affine/FiLM
  int resample;                       // Resample
  int B, H, W;                        // input spatial size
  int out_half;                       // 1: act / rawact are __half, 0: float
  int split;                          // 1 (with out_half): act / rawact hold [hi | lo] fp16 pairs, 2C channels per pixel:
                                      // hi = fp16(v), lo = fp16(v - hi) (split-precision mode); SiLU is then evaluated exactly
  void *act;                          // activated output (resampled size)
  float *raw32;                       // optional: raw (un-normalised) fp32 copy of the concat, resampled
  void *rawact;                       // optional: raw copy in the activation dtype, resampled
};

struct ConvEpi {
  int mode;                 // EpiMode
  const float *bias;        // [Cout]
  const float *bias2;       // second bias (fused 1x1 skip convolution) or null; tensor-core NHWC epilogue only
  const float *residual;    // NHWC [B,H,W,Cout] fp32 (or fp16 when res_half) or null (EPI_NHWC only)
  int res_half;             // 1: residual is __half (tensor-core epilogues only)
  int res_up;               // 1 (swap-AB kernel only): the residual tensor has HALF the resolution (B, H/2, W/2, Cout) and is read through
                            // the nearest-neighbour index map (y >> 1, x >> 1): the h_upd / x_upd of an up-sampling ResBlock
                            // (unet.py:237-244) without materialising the up-sampled tensor
  float *out;               // NHWC [B,H,W,Cout] | NCHW [B,Cout,H,W] | x0 NCHW [B,3,H,W]
  int out_half;             // EPI_NHWC on tensor cores: store __half instead of float (out is then a __half*)
  double *qstats;           // EPI_NHWC on tensor cores: accumulate quad statistics [B][Cout/4][2] of the stored tensor, or null
  const float *x_nchw;      // EPI_X0: x_t (B,3,H,W)
  const float *coef;        // EPI_X0: [T][2] (sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod) as float
  const int64_t *t;         // EPI_X0: (B,) timesteps on device
  const float *acc_scale;   // device scalar multiplied into the accumulator before the bias (split-precision mode: the weights are
                            // stored scaled by a power of two so that their fp16 hi/lo parts stay in the normal range), or null
  int coef_rows;            // EPI_X0: rows of coef (diffusion steps); t is clamped into [0, coef_rows) before the lookup
};

// v4 convolution operand fusion: the activation operand is SiLU(GroupNorm(+FiLM)(src1|src2)) evaluated inside the kernel
struct GnFuse {
  const __half *src1 = nullptr, *src2 = nullptr;   // raw NHWC fp16 source(s); src2 = second half of a virtual concat or null
  int C1 = 0, C2 = 0;
  const float *coef = nullptr;                     // [B][2][C1+C2] from launch_gn_coef: 0.5*scale, 0.5*offset; null = no fusion
  int up = 0;                                      // 1: src1 has HALF the convolution's resolution and is read through the nearest-
                                                   // neighbour index map (y >> 1, x >> 1): conv(upsample(SiLU(GN(x)))) of an up-sampling
                                                   // ResBlock (unet.py:237-241) without the up-sampled tensor (activate-once producers only)
};

// per-(image, channel) affine form of GroupNorm(+FiLM): act = SiLU(scale*x + offset); writes 0.5*scale, 0.5*offset
struct GnCoefArgs {
  int C1, C2;                         // channels of the two sources (C2 = 0: one source)
  const double *q1, *q2;              // quad statistics of the sources
  const float *gamma, *beta;          // [C]
  const float *film; int film_stride; const int64_t *tsteps;   // as GnApplyArgs
  int B, HW;
  float *coef;                        // [B][2][C]
};
int launch_gn_coef(const GnCoefArgs &a, cudaStream_t st);

#ifdef __CUDACC__
// shared conv epilogue (bias / residual / NCHW / fused eps->x0); acc[0..ncols) = channels co.. of pixel (n,y,x)
__device__ __forceinline__ void conv_epilogue_store(const ConvEpi &e, int n, int y, int x, int H, int W, int Cout, int co,
                                                    const float *acc, int ncols) {
  // acc[0..ncols) are output channels co..co+ncols of pixel (n,y,x)
  const float sc = e.acc_scale ? __ldg(e.acc_scale) : 1.f;
  if (e.mode == EPI_NHWC) {
    int64_t pix = ((int64_t)n * H + y) * W + x;
    float *o = e.out + pix * Cout + co;
    const float *r = e.residual ? e.residual + pix * Cout + co : nullptr;
    if (ncols == 4 && (Cout & 3) == 0) {
      float4 b = *reinterpret_cast<const float4 *>(e.bias + co);
      float4 v = make_float4(acc[0] * sc + b.x, acc[1] * sc + b.y, acc[2] * sc + b.z, acc[3] * sc + b.w);
      if (r) { float4 rv = *reinterpret_cast<const float4 *>(r); v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w; }
      *reinterpret_cast<float4 *>(o) = v;
    } else {
      for (int j = 0; j < ncols; ++j) o[j] = acc[j] * sc + e.bias[co + j] + (r ? r[j] : 0.f);
    }
  } else if (e.mode == EPI_NCHW) {
    for (int j = 0; j < ncols; ++j)
      e.out[(((int64_t)n * Cout + co + j) * H + y) * W + x] = acc[j] * sc + e.bias[co + j];
  } else {  // EPI_X0: x0 = clamp(c1[t] x - c2[t] eps, -1, 1), eps = channels 0..2   [gaussian_diffusion.py:305-311,328-333]
    int64_t tt = e.t[n];
    tt = tt < 0 ? 0 : (tt >= e.coef_rows ? e.coef_rows - 1 : tt);
    float c1 = e.coef[2 * tt], c2 = e.coef[2 * tt + 1];
    for (int j = 0; j < ncols; ++j) {
      int c = co + j;
      if (c < 3) {
        int64_t idx = (((int64_t)n * 3 + c) * H + y) * W + x;
        float v = c1 * e.x_nchw[idx] - c2 * (acc[j] * sc + e.bias[c]);
        e.out[idx] = fminf(fmaxf(v, -1.f), 1.f);
      }
    }
  }
}
#endif

int launch_conv_in(const float *x_nchw, const float *w, const float *bias, void *out, int out_half, int B, int H, int W, int Cin,
                   int Cout, cudaStream_t st);
// NCHW fp32 (B,Cin<=8,H,W) -> NHWC fp16 [B][H][W][8] (zero-padded to 8 channels = 16 bytes per pixel: the tensor-core stem operand;
// the convolution's TMA box spans a full 64-channel K chunk and the hardware zero-fills channels 8..63)
int launch_pack_stem(const float *x_nchw, __half *out, int B, int Cin, int H, int W, cudaStream_t st);
int launch_gn_stats(const void *src, int src_half, int C, double *qstats, int B, int HW, cudaStream_t st);
int launch_gn_apply(const GnApplyArgs &a, cudaStream_t st);
int launch_conv_simt(const float *act, const float *w, int ldw, const ConvEpi &epi, int B, int H, int W, int Cin, int Cout,
                     int ksize, cudaStream_t st);
// fp32 CUDA-core attention: qkv fp32 [B,T,3C] -> out [B,T,C] (dtype per out_half); head dim 64, any T
int launch_attention(const float *qkv, void *out, int out_half, int B, int T, int C, int heads, int new_order, cudaStream_t st);
// tcgen05 attention (attention_tc.cu): the plan holds the TMA descriptor over the fp16 qkv tensor [B][T][3C]
struct AttnTcPlan;
int attn_tc_plan_create(const void *qkv_half, int B, int T, int C, AttnTcPlan **out);
void attn_tc_plan_destroy(AttnTcPlan *p);
int launch_attention_tc(const AttnTcPlan *p, void *out, int out_half, int heads, int new_order, cudaStream_t st);
int launch_cast_half(const float *in, __half *out, int64_t n, cudaStream_t st);
// fp32 [npix][C] -> fp16 [npix][hi(C) | lo(C)] (split-precision operand form)
int launch_split_cast(const float *in, __half *out, int64_t npix, int C, cudaStream_t st);
// out[i] = clamp(t[i], 0, T-1): the engine's table lookups never leave [0, T)
int launch_clamp_t(const int64_t *t, int64_t *out, int B, int T, cudaStream_t st);
// sinusoidal embedding + time_embed MLP (nn.py:103-121, unet.py:471-475): t (B,) int64 -> emb_act = SiLU(emb) [B][ted]
int launch_time_embed(const int64_t *t, const float *freqs, const float *w0, const float *b0, const float *w2, const float *b2,
                      float *hidden, float *emb_act, int B, int mc, int ted, cudaStream_t st);
// y[b][o] = act_out(sum_i act_in(x[b][i]) W[o][i] + bias[o])  (emb_layers of every ResBlock in one launch)
int launch_linear_rows(const float *x, const float *W, const float *bias, float *y, int B, int in_dim, int rows, int silu_in,
                       int silu_out, cudaStream_t st);

// tcgen05 convolution (conv_tc.cu)
struct ConvTcPlan;  // opaque: tensor maps + geometry
// split != 0: split-precision operands - act holds [hi | lo] (2*Cin channels per pixel), w is packed [tap][CoutPad][3*Cin] =
// [w_hi | w_hi | w_lo]; the K loop runs x_hi*w_hi + x_lo*w_hi + x_hi*w_lo in one fp32 accumulation (error ~2^-22)
int conv_tc_plan_create(const void *act_half, const void *w_half, int B, int H, int W, int Cin, int CoutPad, int ksize,
                        ConvTcPlan **out, int split = 0, int force_bn = 0, int act_c = 0);
void conv_tc_plan_destroy(ConvTcPlan *p);
bool conv_tc_plan_is_v2(const ConvTcPlan *p);
bool conv_tc_narrow_eligible(int H, int W, int Cin, int CoutPad);   // 3x3, <= 16 couts, >= 16x16 map: k_conv3x3_dxn
int conv_tc_plan_add_skip(ConvTcPlan *p, const void *act2_half, const void *w2_half, int Cin2a, const void *act2b_half = nullptr,
                          int Cin2b = 0);
// math: producer flavour - 3 activate-once (default), 1 per-tap fp32, 2 per-tap packed half, 0 = the process-wide test default
int conv_tc_plan_set_gn_fuse(ConvTcPlan *p, const GnFuse &gf, int math = 0);
int conv_tc_plan_set_io(ConvTcPlan *p, const void *out, int out_half, const void *residual, int res_half, int Cout, int res_up = 0);
int launch_conv_tc(const ConvTcPlan *p, const ConvEpi &epi, int Cout, cudaStream_t st);
bool conv_tc_v2_enabled();
void conv_tc_set_gnfuse_math(int m);   // producer flavour used by the per-kernel test entry (3 default, 1, 2)
bool conv_tc_swapab_enabled();   // the swap-AB persistent kernel (v3 / fused-operand v4) is selectable
void conv_tc_force_v1(int on);   // 1: use the one-tile-per-CTA kernel for every shape (A/B testing)

}  // namespace dpir
