// unet_kernels.cu - GroupNorm/SiLU/FiLM/resample passes, input stem, fp32 CUDA-core implicit-GEMM convolution,
// fused QKV attention and the timestep-embedding MLPs of the guided-diffusion UNet.
//
// Reference behaviour (paths relative to the reference root):
//   guided_diffusion/nn.py:17-19, 93-100  GroupNorm32(32, C), eps 1e-5
//   guided_diffusion/unet.py:236-256      ResBlock._forward (GN+SiLU -> up/down -> conv; FiLM: GN(h)*(1+scale)+shift)
//   guided_diffusion/unet.py:81-140       Upsample (nearest x2) / Downsample (avg-pool 2)
//   guided_diffusion/unet.py:337-354      QKVAttentionLegacy ([head][q,k,v][ch] channel order), :371-389 new order
//   guided_diffusion/nn.py:103-121        timestep_embedding;  unet.py:471-475 time_embed; :199-205 emb_layers
// Layout: activations NHWC so that channels are the contiguous (coalesced, float4) dimension and so that the same
// buffers feed the tcgen05 implicit-GEMM convolution (conv_tc.cu) as K-major operands.
#include "unet_kernels.cuh"
#include <cstdlib>
#include <string>

namespace dpir {

// =========================================================================================== vector load/store helpers
template <typename T> struct Store4;
template <> struct Store4<float> {
  static __device__ __forceinline__ void st(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
};
template <> struct Store4<__half> {
  static __device__ __forceinline__ void st(__half *p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t *>(&a);
    u.y = *reinterpret_cast<uint32_t *>(&b);
    *reinterpret_cast<uint2 *>(p) = u;
  }
};
template <typename T> struct Load4;
template <> struct Load4<float> {
  typedef float4 raw_t;
  static constexpr int UNROLL = 4;
  static __device__ __forceinline__ raw_t raw(const float *p) { return *reinterpret_cast<const float4 *>(p); }
  static __device__ __forceinline__ float4 cvt(raw_t r) { return r; }
  static __device__ __forceinline__ float4 ld(const float *p) { return *reinterpret_cast<const float4 *>(p); }
};
template <> struct Load4<__half> {
  typedef uint2 raw_t;
  static constexpr int UNROLL = 8;      // 8 x 8 B in flight per thread: the streaming passes need ~45 KB in flight per SM
  static __device__ __forceinline__ raw_t raw(const __half *p) { return *reinterpret_cast<const uint2 *>(p); }
  static __device__ __forceinline__ float4 cvt(raw_t u) {
    float2 a = __half22float2(*reinterpret_cast<__half2 *>(&u.x)), b = __half22float2(*reinterpret_cast<__half2 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ float4 ld(const __half *p) { return cvt(raw(p)); }
};

// =========================================================================================== input stem
// input_blocks.0.0: Conv2d(Cin<=4, C, 3, padding=1) reading the caller's NCHW image and writing NHWC fp32.
// w packed [tap*Cin + ci][Cout].  One CTA = an 8 x 32 pixel tile whose (10 x 34 x CIN) input patch is staged in shared
// memory (zero-filled outside the image = the conv padding).  One warp = one tile row; a lane owns 4 output channels
// whose 9*CIN float4 weights stay in registers and produces 4 horizontally adjacent pixels per iteration from a 3 x 6
// patch per input channel (shared-memory broadcasts, no predication in the inner loop); stores are 512 B contiguous.
constexpr int CI_TH = 8, CI_TW = 32;

template <int CIN, typename TO>
__global__ void __launch_bounds__(256) k_conv_in(const float *__restrict__ x, const float *__restrict__ w,
                                                 const float *__restrict__ bias, TO *__restrict__ out, int B, int H, int W,
                                                 int Cout) {
  __shared__ float patch[CIN][CI_TH + 2][CI_TW + 2];
  const int tiles_w = (W + CI_TW - 1) / CI_TW, tiles_h = (H + CI_TH - 1) / CI_TH;
  int t = blockIdx.x;
  const int w0 = (t % tiles_w) * CI_TW;
  const int h0 = ((t / tiles_w) % tiles_h) * CI_TH;
  const int n = t / (tiles_w * tiles_h);
  for (int i = threadIdx.x; i < CIN * (CI_TH + 2) * (CI_TW + 2); i += blockDim.x) {
    int c = i / ((CI_TH + 2) * (CI_TW + 2));
    int r = (i / (CI_TW + 2)) % (CI_TH + 2);
    int col = i % (CI_TW + 2);
    int yy = h0 + r - 1, xx = w0 + col - 1;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(x + (((int64_t)n * CIN + c) * H + yy) * W + xx);
    patch[c][r][col] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int yh = h0 + warp;
  if (yh >= H) return;
  const int C4 = Cout / 4;
  for (int c4 = lane; c4 < C4; c4 += 32) {
    float4 wr[9 * CIN];
#pragma unroll
    for (int i = 0; i < 9 * CIN; ++i) wr[i] = __ldg(reinterpret_cast<const float4 *>(w + (int64_t)i * Cout) + c4);
    const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
#pragma unroll 1
    for (int xs = 0; xs < CI_TW; xs += 4) {
      float4 acc[4] = {bv, bv, bv, bv};
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          float r[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) r[j] = patch[ci][warp + dy][xs + j];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4 wv = wr[(dy * 3 + dx) * CIN + ci];
#pragma unroll
            for (int pxl = 0; pxl < 4; ++pxl) {
              const float v = r[pxl + dx];
              acc[pxl].x += v * wv.x; acc[pxl].y += v * wv.y; acc[pxl].z += v * wv.z; acc[pxl].w += v * wv.w;
            }
          }
        }
      }
      const int64_t p = ((int64_t)n * H + yh) * W + w0 + xs;
#pragma unroll
      for (int pxl = 0; pxl < 4; ++pxl)
        if (w0 + xs + pxl < W) Store4<TO>::st(out + (p + pxl) * Cout + c4 * 4, acc[pxl]);
    }
  }
}

int launch_conv_in(const float *x, const float *w, const float *bias, void *out, int out_half, int B, int H, int W, int Cin,
                   int Cout, cudaStream_t st) {
  DPIR_REQUIRE(Cout % 4 == 0, "conv_in: Cout must be a multiple of 4");
  DPIR_REQUIRE(Cin >= 1 && Cin <= 4, "conv_in: Cin must be 1..4");
  int grid = B * cdiv(H, CI_TH) * cdiv(W, CI_TW);
#define CI_LAUNCH(N)                                                                                      \
  if (out_half) k_conv_in<N, __half><<<grid, 256, 0, st>>>(x, w, bias, (__half *)out, B, H, W, Cout);     \
  else k_conv_in<N, float><<<grid, 256, 0, st>>>(x, w, bias, (float *)out, B, H, W, Cout)
  switch (Cin) {
    case 1: CI_LAUNCH(1); break;
    case 2: CI_LAUNCH(2); break;
    case 3: CI_LAUNCH(3); break;
    default: CI_LAUNCH(4); break;
  }
#undef CI_LAUNCH
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// Tensor-core pipelines run the stem (input_blocks.0.0, K = 27) on the persistent tcgen05 3x3 kernel like every other convolution:
// this pass repacks the caller's NCHW fp32 image as NHWC fp16 with 8 channels per pixel (16 bytes, zeros beyond Cin).  The
// convolution's 4-D TMA box spans a whole 64-channel K chunk of this 8-channel tensor: the hardware zero-fills channels 8..63 like
// it zero-fills the halo, so the padding costs no memory traffic (the weights are zero-padded to 64 the ordinary way).
__global__ void k_pack_stem(const float *__restrict__ x, __half *__restrict__ out, int Cin, int64_t HW, int64_t npix) {
  for (int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = pix / HW, p = pix - n * HW;
    const float *src = x + n * Cin * HW + p;
    __half h[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = __float2half_rn(c < Cin ? __ldg(src + (int64_t)c * HW) : 0.f);
    *reinterpret_cast<uint4 *>(out + pix * 8) = *reinterpret_cast<const uint4 *>(h);
  }
}
int launch_pack_stem(const float *x_nchw, __half *out, int B, int Cin, int H, int W, cudaStream_t st) {
  DPIR_REQUIRE(Cin >= 1 && Cin <= 8, "pack_stem: 1..8 input channels");
  const int64_t npix = (int64_t)B * H * W;
  int64_t blocks = (npix + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  k_pack_stem<<<(int)blocks, 256, 0, st>>>(x_nchw, out, Cin, (int64_t)H * W, npix);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== GroupNorm statistics
// Statistics are attached to the TENSOR, not to the norm: qstats[b][c/4] = (sum, sumsq) over all pixels of channel quad
// c/4 (fp64).  They are produced where the tensor is produced - the tcgen05 conv epilogue (conv_tc.cu) or this kernel
// (stem output, fp32 mode) - and any later GroupNorm over the tensor, alone or inside a channel concat whose groups
// straddle the two sources (e.g. 768 = 512 + 256 -> 24 channels per group), folds the quads of its 32 groups on the fly.
constexpr int GN_PIX_PER_BLOCK = 512;

template <typename TI>
__global__ void k_gn_stats(const TI *__restrict__ src, int C, double *__restrict__ qstats, int HW) {
  int C4 = C / 4;
  int ppb = blockDim.x / C4;
  int lane_p = threadIdx.x / C4, c4 = threadIdx.x % C4;
  int b = blockIdx.y;
  int p0 = blockIdx.x * GN_PIX_PER_BLOCK;
  int p1 = min(p0 + GN_PIX_PER_BLOCK, HW);
  const TI *base = src + (int64_t)b * HW * C + c4 * 4;
  float s = 0.f, ss = 0.f;
  int p = p0 + lane_p;
  for (; p + 3 * ppb < p1; p += 4 * ppb) {
    float4 v0 = Load4<TI>::ld(base + (int64_t)(p + 0 * ppb) * C);
    float4 v1 = Load4<TI>::ld(base + (int64_t)(p + 1 * ppb) * C);
    float4 v2 = Load4<TI>::ld(base + (int64_t)(p + 2 * ppb) * C);
    float4 v3 = Load4<TI>::ld(base + (int64_t)(p + 3 * ppb) * C);
    s += ((v0.x + v0.y) + (v0.z + v0.w)) + ((v1.x + v1.y) + (v1.z + v1.w)) + ((v2.x + v2.y) + (v2.z + v2.w)) +
         ((v3.x + v3.y) + (v3.z + v3.w));
    ss += ((v0.x * v0.x + v0.y * v0.y) + (v0.z * v0.z + v0.w * v0.w)) + ((v1.x * v1.x + v1.y * v1.y) + (v1.z * v1.z + v1.w * v1.w)) +
          ((v2.x * v2.x + v2.y * v2.y) + (v2.z * v2.z + v2.w * v2.w)) + ((v3.x * v3.x + v3.y * v3.y) + (v3.z * v3.z + v3.w * v3.w));
  }
  for (; p < p1; p += ppb) {
    float4 v = Load4<TI>::ld(base + (int64_t)p * C);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  // combine the ppb pixel-lanes of a quad through shared memory, then one fp64 atomic pair per quad per block
  extern __shared__ float red[];   // [ppb][C4][2]
  red[(lane_p * C4 + c4) * 2] = s;
  red[(lane_p * C4 + c4) * 2 + 1] = ss;
  __syncthreads();
  if (lane_p == 0) {
    double ds = 0.0, dss = 0.0;
    for (int l = 0; l < ppb; ++l) { ds += (double)red[(l * C4 + c4) * 2]; dss += (double)red[(l * C4 + c4) * 2 + 1]; }
    double *q = qstats + ((int64_t)b * C4 + c4) * 2;
    atomicAdd(q, ds);
    atomicAdd(q + 1, dss);
  }
}

static inline int gn_threads(int C) {
  int C4 = C / 4;
  int ppb = 256 / C4;
  if (ppb < 1) ppb = 1;
  return C4 * ppb;
}

int launch_gn_stats(const void *src, int src_half, int C, double *qstats, int B, int HW, cudaStream_t st) {
  DPIR_REQUIRE(C % 4 == 0 && C <= 4096, "gn_stats: channels must be a multiple of 4 (got %d)", C);
  dim3 grid(cdiv(HW, GN_PIX_PER_BLOCK), B);
  int threads = gn_threads(C);
  size_t smem = (size_t)threads * 2 * sizeof(float);
  if (src_half) k_gn_stats<__half><<<grid, threads, smem, st>>>((const __half *)src, C, qstats, HW);
  else k_gn_stats<float><<<grid, threads, smem, st>>>((const float *)src, C, qstats, HW);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== GroupNorm apply
// store 4 activated channels of work pixel `pix` (C channels per pixel): plain, or split into fp16 hi | lo halves of a 2C row
template <typename T, bool SPLIT>
__device__ __forceinline__ void store_act4(T *base, int64_t pix, int C, int c, float4 v) {
  if constexpr (SPLIT) {
    __half *row = reinterpret_cast<__half *>(base) + pix * (2 * (int64_t)C);
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t *>(&h0); uh.y = *reinterpret_cast<const uint32_t *>(&h1);
    ul.x = *reinterpret_cast<const uint32_t *>(&l0); ul.y = *reinterpret_cast<const uint32_t *>(&l1);
    *reinterpret_cast<uint2 *>(row + c) = uh;
    *reinterpret_cast<uint2 *>(row + C + c) = ul;
  } else {
    Store4<T>::st(base + pix * C + c, v);
  }
}

struct GnDev {
  const void *s1; int C1; const void *s2; int C2;
  const double *q1, *q2;             // quad statistics of the two sources
  const float *gamma, *beta; const float *film; int film_stride; const int64_t *tsteps;
  int silu, H, W;
  float *raw32;
};

// FAST: fp16-output pipelines use ex2/rcp approximations (2 ulp, far below the fp16 rounding of the stored value);
// the fp32 parity mode keeps expf and IEEE division.
// FAST (fp16-output pipelines): x*sigmoid(x) = 0.5x + 0.5x*tanh(0.5x) with ONE special-function op (tanh.approx.f32, max
// relative error 2^-11 = the fp16 rounding of the stored value) instead of ex2 + rcp: the pass is otherwise co-limited by
// the 16/clk/SM SFU (2 MUFU per element = 0.13 ms of a 0.24 ms pass at 256x256x128, B=32).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool FAST> __device__ __forceinline__ float silu1(float o) {
  if (FAST) { const float h = 0.5f * o; return fmaf(h, tanh_approx(h), h); }
  return o / (1.f + expf(-o));
}

template <bool FAST>
__device__ __forceinline__ float4 gn_act(float4 v, float mean, float rstd, float4 g, float4 be, bool has_film, float4 sc,
                                         float4 sh, int silu) {
  float4 o;
  o.x = (v.x - mean) * rstd * g.x + be.x;
  o.y = (v.y - mean) * rstd * g.y + be.y;
  o.z = (v.z - mean) * rstd * g.z + be.z;
  o.w = (v.w - mean) * rstd * g.w + be.w;
  if (has_film) {
    o.x = o.x * (1.f + sc.x) + sh.x; o.y = o.y * (1.f + sc.y) + sh.y;
    o.z = o.z * (1.f + sc.z) + sh.z; o.w = o.w * (1.f + sc.w) + sh.w;
  }
  if (silu) { o.x = silu1<FAST>(o.x); o.y = silu1<FAST>(o.y); o.z = silu1<FAST>(o.z); o.w = silu1<FAST>(o.w); }
  return o;
}

constexpr int GN_APPLY_PIX = 256;  // work pixels per block

// MODE none: one work pixel = one pixel.  MODE down: work pixel = output pixel (reads 2x2 inputs, averages the
// ACTIVATED values as unet.py:238-241 does: in_rest(x) then h_upd).  MODE up: work pixel = input pixel, written 2x2.
template <typename T, typename TI, int MODE, bool SPLIT = false>
__global__ void k_gn_apply(GnDev a, T *__restrict__ act, T *__restrict__ rawact) {
  constexpr bool FAST = sizeof(T) == 2 && !SPLIT;      // split-precision operands keep the exact SiLU
  __shared__ float s_mean[32], s_rstd[32];
  int C = a.C1 + a.C2, C4 = C / 4;
  int b = blockIdx.y;
  int H = a.H, W = a.W, HW = H * W;
  if (threadIdx.x < 32) {
    int qpg = C4 / 32, Q1 = a.C1 / 4;
    double su = 0.0, sq = 0.0;
    for (int j = 0; j < qpg; ++j) {
      int q = threadIdx.x * qpg + j;
      const double *src = q < Q1 ? a.q1 + ((int64_t)b * Q1 + q) * 2 : a.q2 + ((int64_t)b * (C4 - Q1) + (q - Q1)) * 2;
      su += src[0];
      sq += src[1];
    }
    double n = (double)HW * (double)(C / 32);
    double mean = su / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  int ppb = blockDim.x / C4;
  int lane_p = threadIdx.x / C4, c4 = threadIdx.x % C4;
  int g = c4 / (C4 / 32);
  float mean = s_mean[g], rstd = s_rstd[g];
  float4 ga = reinterpret_cast<const float4 *>(a.gamma)[c4];
  float4 be = reinterpret_cast<const float4 *>(a.beta)[c4];
  bool has_film = a.film != nullptr;
  float4 sc = make_float4(0, 0, 0, 0), sh = sc;
  if (has_film) {
    const float *f = a.film + (a.tsteps ? a.tsteps[b] : (int64_t)b) * a.film_stride;
    sc = reinterpret_cast<const float4 *>(f)[c4];
    sh = reinterpret_cast<const float4 *>(f + C)[c4];
  }
  const TI *src; int Cs, cc;
  if (c4 * 4 < a.C1) { src = reinterpret_cast<const TI *>(a.s1) + (int64_t)b * HW * a.C1; Cs = a.C1; cc = c4 * 4; }
  else               { src = reinterpret_cast<const TI *>(a.s2) + (int64_t)b * HW * a.C2; Cs = a.C2; cc = c4 * 4 - a.C1; }
  int Hw = (MODE == RS_DOWN) ? H / 2 : H, Ww = (MODE == RS_DOWN) ? W / 2 : W;   // work grid
  int Ho = (MODE == RS_UP) ? H * 2 : Hw, Wo = (MODE == RS_UP) ? W * 2 : Ww;     // output grid
  int nwork = Hw * Ww;
  int p0 = blockIdx.x * GN_APPLY_PIX;
  int p1 = min(p0 + GN_APPLY_PIX, nwork);
  int64_t obase = (int64_t)b * Ho * Wo;
  if (MODE == RS_NONE) {
    int p = p0 + lane_p;
    constexpr int UN = Load4<TI>::UNROLL;
    for (; p + (UN - 1) * ppb < p1; p += UN * ppb) {
      typename Load4<TI>::raw_t rv[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) rv[u] = Load4<TI>::raw(src + (int64_t)(p + u * ppb) * Cs + cc);
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const float4 v = Load4<TI>::cvt(rv[u]);
        int64_t o = (obase + p + u * ppb) * C + c4 * 4;
        store_act4<T, SPLIT>(act, obase + p + u * ppb, C, c4 * 4, gn_act<FAST>(v, mean, rstd, ga, be, has_film, sc, sh, a.silu));
        if (a.raw32) *reinterpret_cast<float4 *>(a.raw32 + o) = v;
        if (rawact) store_act4<T, SPLIT>(rawact, obase + p + u * ppb, C, c4 * 4, v);
      }
    }
    for (; p < p1; p += ppb) {
      float4 v = Load4<TI>::ld(src + (int64_t)p * Cs + cc);
      int64_t o = (obase + p) * C + c4 * 4;
      store_act4<T, SPLIT>(act, obase + p, C, c4 * 4, gn_act<FAST>(v, mean, rstd, ga, be, has_film, sc, sh, a.silu));
      if (a.raw32) *reinterpret_cast<float4 *>(a.raw32 + o) = v;
      if (rawact) store_act4<T, SPLIT>(rawact, obase + p, C, c4 * 4, v);
    }
  } else if (MODE == RS_DOWN) {
    for (int p = p0 + lane_p; p < p1; p += ppb) {
      int yo = p / Ww, xo = p % Ww;
      float4 v[4];
#pragma unroll
      for (int d = 0; d < 4; ++d)
        v[d] = Load4<TI>::ld(src + ((int64_t)(2 * yo + (d >> 1)) * W + 2 * xo + (d & 1)) * Cs + cc);
      float4 accA = make_float4(0, 0, 0, 0), accR = accA;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        float4 o = gn_act<FAST>(v[d], mean, rstd, ga, be, has_film, sc, sh, a.silu);
        accA.x += o.x; accA.y += o.y; accA.z += o.z; accA.w += o.w;
        accR.x += v[d].x; accR.y += v[d].y; accR.z += v[d].z; accR.w += v[d].w;
      }
      accA.x *= 0.25f; accA.y *= 0.25f; accA.z *= 0.25f; accA.w *= 0.25f;
      accR.x *= 0.25f; accR.y *= 0.25f; accR.z *= 0.25f; accR.w *= 0.25f;
      int64_t o = (obase + p) * C + c4 * 4;
      store_act4<T, SPLIT>(act, obase + p, C, c4 * 4, accA);
      if (a.raw32) *reinterpret_cast<float4 *>(a.raw32 + o) = accR;
      if (rawact) store_act4<T, SPLIT>(rawact, obase + p, C, c4 * 4, accR);
    }
  } else {
    for (int p = p0 + lane_p; p < p1; p += ppb) {
      int yi = p / W, xi = p % W;
      float4 v = Load4<TI>::ld(src + (int64_t)p * Cs + cc);
      float4 o = gn_act<FAST>(v, mean, rstd, ga, be, has_film, sc, sh, a.silu);
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int64_t qp = obase + (int64_t)(2 * yi + (d >> 1)) * Wo + 2 * xi + (d & 1);
        int64_t q = qp * C + c4 * 4;
        store_act4<T, SPLIT>(act, qp, C, c4 * 4, o);
        if (a.raw32) *reinterpret_cast<float4 *>(a.raw32 + q) = v;
        if (rawact) store_act4<T, SPLIT>(rawact, qp, C, c4 * 4, v);
      }
    }
  }
}

int launch_gn_apply(const GnApplyArgs &a, cudaStream_t st) {
  int C = a.C1 + a.C2;
  DPIR_REQUIRE(C % 128 == 0 && a.C1 % 4 == 0 && C <= 4096, "gn_apply: channels must be a multiple of 128 (got %d)", C);
  DPIR_REQUIRE(a.resample != RS_DOWN || (a.H % 2 == 0 && a.W % 2 == 0), "gn_apply: downsample needs even H, W");
  DPIR_REQUIRE(!a.src_half || a.out_half, "gn_apply: fp16 sources only in an fp16 pipeline");
  DPIR_REQUIRE(!a.src_half || !a.raw32, "gn_apply: fp16 sources produce their raw copy through rawact");
  GnDev d{a.src1, a.C1, a.src2, a.C2, a.q1, a.q2, a.gamma, a.beta, a.film, a.film_stride, a.tsteps, a.silu, a.H, a.W, a.raw32};
  int nwork = (a.resample == RS_DOWN) ? (a.H / 2) * (a.W / 2) : a.H * a.W;
  dim3 grid(cdiv(nwork, GN_APPLY_PIX), a.B);
  int threads = gn_threads(C);
#define GN_LAUNCH(T, TI, M) k_gn_apply<T, TI, M><<<grid, threads, 0, st>>>(d, (T *)a.act, (T *)a.rawact)
  DPIR_REQUIRE(!a.split || (a.out_half && !a.src_half), "gn_apply: split operands are written from fp32 sources as fp16 pairs");
#define GN_LAUNCH_S(M) k_gn_apply<__half, float, M, true><<<grid, threads, 0, st>>>(d, (__half *)a.act, (__half *)a.rawact)
  if (a.split) {
    if (a.resample == RS_NONE) GN_LAUNCH_S(RS_NONE);
    else if (a.resample == RS_DOWN) GN_LAUNCH_S(RS_DOWN);
    else GN_LAUNCH_S(RS_UP);
  } else if (a.out_half && a.src_half) {
    if (a.resample == RS_NONE) GN_LAUNCH(__half, __half, RS_NONE);
    else if (a.resample == RS_DOWN) GN_LAUNCH(__half, __half, RS_DOWN);
    else GN_LAUNCH(__half, __half, RS_UP);
  } else if (a.out_half) {
    if (a.resample == RS_NONE) GN_LAUNCH(__half, float, RS_NONE);
    else if (a.resample == RS_DOWN) GN_LAUNCH(__half, float, RS_DOWN);
    else GN_LAUNCH(__half, float, RS_UP);
  } else {
    if (a.resample == RS_NONE) GN_LAUNCH(float, float, RS_NONE);
    else if (a.resample == RS_DOWN) GN_LAUNCH(float, float, RS_DOWN);
    else GN_LAUNCH(float, float, RS_UP);
  }
#undef GN_LAUNCH
#undef GN_LAUNCH_S
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== GroupNorm coefficients
// For the v4 convolution (conv_tc.cu) GroupNorm(+FiLM) is folded into a per-(image, channel) affine map
//   y = ((x - mean) * rstd * gamma + beta) * (1 + film_scale) + film_shift = scale * x + offset       (nn.py:17-19, unet.py:246-249)
// and SiLU(y) = h + h * tanh(h) with h = y / 2, so the table holds hs = scale / 2 and ho = offset / 2: coef[b][0][c], coef[b][1][c].
__global__ void k_gn_coef(GnCoefArgs a) {
  __shared__ float s_mean[32], s_rstd[32];
  const int C = a.C1 + a.C2, C4 = C / 4, b = blockIdx.x;
  if (threadIdx.x < 32) {
    const int qpg = C4 / 32, Q1 = a.C1 / 4;
    double su = 0.0, sq = 0.0;
    for (int j = 0; j < qpg; ++j) {
      const int q = threadIdx.x * qpg + j;
      const double *src = q < Q1 ? a.q1 + ((int64_t)b * Q1 + q) * 2 : a.q2 + ((int64_t)b * (C4 - Q1) + (q - Q1)) * 2;
      su += src[0];
      sq += src[1];
    }
    const double n = (double)a.HW * (double)(C / 32);
    const double mean = su / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  const float *f = a.film ? a.film + (a.tsteps ? a.tsteps[b] : (int64_t)b) * a.film_stride : nullptr;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / (C / 32);
    float sc = s_rstd[g] * a.gamma[c];
    float of = a.beta[c] - s_mean[g] * sc;
    if (f) { const float m = 1.f + f[c]; sc *= m; of = of * m + f[C + c]; }
    a.coef[((int64_t)b * 2) * C + c] = 0.5f * sc;
    a.coef[((int64_t)b * 2 + 1) * C + c] = 0.5f * of;
  }
}

int launch_gn_coef(const GnCoefArgs &a, cudaStream_t st) {
  const int C = a.C1 + a.C2;
  DPIR_REQUIRE(C % 128 == 0 && a.C1 % 4 == 0, "gn_coef: channels must be a multiple of 128 (got %d)", C);
  k_gn_coef<<<a.B, 256, 0, st>>>(a);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== fp32 implicit-GEMM conv
// D[pixel][co] = sum_{tap, ci} act[pixel + tap][ci] * w[tap][ci][co]  on CUDA cores (parity mode / debugging aid).
// Tile 64 pixels x 64 couts x 16 channels, 256 threads, 4x4 register micro-tile.
constexpr int SB_M = 64, SB_N = 64, SB_K = 16, SB_PAD = 4;

__global__ void __launch_bounds__(256) k_conv_simt(const float *__restrict__ act, const float *__restrict__ w, int ldw,
                                                   ConvEpi epi, int B, int H, int W, int Cin, int Cout, int ksize) {
  __shared__ __align__(16) float As[SB_K][SB_M + SB_PAD];
  __shared__ __align__(16) float Bs[SB_K][SB_N];
  int tid = threadIdx.x;
  int64_t npix = (int64_t)B * H * W;
  int64_t m0 = (int64_t)blockIdx.x * SB_M;
  int n0 = blockIdx.y * SB_N;
  // A-load role: pixel row = tid/4, channel quad = tid%4
  int ar = tid >> 2, aq = tid & 3;
  int64_t ap = m0 + ar;
  bool apv = ap < npix;
  int ax = 0, ay = 0, an = 0;
  if (apv) { ax = (int)(ap % W); ay = (int)((ap / W) % H); an = (int)(ap / ((int64_t)W * H)); }
  // B-load role: k row = tid/16, col quad = tid%16
  int br = tid >> 4, bq = tid & 15;
  int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  int taps = ksize * ksize, pad = ksize / 2;
  for (int tap = 0; tap < taps; ++tap) {
    int dy = tap / ksize - pad, dx = tap % ksize - pad;
    int yy = ay + dy, xx = ax + dx;
    bool v = apv && yy >= 0 && yy < H && xx >= 0 && xx < W;
    const float *arow = act + (((int64_t)an * H + yy) * W + xx) * Cin;
    const float *wt = w + (int64_t)tap * Cin * ldw;
    for (int c0 = 0; c0 < Cin; c0 += SB_K) {
      float4 av = v ? *reinterpret_cast<const float4 *>(arow + c0 + aq * 4) : make_float4(0, 0, 0, 0);
      float4 bv = make_float4(0, 0, 0, 0);
      if (n0 + bq * 4 < ldw) bv = *reinterpret_cast<const float4 *>(wt + (int64_t)(c0 + br) * ldw + n0 + bq * 4);
      __syncthreads();
      As[aq * 4 + 0][ar] = av.x; As[aq * 4 + 1][ar] = av.y; As[aq * 4 + 2][ar] = av.z; As[aq * 4 + 3][ar] = av.w;
      *reinterpret_cast<float4 *>(&Bs[br][bq * 4]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SB_K; ++k) {
        float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
        float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
        float a_[4] = {a4.x, a4.y, a4.z, a4.w}, b_[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += a_[i] * b_[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t p = m0 + ty * 4 + i;
    if (p >= npix) continue;
    int x = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    int co = n0 + tx * 4;
    int ncols = Cout - co;
    if (ncols <= 0) continue;
    if (ncols > 4) ncols = 4;
    conv_epilogue_store(epi, n, y, x, H, W, Cout, co, acc[i], ncols);
  }
}

int launch_conv_simt(const float *act, const float *w, int ldw, const ConvEpi &epi, int B, int H, int W, int Cin, int Cout,
                     int ksize, cudaStream_t st) {
  DPIR_REQUIRE(Cin % SB_K == 0 && ldw % 4 == 0 && ldw >= Cout, "conv_simt: Cin %% 16 == 0 and padded Cout %% 4 == 0 required");
  DPIR_REQUIRE(ksize == 1 || ksize == 3, "conv_simt: ksize must be 1 or 3");
  int64_t npix = (int64_t)B * H * W;
  dim3 grid(cdiv(npix, SB_M), cdiv(Cout, SB_N));
  k_conv_simt<<<grid, 256, 0, st>>>(act, w, ldw, epi, B, H, W, Cin, Cout, ksize);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== attention
// softmax((q*s)^T (k*s)) v per (sample, head), s = ch^-1/4, ch = 64  [unet.py:337-354].  Flash-style: 64-query tile per
// CTA, streams 64-key tiles, online softmax in fp32; nothing of size T x T is materialised.
constexpr int AT_D = 64, AT_Q = 64, AT_K = 64, AT_LD = 68;

template <typename T>
__global__ void __launch_bounds__(256) k_attention(const float *__restrict__ qkv, T *__restrict__ out, int Tn, int C, int heads,
                                                   int new_order) {
  extern __shared__ __align__(16) float sm[];
  float *Qt = sm;                       // [D][AT_LD]  (d-major: Qt[d][q])
  float *Kt = Qt + AT_D * AT_LD;        // [D][AT_LD]
  float *Vs = Kt + AT_D * AT_LD;        // [K][AT_LD]  (Vs[k][c])
  float *Pt = Vs + AT_K * AT_LD;        // [K][AT_LD]  (Pt[k][q])
  int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  int q0 = blockIdx.x * AT_Q;
  int h = blockIdx.y, b = blockIdx.z;
  int C3 = 3 * C;
  int qoff, koff, voff;
  if (new_order) { qoff = h * AT_D; koff = C + h * AT_D; voff = 2 * C + h * AT_D; }
  else           { qoff = h * 3 * AT_D; koff = qoff + AT_D; voff = qoff + 2 * AT_D; }
  const float *base = qkv + (int64_t)b * Tn * C3;
  // load Q tile transposed: thread loads float4 along d for row r
  {
    int r = tid >> 2, qd = tid & 3;   // 64 rows x 4 quads of 16 -> loop 4 times over d quads
    for (int it = 0; it < 4; ++it) {
      int d = (qd * 4 + it) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + r < Tn) v = *reinterpret_cast<const float4 *>(base + (int64_t)(q0 + r) * C3 + qoff + d);   // rows >= T: zeros, never stored
      Qt[(d + 0) * AT_LD + r] = v.x; Qt[(d + 1) * AT_LD + r] = v.y; Qt[(d + 2) * AT_LD + r] = v.z; Qt[(d + 3) * AT_LD + r] = v.w;
    }
  }
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  const float scale2 = 0.125f;  // (64^-1/4)^2
  for (int k0 = 0; k0 < Tn; k0 += AT_K) {
    __syncthreads();   // previous tile fully consumed (also orders the Q store on the first pass)
    {
      int r = tid >> 2, qd = tid & 3;
      for (int it = 0; it < 4; ++it) {
        int d = (qd * 4 + it) * 4;
        const float *rowp = base + (int64_t)(k0 + r) * C3;
        const bool kin = k0 + r < Tn;                                    // keys >= T: zero operands, masked scores
        float4 kv = kin ? *reinterpret_cast<const float4 *>(rowp + koff + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        Kt[(d + 0) * AT_LD + r] = kv.x; Kt[(d + 1) * AT_LD + r] = kv.y; Kt[(d + 2) * AT_LD + r] = kv.z; Kt[(d + 3) * AT_LD + r] = kv.w;
        float4 vv = kin ? *reinterpret_cast<const float4 *>(rowp + voff + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4 *>(&Vs[r * AT_LD + d]) = vv;
      }
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < AT_D; ++d) {
      float4 a4 = *reinterpret_cast<const float4 *>(&Qt[d * AT_LD + ty * 4]);
      float4 b4 = *reinterpret_cast<const float4 *>(&Kt[d * AT_LD + tx * 4]);
      float a_[4] = {a4.x, a4.y, a4.z, a4.w}, b_[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] += a_[i] * b_[j];
    }
    // online softmax over this key tile; the 16 tx-lanes of a row share it (xor-shuffles stay inside the 16-lane group)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx * 4 + j < Tn) ? s[i][j] * scale2 : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float m_new = fmaxf(m_run[i], mx);
      float corr = expf(m_run[i] - m_new);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[i][j] = expf(s[i][j] - m_new); ps += s[i][j]; }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l_run[i] = l_run[i] * corr + ps;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
#pragma unroll
      for (int j = 0; j < 4; ++j) Pt[(tx * 4 + j) * AT_LD + ty * 4 + i] = s[i][j];
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < AT_K; ++k) {
      float4 p4 = *reinterpret_cast<const float4 *>(&Pt[k * AT_LD + ty * 4]);
      float4 v4 = *reinterpret_cast<const float4 *>(&Vs[k * AT_LD + tx * 4]);
      float p_[4] = {p4.x, p4.y, p4.z, p4.w}, v_[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] += p_[i] * v_[j];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float inv = 1.f / l_run[i];
    float4 v = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    if (q0 + ty * 4 + i < Tn) Store4<T>::st(out + ((int64_t)b * Tn + q0 + ty * 4 + i) * C + h * AT_D + tx * 4, v);
  }
}

// fp32 CUDA-core attention (fp32 parity mode and the fp16x3 split-precision mode): qkv fp32 [B][T][3C] -> out fp32 / fp16 [B][T][C];
// any T.  The tensor-core pipelines use the tcgen05 kernel of attention_tc.cu on the fp16 qkv tensor (launch_attention_tc).
int launch_attention(const float *qkv, void *out, int out_half, int B, int T, int C, int heads, int new_order, cudaStream_t st) {
  DPIR_REQUIRE(C == heads * AT_D, "attention: head dim must be 64 (C=%d heads=%d)", C, heads);
  DPIR_REQUIRE(T >= 1, "attention: empty sequence");
  size_t smem = (size_t)4 * AT_D * AT_LD * sizeof(float);
  dim3 grid(cdiv(T, AT_Q), heads, B);
  static bool attr_set[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_attention<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_attention<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev] = true;
  }
  if (out_half) k_attention<__half><<<grid, 256, smem, st>>>(qkv, (__half *)out, T, C, heads, new_order);
  else k_attention<float><<<grid, 256, smem, st>>>(qkv, (float *)out, T, C, heads, new_order);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// fp32 -> fp16 elementwise (test entry of the tcgen05 attention: the engine's qkv convolution writes fp16 itself)
__global__ void k_cast_half(const float *__restrict__ in, __half *__restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    Store4<__half>::st(out + 4 * i, *reinterpret_cast<const float4 *>(in + 4 * i));
}
int launch_cast_half(const float *in, __half *out, int64_t n, cudaStream_t st) {
  DPIR_REQUIRE(n % 4 == 0, "cast_half: element count must be a multiple of 4");
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_cast_half<<<(int)blocks, 256, 0, st>>>(in, out, n / 4);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== split cast
// fp32 [npix][C] -> fp16 [npix][hi(C) | lo(C)]: the operand form of the split-precision convolutions (attention output -> proj)
__global__ void k_split_cast(const float *__restrict__ in, __half *__restrict__ out, int64_t n4, int C) {
  const int C4 = C / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / C4;
    const int c = (int)(i - pix * C4) * 4;
    store_act4<__half, true>(out, pix, C, c, *reinterpret_cast<const float4 *>(in + pix * C + c));
  }
}
int launch_split_cast(const float *in, __half *out, int64_t npix, int C, cudaStream_t st) {
  DPIR_REQUIRE(C % 4 == 0, "split_cast: channels must be a multiple of 4");
  const int64_t n4 = npix * (C / 4);
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_split_cast<<<(int)blocks, 256, 0, st>>>(in, out, n4, C);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== timestep clamp
// Device-side timesteps index the per-timestep FiLM table and the eps->x0 coefficient table; a value outside [0, T) would read
// out of bounds (the reference raises IndexError, which needs a host read).  The engine works on a clamped copy.
__global__ void k_clamp_t(const int64_t *__restrict__ t, int64_t *__restrict__ out, int B, int64_t T) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { int64_t v = t[i]; out[i] = v < 0 ? 0 : (v >= T ? T - 1 : v); }
}
int launch_clamp_t(const int64_t *t, int64_t *out, int B, int T, cudaStream_t st) {
  k_clamp_t<<<cdiv(B, 128), 128, 0, st>>>(t, out, B, (int64_t)T);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

// =========================================================================================== timestep embedding MLPs
// hidden[b][o] = sum_i e(t_b)[i] * w0[o][i] + b0[o],  e = [cos(t f_j) | sin(t f_j)], f_j = freqs[j] (nn.py:112-118; the
// table is computed once on the host with the reference's formula and uploaded).
__global__ void k_time_embed0(const int64_t *__restrict__ t, const float *__restrict__ freqs, const float *__restrict__ w0,
                              const float *__restrict__ b0, float *__restrict__ hidden, int mc, int ted) {
  int b = blockIdx.y;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ted) return;
  float tv = (float)t[b];
  int half = mc / 2;
  float acc = 0.f;
  for (int i = lane; i < 2 * half; i += 32) {
    int j = i < half ? i : i - half;
    float a = tv * freqs[j];
    float e = i < half ? cosf(a) : sinf(a);
    acc += e * w0[(int64_t)warp * mc + i];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) hidden[(int64_t)b * ted + warp] = acc + b0[warp];
}

// y[b][o] = act_out(sum_i act_in(x[b][i]) W[o][i] + bias[o]); one warp per output row, the weight row is read once and
// reused for every sample.
__global__ void k_linear_rows(const float *__restrict__ x, const float *__restrict__ Wm, const float *__restrict__ bias,
                              float *__restrict__ y, int B, int in_dim, int rows, int silu_in, int silu_out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float *wr = Wm + (int64_t)warp * in_dim;
  float bz = bias[warp];
  for (int b = 0; b < B; ++b) {
    const float *xb = x + (int64_t)b * in_dim;
    float acc = 0.f;
    for (int i = lane; i < in_dim; i += 32) {
      float v = xb[i];
      if (silu_in) v = v / (1.f + expf(-v));
      acc += v * wr[i];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      float o = acc + bz;
      if (silu_out) o = o / (1.f + expf(-o));
      y[(int64_t)b * rows + warp] = o;
    }
  }
}

// emb_act = SiLU(time_embed(timestep_embedding(t)))  - every consumer (ResBlock.emb_layers, unet.py:199-205) starts with SiLU
int launch_time_embed(const int64_t *t, const float *freqs, const float *w0, const float *b0, const float *w2, const float *b2,
                      float *hidden, float *emb_act, int B, int mc, int ted, cudaStream_t st) {
  dim3 grid(cdiv((int64_t)ted * 32, 256), B);
  k_time_embed0<<<grid, 256, 0, st>>>(t, freqs, w0, b0, hidden, mc, ted);
  DPIR_LAUNCH_CHECK();
  return launch_linear_rows(hidden, w2, b2, emb_act, B, ted, ted, 1, 1, st);
}

int launch_linear_rows(const float *x, const float *W, const float *bias, float *y, int B, int in_dim, int rows, int silu_in,
                       int silu_out, cudaStream_t st) {
  k_linear_rows<<<cdiv((int64_t)rows * 32, 256), 256, 0, st>>>(x, W, bias, y, B, in_dim, rows, silu_in, silu_out);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

}  // namespace dpir
