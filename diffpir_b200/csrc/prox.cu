// prox.cu - HBM-bound kernels of the DiffPIR data-fidelity / re-noise steps, and the cuFFT plumbing.
//
// Reference behaviour (paths relative to the reference root):
//   utils/utils_sisr.py:22-41 p2o, :44-52 upsample, :78-95 pre_calculate, :65-75 data_solution, :9-19 splits
//   main_ddpir.py:393-394 inpaint prox, :396-400 deblur/SR prox wrapper, :403-406 IBP, :451-456 re-noise
//   guided_diffusion/gaussian_diffusion.py:328-333 eps->x0, :305-307 clamp
//   utils/utils_resizer.py:55-74 Resizer.forward;  utils/utils_deblur.py:641-727 blur operator forward
// All kernels are coalesced float4 / float2 streams; spectra are Hermitian half-spectra (H, W/2+1).
#include "common.cuh"
#include <cufft.h>

namespace dpir {

thread_local std::string g_err;
thread_local int64_t g_launch_count = 0;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

// ------------------------------------------------------------------------------------------ elementwise
__global__ void k_affine(const float *__restrict__ x, float a, float b, float *__restrict__ out, int64_t n4, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n4; j += stride) {
    float4 v = reinterpret_cast<const float4 *>(x)[j];
    v.x = a * v.x + b; v.y = a * v.y + b; v.z = a * v.z + b; v.w = a * v.w + b;
    reinterpret_cast<float4 *>(out)[j] = v;
  }
  for (int64_t j = n4 * 4 + i; j < n; j += stride) out[j] = a * x[j] + b;
}

// (mask*(ys*y+yb) + tau*x0)/(mask+tau), then x0 + g*(x0_p - x0)   [main_ddpir.py:393-394]
__device__ __forceinline__ float inpaint1(float x0, float y, float m, float tau, float ys, float yb, float g) {
  float x0p = (m * (ys * y + yb) + tau * x0) / (m + tau);
  return x0 + g * (x0p - x0);
}
__global__ void k_inpaint_prox(const float *__restrict__ x0, const float *__restrict__ y, const float *__restrict__ mask,
                               float tau, const float *__restrict__ tau_p, float ys, float yb, float g, float *__restrict__ out,
                               int64_t n4, int64_t n) {
  if (tau_p) tau = __ldg(tau_p);      // device-resident scalar (the drivers keep rhos[t] on the GPU, main_ddpir.py:389): no host sync
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n4; j += stride) {
    float4 a = reinterpret_cast<const float4 *>(x0)[j];
    float4 b = reinterpret_cast<const float4 *>(y)[j];
    float4 m = reinterpret_cast<const float4 *>(mask)[j];
    float4 o;
    o.x = inpaint1(a.x, b.x, m.x, tau, ys, yb, g);
    o.y = inpaint1(a.y, b.y, m.y, tau, ys, yb, g);
    o.z = inpaint1(a.z, b.z, m.z, tau, ys, yb, g);
    o.w = inpaint1(a.w, b.w, m.w, tau, ys, yb, g);
    reinterpret_cast<float4 *>(out)[j] = o;
  }
  for (int64_t j = n4 * 4 + i; j < n; j += stride) out[j] = inpaint1(x0[j], y[j], mask[j], tau, ys, yb, g);
}

// eps = (x - sa_t x0)/s1_t ; out = sa_p x0 + c_eps eps + c_eta n1 + c_zeta n2   [main_ddpir.py:451-456]
__device__ __forceinline__ float renoise1(float x, float x0, float n1, float n2, float sa_t, float s1_t, float sa_p,
                                          float c_eps, float c_eta, float c_zeta) {
  float eps = (x - sa_t * x0) / s1_t;
  return sa_p * x0 + (c_eps * eps + c_eta * n1) + c_zeta * n2;
}
__global__ void k_renoise(const float *__restrict__ x, const float *__restrict__ x0, const float *__restrict__ n1,
                          const float *__restrict__ n2, float sa_t, float s1_t, float sa_p, float c_eps, float c_eta,
                          float c_zeta, float *__restrict__ out, int64_t n4, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < n4; j += stride) {
    float4 a = reinterpret_cast<const float4 *>(x)[j];
    float4 b = reinterpret_cast<const float4 *>(x0)[j];
    float4 e = n1 ? reinterpret_cast<const float4 *>(n1)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 z = reinterpret_cast<const float4 *>(n2)[j];
    float4 o;
    o.x = renoise1(a.x, b.x, e.x, z.x, sa_t, s1_t, sa_p, c_eps, c_eta, c_zeta);
    o.y = renoise1(a.y, b.y, e.y, z.y, sa_t, s1_t, sa_p, c_eps, c_eta, c_zeta);
    o.z = renoise1(a.z, b.z, e.z, z.z, sa_t, s1_t, sa_p, c_eps, c_eta, c_zeta);
    o.w = renoise1(a.w, b.w, e.w, z.w, sa_t, s1_t, sa_p, c_eps, c_eta, c_zeta);
    reinterpret_cast<float4 *>(out)[j] = o;
  }
  for (int64_t j = n4 * 4 + i; j < n; j += stride)
    out[j] = renoise1(x[j], x0[j], n1 ? n1[j] : 0.f, n2[j], sa_t, s1_t, sa_p, c_eps, c_eta, c_zeta);
}

// x0 = clamp(c1 x - c2 eps)   [gaussian_diffusion.py:328-333, 305-307]
__global__ void k_eps_to_x0(const float *__restrict__ mo, const float *__restrict__ x, const float *__restrict__ coef,
                            float *__restrict__ x0, int B, int Cout, int HW, int clip) {
  int64_t total = (int64_t)B * 3 * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / (3 * (int64_t)HW));
    int64_t r = i - (int64_t)b * 3 * HW;
    float c1 = coef[2 * b], c2 = coef[2 * b + 1];
    float v = c1 * x[i] - c2 * mo[(int64_t)b * Cout * HW + r];
    if (clip) v = fminf(fmaxf(v, -1.f), 1.f);
    x0[i] = v;
  }
}

// p_sample / ddim_sample tails (gaussian_diffusion.py:258-311 + 430-438, and :566-585) as ONE pass over the UNet output:
//   eps = mo[:, :3], v = mo[:, 3:6];  x0 = clamp(c1 x - c2 eps)
//   mode 0 (p_sample, LEARNED_RANGE):  mean = pm1 x0 + pm2 x;  logvar = f max_log + (1-f) min_log, f = (v+1)/2;
//                                      sample = mean + [t != 0] exp(0.5 logvar) noise
//   mode 1 (ddim_sample):              eps' = (c1 x - x0)/c2;  sample = sqrt_ac_prev x0 + ce eps' + [t != 0] sigma noise
// tab[t] = (c1, c2, pm1, pm2, min_log, max_log, sqrt_ac_prev, ce, sigma, -) as float (float64 tables cast like
// _extract_into_tensor :905).  Either output may be null.
__global__ void k_sample_step(const float *__restrict__ mo, const float *__restrict__ x, const float *__restrict__ noise,
                              const int64_t *__restrict__ t, const float *__restrict__ tab, int T, float *__restrict__ sample,
                              float *__restrict__ x0o, int B, int Cout, int HW, int mode, int clip) {
  int64_t total = (int64_t)B * 3 * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int b = (int)(i / (3 * (int64_t)HW));
    int64_t r = i - (int64_t)b * 3 * HW;
    int64_t tt = t[b];
    if (tt < 0 || tt >= T) __trap();                 // the reference raises IndexError
    const float *c = tab + tt * 10;
    const float xv = x[i];
    const float eps = mo[(int64_t)b * Cout * HW + r];
    float v0 = c[0] * xv - c[1] * eps;
    if (clip) v0 = fminf(fmaxf(v0, -1.f), 1.f);
    if (x0o) x0o[i] = v0;
    if (sample) {
      const float nz = (noise && tt != 0) ? noise[i] : 0.f;
      float sv;
      if (mode == 0) {
        const float mean = c[2] * v0 + c[3] * xv;
        const float f = (mo[(int64_t)b * Cout * HW + (int64_t)3 * HW + r] + 1.f) / 2.f;
        const float logvar = f * c[5] + (1.f - f) * c[4];
        sv = mean + expf(0.5f * logvar) * nz;
      } else {
        const float epsd = (c[0] * xv - v0) / c[1];
        sv = v0 * c[6] + c[7] * epsd + c[8] * nz;
      }
      sample[i] = sv;
    }
  }
}

// ------------------------------------------------------------------------------------------ FFT helpers
// psf -> zero-padded, centre rolled to (0,0)   [utils_sisr.py:34-37]
__global__ void k_psf_place(const float *__restrict__ k, float *__restrict__ otf_in, int Bk, int kh, int kw, int H, int W) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int total = Bk * kh * kw;
  if (idx >= total) return;
  int b = idx / (kh * kw);
  int r = idx - b * kh * kw;
  int i = r / kw, j = r - i * kw;
  int oi = ((i - kh / 2) % H + H) % H;
  int oj = ((j - kw / 2) % W + W) % W;
  // several taps can land on one cell only when the psf is larger than the image; accumulate like copy_+roll would not
  otf_in[((int64_t)b * H + oi) * W + oj] = k[idx];
}

// F2B = |FB|^2   [utils_sisr.py:91]
__global__ void k_abs2(const float2 *__restrict__ FB, float *__restrict__ F2B, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 v = FB[i];
    F2B[i] = v.x * v.x + v.y * v.y;
  }
}

// STy = zero-stuffed y   [utils_sisr.py:44-52]
__global__ void k_zero_stuff(const float *__restrict__ y, float *__restrict__ out, int planes, int Hl, int Wl, int sf) {
  int H = Hl * sf, W = Wl * sf;
  int64_t total = (int64_t)planes * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    int h = (int)((i / W) % H);
    int64_t p = i / ((int64_t)W * H);
    float v = 0.f;
    if (h % sf == 0 && w % sf == 0) v = y[(p * Hl + h / sf) * Wl + w / sf];
    out[i] = v;
  }
}

// FBFy = conj(FB) * Fy (in place on Fy)   [utils_sisr.py:93-94]
__global__ void k_mul_conj(const float2 *__restrict__ FB, float2 *__restrict__ Fy, int B, int C, int Bk, int64_t plane) {
  int64_t total = (int64_t)B * C * plane;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t f = i % plane;
    int b = (int)(i / (plane * C));
    float2 k = FB[(int64_t)(Bk == 1 ? 0 : b) * plane + f];
    float2 v = Fy[i];
    Fy[i] = make_float2(k.x * v.x + k.y * v.y, k.x * v.y - k.y * v.x);
  }
}

// sf == 1: FX = (alpha Xi + FBFy) / (F2B + alpha), Xi = in_scale X + in_bias HW delta, then the affine post step;
// in place on spec.  Algorithmic bytes per image (3x256x256): X r/w folded into the cuFFT passes; here
// spec read+write (2 x 792,576) + FBFy 792,576 + F2B 132,096.
__global__ void k_solve_sf1(float2 *__restrict__ spec, const float *__restrict__ F2B, const float2 *__restrict__ FBFy,
                            float alpha, const float *__restrict__ alpha_p, int B, int C, int Bk, int64_t plane, float HW,
                            float in_scale, float in_bias, float out_scale, float out_bias, float blend) {
  if (alpha_p) alpha = __ldg(alpha_p);
  int64_t total = (int64_t)B * C * plane;
  float inv = 1.0f / HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t f = i % plane;
    int b = (int)(i / (plane * C));
    float2 X = spec[i];
    float2 Xi = make_float2(in_scale * X.x, in_scale * X.y);
    if (f == 0) Xi.x += in_bias * HW;
    float w = F2B[(int64_t)(Bk == 1 ? 0 : b) * plane + f];
    float2 y = FBFy[i];
    float d = 1.0f / (w + alpha);
    float2 FX = make_float2((alpha * Xi.x + y.x) * d, (alpha * Xi.y + y.y) * d);
    float2 o = make_float2(blend * out_scale * FX.x + (1.f - blend) * X.x, blend * out_scale * FX.y + (1.f - blend) * X.y);
    if (f == 0) o.x += blend * out_bias * HW;
    spec[i] = make_float2(o.x * inv, o.y * inv);
  }
}

__device__ __forceinline__ float2 herm_load(const float2 *__restrict__ p, int r, int c, int H, int W, int Wh) {
  if (c < Wh) return p[(int64_t)r * Wh + c];
  float2 v = p[(int64_t)((H - r) % H) * Wh + (W - c)];
  return make_float2(v.x, -v.y);
}

// sf > 1: one thread per low-resolution frequency (u,v); gathers the sf^2 aliases (Hermitian mirror for columns > W/2)
//   S = mean_i FB_i Xi_i,  Wm = mean_i |FB_i|^2,  FX_j = Xi_j + (FBFy_j - conj(FB_j) S)/(Wm + alpha)
// Equivalent to utils_sisr.py:66-72 but without the (FR - ...)/alpha cancellation.  Out-of-place.
__global__ void k_solve_sf(const float2 *__restrict__ spec, float2 *__restrict__ spec_out, const float2 *__restrict__ FB,
                           const float2 *__restrict__ FBFy, float alpha, const float *__restrict__ alpha_p, int B, int C, int Bk,
                           int H, int W, int sf, float in_scale, float in_bias, float out_scale, float out_bias, float blend) {
  if (alpha_p) alpha = __ldg(alpha_p);
  int Hl = H / sf, Wl = W / sf, Wh = W / 2 + 1;
  int64_t total = (int64_t)B * C * Hl * Wl;
  float HWf = (float)H * (float)W;
  float inv = 1.0f / HWf;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int v = (int)(i % Wl);
    int u = (int)((i / Wl) % Hl);
    int64_t p = i / ((int64_t)Wl * Hl);
    int b = (int)(p / C);
    const float2 *X = spec + p * (int64_t)H * Wh;
    const float2 *Bf = FB + (int64_t)(Bk == 1 ? 0 : b) * H * Wh;
    const float2 *Yf = FBFy + p * (int64_t)H * Wh;
    float2 *O = spec_out + p * (int64_t)H * Wh;
    float2 S = make_float2(0.f, 0.f);
    float Wm = 0.f;
    for (int a = 0; a < sf; ++a)
      for (int c = 0; c < sf; ++c) {
        int r = u + a * Hl, cc = v + c * Wl;
        float2 xi = herm_load(X, r, cc, H, W, Wh);
        xi.x *= in_scale; xi.y *= in_scale;
        if (r == 0 && cc == 0) xi.x += in_bias * HWf;
        float2 k = herm_load(Bf, r, cc, H, W, Wh);
        S.x += k.x * xi.x - k.y * xi.y;
        S.y += k.x * xi.y + k.y * xi.x;
        Wm += k.x * k.x + k.y * k.y;
      }
    float n = 1.0f / (float)(sf * sf);
    S.x *= n; S.y *= n; Wm *= n;
    float d = 1.0f / (Wm + alpha);
    for (int a = 0; a < sf; ++a)
      for (int c = 0; c < sf; ++c) {
        int r = u + a * Hl, cc = v + c * Wl;
        if (cc >= Wh) continue;
        float2 x = X[(int64_t)r * Wh + cc];
        float2 xi = make_float2(in_scale * x.x, in_scale * x.y);
        if (r == 0 && cc == 0) xi.x += in_bias * HWf;
        float2 k = Bf[(int64_t)r * Wh + cc];
        float2 y = Yf[(int64_t)r * Wh + cc];
        // conj(k) * S
        float2 kS = make_float2(k.x * S.x + k.y * S.y, k.x * S.y - k.y * S.x);
        float2 FX = make_float2(xi.x + (y.x - kS.x) * d, xi.y + (y.y - kS.y) * d);
        float2 o = make_float2(blend * out_scale * FX.x + (1.f - blend) * x.x, blend * out_scale * FX.y + (1.f - blend) * x.y);
        if (r == 0 && cc == 0) o.x += blend * out_bias * HWf;
        O[(int64_t)r * Wh + cc] = make_float2(o.x * inv, o.y * inv);
      }
  }
}

// ------------------------------------------------------------------------------------------ resize / IBP / blur
// one pass of Resizer.forward along H (dim_is_h=1) or W: out[o] = sum_t w[t][o] * in[fov[t][o]]
__global__ void k_resize_pass(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ w,
                              const int *__restrict__ fov, int taps, int N, int Hin, int Win, int Hout, int Wout, int along_h) {
  int64_t total = (int64_t)N * Hout * Wout;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % Wout);
    int y = (int)((i / Wout) % Hout);
    int64_t n = i / ((int64_t)Wout * Hout);
    const float *src = in + n * (int64_t)Hin * Win;
    float acc = 0.f;
    if (along_h) {
      for (int t = 0; t < taps; ++t) acc += src[(int64_t)fov[t * Hout + y] * Win + x] * w[t * Hout + y];
    } else {
      for (int t = 0; t < taps; ++t) acc += src[(int64_t)y * Win + fov[t * Wout + x]] * w[t * Wout + x];
    }
    out[i] = acc;
  }
}

// out = 2*((x0/2+.5) + gamma * up_nearest(y - down)/(1+rho)) - 1    [main_ddpir.py:403-406]
__global__ void k_ibp(const float *__restrict__ x0, const float *__restrict__ y, const float *__restrict__ down, float gamma,
                      float rho, const float *__restrict__ rho_p, int N, int H, int W, int sf, float *__restrict__ out) {
  if (rho_p) rho = __ldg(rho_p);
  int Hl = H / sf, Wl = W / sf;
  int64_t total = (int64_t)N * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    int yy = (int)((i / W) % H);
    int64_t n = i / ((int64_t)W * H);
    int64_t l = (n * Hl + yy / sf) * Wl + x / sf;
    float v = x0[i] / 2 + 0.5f;
    v = v + gamma * (y[l] - down[l]) / (1 + rho);
    out[i] = v * 2 - 1;
  }
}

// ReflectionPad2d(k//2) + depthwise correlation (nn.Conv2d)   [utils_deblur.py:647-655]
__global__ void k_blur_reflect(const float *__restrict__ x, const float *__restrict__ k, int planes, int H, int W, int kh,
                               int kw, float *__restrict__ out) {
  extern __shared__ float sk[];
  for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) sk[i] = k[i];
  __syncthreads();
  int64_t total = (int64_t)planes * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int xx = (int)(i % W);
    int yy = (int)((i / W) % H);
    const float *src = x + (i / ((int64_t)W * H)) * (int64_t)H * W;
    float acc = 0.f;
    for (int a = 0; a < kh; ++a) {
      int r = yy + a - kh / 2;
      r = r < 0 ? -r : (r >= H ? 2 * H - 2 - r : r);
      for (int b = 0; b < kw; ++b) {
        int c = xx + b - kw / 2;
        c = c < 0 ? -c : (c >= W ? 2 * W - 2 - c : c);
        acc += src[(int64_t)r * W + c] * sk[a * kw + b];
      }
    }
    out[i] = acc;
  }
}


// ------------------------------------------------------------------------------------------ output / metrics tail
// tensor2uint_batch (utils/utils_image.py:238-242): clamp(0,1) * 255, round-half-even (np.round), uint8, NCHW -> NHWC
__global__ void k_tensor2uint(const float *__restrict__ x, uint8_t *__restrict__ out, int B, int C, int H, int W) {
  int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t p = i / C;                    // output index order: (b, h, w, c)
    int w = (int)(p % W);
    int h = (int)((p / W) % H);
    int b = (int)(p / ((int64_t)W * H));
    float v = x[(((int64_t)b * C + c) * H + h) * W + w];
    v = fminf(fmaxf(v, 0.f), 1.f) * 255.0f;
    out[i] = (uint8_t)__float2int_rn(v);
  }
}

// per-image sum of squared differences of (sa*a+ba) and (sb*b+bb); luma != 0: on Y = 0.299R+0.587G+0.114B only
// (rgb2ycbcr_batch only_y, utils_image.py:470-490).  Block-level warp-shuffle reduction, one fp64 atomic per block.
__global__ void k_sqdiff(const float *__restrict__ a, const float *__restrict__ b, float sa, float ba, float sb, float bb,
                         int luma, int C, int HW, double *__restrict__ sse) {
  int img = blockIdx.y;
  const float *pa = a + (int64_t)img * C * HW, *pb = b + (int64_t)img * C * HW;
  float acc = 0.f;
  if (luma) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      float ya = 0.299f * (sa * pa[i] + ba) + 0.587f * (sa * pa[HW + i] + ba) + 0.114f * (sa * pa[2 * HW + i] + ba);
      float yb = 0.299f * (sb * pb[i] + bb) + 0.587f * (sb * pb[HW + i] + bb) + 0.114f * (sb * pb[2 * HW + i] + bb);
      float d = ya - yb;
      acc += d * d;
    }
  } else {
    int n = C * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      float d = (sa * pa[i] + ba) - (sb * pb[i] + bb);
      acc += d * d;
    }
  }
  __shared__ float red[32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) atomicAdd(&sse[img], (double)v);
  }
}

// measurement synthesis (main_ddpir.py:97-114): circular blur on the uint8 image exactly as
// scipy.ndimage.convolve(img_H, k[..., None], mode='wrap') evaluates it - correlation with the flipped kernel, taps in
// row-major order of the flipped kernel, fp64 multiply and add WITHOUT fma contraction, result truncated to uint8 - then
// /255, *2-1, + noise (host-drawn AWGN, std 2*sigma), /2+.5, written as float NCHW.
__global__ void k_measure_blur_u8(const uint8_t *__restrict__ img, const double *__restrict__ kf, const float *__restrict__ noise,
                                  uint8_t *__restrict__ blur_u8, float *__restrict__ out, int B, int H, int W, int C, int kh,
                                  int kw) {
  extern __shared__ double skf[];
  for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) skf[i] = kf[(int64_t)blockIdx.y * kh * kw + i];
  __syncthreads();
  int b = blockIdx.y;
  int64_t n = (int64_t)H * W * C;
  const uint8_t *src = img + (int64_t)b * n;
  // scipy: convolve == correlate with weights[::-1, ::-1], filter centre at size/2 (odd) or size/2 - 1 (even); taps with
  // |w| <= DBL_EPSILON are dropped from the footprint (NI_Correlate)
  const int sh = (kh & 1) ? kh / 2 : kh / 2 - 1, sw = (kw & 1) ? kw / 2 : kw / 2 - 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int x = (int)((i / C) % W);
    int y = (int)(i / ((int64_t)C * W));
    double acc = 0.0;
    for (int a = 0; a < kh; ++a) {
      int yy = y + a - sh;
      yy %= H; if (yy < 0) yy += H;
      for (int bb = 0; bb < kw; ++bb) {
        double wgt = skf[(kh - 1 - a) * kw + (kw - 1 - bb)];
        if (fabs(wgt) <= 2.220446049250313e-16) continue;
        int xx = x + bb - sw;
        xx %= W; if (xx < 0) xx += W;
        acc = __dadd_rn(acc, __dmul_rn((double)src[((int64_t)yy * W + xx) * C + c], wgt));
      }
    }
    uint8_t q = (uint8_t)(int)acc;          // C cast: truncation
    if (blur_u8) blur_u8[(int64_t)b * n + i] = q;
    if (out) {
      float v = (float)q / 255.0f;
      float nz = noise ? noise[(int64_t)b * n + i] : 0.f;
      v = (v * 2.f - 1.f + nz) / 2.f + 0.5f;
      out[(((int64_t)b * C + c) * H + y) * W + x] = v;
    }
  }
}

static inline int grid_for(int64_t n, int block = 256, int max_blocks = 148 * 16) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace dpir

using namespace dpir;

struct dpir_fft_plan {
  cufftHandle r2c = 0, c2r = 0;
  int batch = 0, H = 0, W = 0;
};

extern "C" {

const char *dpir_last_error(void) { return g_err.c_str(); }
int dpir_version(void) { return 200; }
int64_t dpir_launch_count(void) { return g_launch_count; }
void dpir_launch_count_reset(void) { g_launch_count = 0; }

int dpir_affine(const float *x, float a, float b, float *out, int64_t n, void *stream) {
  DPIR_REQUIRE(x && out && n >= 0, "dpir_affine: null pointer");
  bool al = (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
  int64_t n4 = al ? n / 4 : 0;
  k_affine<<<grid_for(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(x, a, b, out, n4, n);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

static int inpaint_prox_impl(const float *x0, const float *y, const float *mask, float tau, const float *tau_p, float y_scale,
                             float y_bias, float guidance, float *out, int64_t n, void *stream) {
  DPIR_REQUIRE(x0 && y && mask && out, "dpir_inpaint_prox: null pointer");
  bool al = (((uintptr_t)x0 | (uintptr_t)y | (uintptr_t)mask | (uintptr_t)out) & 15) == 0;
  int64_t n4 = al ? n / 4 : 0;
  k_inpaint_prox<<<grid_for(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(x0, y, mask, tau, tau_p, y_scale, y_bias, guidance, out, n4, n);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}
int dpir_inpaint_prox(const float *x0, const float *y, const float *mask, float tau, float y_scale, float y_bias,
                      float guidance, float *out, int64_t n, void *stream) {
  return inpaint_prox_impl(x0, y, mask, tau, nullptr, y_scale, y_bias, guidance, out, n, stream);
}
int dpir_inpaint_prox_dev(const float *x0, const float *y, const float *mask, const float *tau_dev, float y_scale, float y_bias,
                          float guidance, float *out, int64_t n, void *stream) {
  DPIR_REQUIRE(tau_dev, "dpir_inpaint_prox_dev: null tau");
  return inpaint_prox_impl(x0, y, mask, 0.f, tau_dev, y_scale, y_bias, guidance, out, n, stream);
}

int dpir_renoise(const float *x, const float *x0, const float *noise_eta, const float *noise_zeta, float sa_t, float s1_t,
                 float sa_p, float c_eps, float c_eta, float c_zeta, float *out, int64_t n, void *stream) {
  DPIR_REQUIRE(x && x0 && noise_zeta && out, "dpir_renoise: null pointer");
  DPIR_REQUIRE(noise_eta || c_eta == 0.f, "dpir_renoise: noise_eta required when c_eta != 0");
  bool al = (((uintptr_t)x | (uintptr_t)x0 | (uintptr_t)noise_eta | (uintptr_t)noise_zeta | (uintptr_t)out) & 15) == 0;
  int64_t n4 = al ? n / 4 : 0;
  k_renoise<<<grid_for(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(x, x0, noise_eta, noise_zeta, sa_t, s1_t, sa_p, c_eps,
                                                                    c_eta, c_zeta, out, n4, n);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

int dpir_eps_to_x0(const float *model_out, const float *x, const float *coef, float *x0, int32_t B, int32_t Cout,
                   int32_t HW, int32_t clip, void *stream) {
  DPIR_REQUIRE(model_out && x && coef && x0 && Cout >= 3, "dpir_eps_to_x0: bad arguments");
  k_eps_to_x0<<<grid_for((int64_t)B * 3 * HW), 256, 0, (cudaStream_t)stream>>>(model_out, x, coef, x0, B, Cout, HW, clip);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

int dpir_sample_step(const float *model_out, const float *x, const float *noise, const int64_t *t, const float *tab, int32_t T,
                     float *sample, float *x0, int32_t B, int32_t Cout, int32_t HW, int32_t mode, int32_t clip, void *stream) {
  DPIR_REQUIRE(model_out && x && t && tab && (sample || x0), "dpir_sample_step: null pointer");
  DPIR_REQUIRE(mode == 1 || Cout >= 6, "dpir_sample_step: p_sample needs the learned-range variance channels (Cout >= 6)");
  DPIR_REQUIRE(Cout >= 3 && (mode == 0 || mode == 1) && T >= 1, "dpir_sample_step: bad arguments");
  k_sample_step<<<grid_for((int64_t)B * 3 * HW), 256, 0, (cudaStream_t)stream>>>(model_out, x, noise, t, tab, T, sample, x0, B, Cout,
                                                                                  HW, mode, clip);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

int dpir_fft_plan_create(int32_t batch, int32_t H, int32_t W, dpir_fft_plan **out) {
  DPIR_REQUIRE(out && batch > 0 && H > 0 && W > 0 && W % 2 == 0, "dpir_fft_plan_create: bad shape (W must be even)");
  dpir_fft_plan *p = new dpir_fft_plan();
  p->batch = batch; p->H = H; p->W = W;
  int n[2] = {H, W};
  cufftResult r = cufftPlanMany(&p->r2c, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, batch);
  if (r != CUFFT_SUCCESS) { set_error("cufftPlanMany(R2C) failed: %d", (int)r); delete p; return DPIR_E_CUFFT; }
  r = cufftPlanMany(&p->c2r, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, batch);
  if (r != CUFFT_SUCCESS) { set_error("cufftPlanMany(C2R) failed: %d", (int)r); cufftDestroy(p->r2c); delete p; return DPIR_E_CUFFT; }
  *out = p;
  return DPIR_OK;
}

int dpir_fft_plan_destroy(dpir_fft_plan *p) {
  if (!p) return DPIR_OK;
  cufftDestroy(p->r2c);
  cufftDestroy(p->c2r);
  delete p;
  return DPIR_OK;
}

int dpir_rfft2(dpir_fft_plan *p, const float *in, float *out_spec, void *stream) {
  DPIR_REQUIRE(p && in && out_spec, "dpir_rfft2: null pointer");
  cufftResult r = cufftSetStream(p->r2c, (cudaStream_t)stream);
  if (r == CUFFT_SUCCESS) r = cufftExecR2C(p->r2c, const_cast<float *>(in), reinterpret_cast<cufftComplex *>(out_spec));
  if (r != CUFFT_SUCCESS) { set_error("cufftExecR2C failed: %d", (int)r); return DPIR_E_CUFFT; }
  return DPIR_OK;
}

int dpir_irfft2(dpir_fft_plan *p, float *spec, float *out, void *stream) {
  DPIR_REQUIRE(p && spec && out, "dpir_irfft2: null pointer");
  cufftResult r = cufftSetStream(p->c2r, (cudaStream_t)stream);
  if (r == CUFFT_SUCCESS) r = cufftExecC2R(p->c2r, reinterpret_cast<cufftComplex *>(spec), out);
  if (r != CUFFT_SUCCESS) { set_error("cufftExecC2R failed: %d", (int)r); return DPIR_E_CUFFT; }
  return DPIR_OK;
}

int dpir_fft_precalc(dpir_fft_plan *plan_k, dpir_fft_plan *plan_y, const float *y, const float *k, int32_t B, int32_t C,
                     int32_t Bk, int32_t kh, int32_t kw, int32_t sf, float *FB, float *F2B, float *FBFy, float *scratch,
                     void *stream) {
  DPIR_REQUIRE(plan_k && plan_y && y && k && FB && F2B && FBFy && scratch, "dpir_fft_precalc: null pointer");
  DPIR_REQUIRE(Bk == 1 || Bk == B, "dpir_fft_precalc: kernel batch must be 1 or B");
  DPIR_REQUIRE(plan_k->batch == Bk && plan_y->batch == B * C && plan_k->H == plan_y->H && plan_k->W == plan_y->W,
               "dpir_fft_precalc: plan/batch mismatch");
  int H = plan_y->H, W = plan_y->W, Wh = W / 2 + 1;
  DPIR_REQUIRE(sf >= 1 && H % sf == 0 && W % sf == 0, "dpir_fft_precalc: H, W must be divisible by sf");
  DPIR_REQUIRE(kh <= H && kw <= W, "dpir_fft_precalc: psf larger than the image");
  cudaStream_t st = (cudaStream_t)stream;
  DPIR_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)Bk * H * W * sizeof(float), st));
  k_psf_place<<<cdiv((int64_t)Bk * kh * kw, 256), 256, 0, st>>>(k, scratch, Bk, kh, kw, H, W);
  DPIR_LAUNCH_CHECK();
  int rc = dpir_rfft2(plan_k, scratch, FB, stream);
  if (rc) return rc;
  k_abs2<<<grid_for((int64_t)Bk * H * Wh), 256, 0, st>>>(reinterpret_cast<const float2 *>(FB), F2B, (int64_t)Bk * H * Wh);
  DPIR_LAUNCH_CHECK();
  const float *src = y;
  if (sf > 1) {
    k_zero_stuff<<<grid_for((int64_t)B * C * H * W), 256, 0, st>>>(y, scratch, B * C, H / sf, W / sf, sf);
    DPIR_LAUNCH_CHECK();
    src = scratch;
  }
  rc = dpir_rfft2(plan_y, src, FBFy, stream);
  if (rc) return rc;
  k_mul_conj<<<grid_for((int64_t)B * C * H * Wh), 256, 0, st>>>(reinterpret_cast<const float2 *>(FB),
                                                               reinterpret_cast<float2 *>(FBFy), B, C, Bk, (int64_t)H * Wh);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

static int fft_solve_impl(dpir_fft_plan *plan_x, const float *x, const float *FB, const float *F2B, const float *FBFy, float alpha,
                          const float *alpha_p, int32_t B, int32_t C, int32_t Bk, int32_t sf, float in_scale, float in_bias,
                          float out_scale, float out_bias, float blend, float *spec, float *out, void *stream) {
  DPIR_REQUIRE(plan_x && x && FB && F2B && FBFy && spec && out, "dpir_fft_solve: null pointer");
  DPIR_REQUIRE(plan_x->batch == B * C, "dpir_fft_solve: plan batch != B*C");
  DPIR_REQUIRE(Bk == 1 || Bk == B, "dpir_fft_solve: kernel batch must be 1 or B");
  int H = plan_x->H, W = plan_x->W, Wh = W / 2 + 1;
  DPIR_REQUIRE(sf >= 1 && H % sf == 0 && W % sf == 0, "dpir_fft_solve: H, W must be divisible by sf");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = dpir_rfft2(plan_x, x, spec, stream);
  if (rc) return rc;
  int64_t plane = (int64_t)H * Wh;
  float *inv_in = spec;
  if (sf == 1) {
    k_solve_sf1<<<grid_for((int64_t)B * C * plane), 256, 0, st>>>(reinterpret_cast<float2 *>(spec), F2B,
                                                                  reinterpret_cast<const float2 *>(FBFy), alpha, alpha_p, B, C,
                                                                  Bk, plane, (float)H * (float)W, in_scale, in_bias, out_scale,
                                                                  out_bias, blend);
  } else {
    float *spec2 = spec + (size_t)B * C * plane * 2;
    k_solve_sf<<<grid_for((int64_t)B * C * (H / sf) * (W / sf)), 256, 0, st>>>(
        reinterpret_cast<const float2 *>(spec), reinterpret_cast<float2 *>(spec2), reinterpret_cast<const float2 *>(FB),
        reinterpret_cast<const float2 *>(FBFy), alpha, alpha_p, B, C, Bk, H, W, sf, in_scale, in_bias, out_scale, out_bias, blend);
    inv_in = spec2;
  }
  DPIR_LAUNCH_CHECK();
  return dpir_irfft2(plan_x, inv_in, out, stream);
}
int dpir_fft_solve(dpir_fft_plan *plan_x, const float *x, const float *FB, const float *F2B, const float *FBFy, float alpha,
                   int32_t B, int32_t C, int32_t Bk, int32_t sf, float in_scale, float in_bias, float out_scale,
                   float out_bias, float blend, float *spec, float *out, void *stream) {
  return fft_solve_impl(plan_x, x, FB, F2B, FBFy, alpha, nullptr, B, C, Bk, sf, in_scale, in_bias, out_scale, out_bias, blend, spec,
                        out, stream);
}
int dpir_fft_solve_dev(dpir_fft_plan *plan_x, const float *x, const float *FB, const float *F2B, const float *FBFy,
                       const float *alpha_dev, int32_t B, int32_t C, int32_t Bk, int32_t sf, float in_scale, float in_bias,
                       float out_scale, float out_bias, float blend, float *spec, float *out, void *stream) {
  DPIR_REQUIRE(alpha_dev, "dpir_fft_solve_dev: null alpha");
  return fft_solve_impl(plan_x, x, FB, F2B, FBFy, 0.f, alpha_dev, B, C, Bk, sf, in_scale, in_bias, out_scale, out_bias, blend, spec,
                        out, stream);
}

int dpir_resize_separable(const float *in, int32_t N, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, const float *w_h,
                          const int32_t *fov_h, int32_t taps_h, const float *w_w, const int32_t *fov_w, int32_t taps_w,
                          float *tmp, float *out, void *stream) {
  DPIR_REQUIRE(in && w_h && fov_h && w_w && fov_w && tmp && out, "dpir_resize_separable: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  k_resize_pass<<<grid_for((int64_t)N * Hout * Win), 256, 0, st>>>(in, tmp, w_h, fov_h, taps_h, N, Hin, Win, Hout, Win, 1);
  DPIR_LAUNCH_CHECK();
  k_resize_pass<<<grid_for((int64_t)N * Hout * Wout), 256, 0, st>>>(tmp, out, w_w, fov_w, taps_w, N, Hout, Win, Hout, Wout, 0);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

static int ibp_impl(const float *x0, const float *y, const float *down, float gamma, float rho, const float *rho_p, int32_t N,
                    int32_t H, int32_t W, int32_t sf, float *out, void *stream) {
  DPIR_REQUIRE(x0 && y && down && out && sf >= 1 && H % sf == 0 && W % sf == 0, "dpir_ibp_update: bad arguments");
  k_ibp<<<grid_for((int64_t)N * H * W), 256, 0, (cudaStream_t)stream>>>(x0, y, down, gamma, rho, rho_p, N, H, W, sf, out);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}
int dpir_ibp_update(const float *x0, const float *y, const float *down, float gamma, float rho, int32_t N, int32_t H,
                    int32_t W, int32_t sf, float *out, void *stream) {
  return ibp_impl(x0, y, down, gamma, rho, nullptr, N, H, W, sf, out, stream);
}
int dpir_ibp_update_dev(const float *x0, const float *y, const float *down, float gamma, const float *rho_dev, int32_t N, int32_t H,
                        int32_t W, int32_t sf, float *out, void *stream) {
  DPIR_REQUIRE(rho_dev, "dpir_ibp_update_dev: null rho");
  return ibp_impl(x0, y, down, gamma, 0.f, rho_dev, N, H, W, sf, out, stream);
}

int dpir_blur_reflect(const float *x, const float *k, int32_t planes, int32_t H, int32_t W, int32_t kh, int32_t kw, float *out,
                      void *stream) {
  DPIR_REQUIRE(x && k && out && kh / 2 < H && kw / 2 < W, "dpir_blur_reflect: bad arguments");
  size_t smem = (size_t)kh * kw * sizeof(float);
  DPIR_REQUIRE(smem <= 48 * 1024, "dpir_blur_reflect: kernel too large");
  k_blur_reflect<<<grid_for((int64_t)planes * H * W), 256, smem, (cudaStream_t)stream>>>(x, k, planes, H, W, kh, kw, out);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

int dpir_tensor2uint(const float *x, uint8_t *out_nhwc, int32_t B, int32_t C, int32_t H, int32_t W, void *stream) {
  DPIR_REQUIRE(x && out_nhwc, "dpir_tensor2uint: null pointer");
  k_tensor2uint<<<grid_for((int64_t)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(x, out_nhwc, B, C, H, W);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

int dpir_sqdiff_batch(const float *a, const float *b, float a_scale, float a_bias, float b_scale, float b_bias, int32_t luma,
                      int32_t B, int32_t C, int32_t HW, double *sse, void *stream) {
  DPIR_REQUIRE(a && b && sse, "dpir_sqdiff_batch: null pointer");
  DPIR_REQUIRE(!luma || C == 3, "dpir_sqdiff_batch: luma needs 3 channels");
  cudaStream_t st = (cudaStream_t)stream;
  DPIR_CUDA_CHECK(cudaMemsetAsync(sse, 0, (size_t)B * sizeof(double), st));
  int64_t n = luma ? HW : (int64_t)C * HW;
  int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  k_sqdiff<<<dim3(gx, B), 256, 0, st>>>(a, b, a_scale, a_bias, b_scale, b_bias, luma, C, HW, sse);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

int dpir_measure_blur_u8(const uint8_t *img_nhwc, const double *k, const float *noise_nhwc, uint8_t *blur_u8, float *out_nchw,
                         int32_t B, int32_t H, int32_t W, int32_t C, int32_t kh, int32_t kw, void *stream) {
  DPIR_REQUIRE(img_nhwc && k && (blur_u8 || out_nchw), "dpir_measure_blur_u8: null pointer");
  size_t smem = (size_t)kh * kw * sizeof(double);
  DPIR_REQUIRE(smem <= 48 * 1024, "dpir_measure_blur_u8: kernel too large");
  int64_t n = (int64_t)H * W * C;
  int gx = (int)((n + 255) / 256);
  if (gx > 148 * 8) gx = 148 * 8;
  k_measure_blur_u8<<<dim3(gx, B), 256, smem, (cudaStream_t)stream>>>(img_nhwc, k, noise_nhwc, blur_u8, out_nchw, B, H, W, C, kh, kw);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

}  // extern "C"
