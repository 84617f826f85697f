// conv_tc.cu - 3x3 / 1x1 stride-1 convolution as an implicit GEMM on the Blackwell 5th-gen tensor cores.
//
// Stands in for every nn.Conv2d(k=3,pad=1) / Conv2d(k=1) / Conv1d(k=1) of the UNet
// (reference guided_diffusion/unet.py:185, 211, 222, 286, 294, 482, 615) in the fp16-operand / fp32-accumulate mode.
//
//   D[m][n] = sum_{tap} sum_{c} A_tap[m][c] * Wt[tap][n][c]        m = output pixel, n = output channel
//
// * A operand: activations NHWC fp16.  One M tile = 128 output pixels = a (TN x TH x TW) box of the image
//   (TW*TH*TN = 128).  For tap (dy,dx) the A tile is the SAME box shifted by (dy-1, dx-1); it is fetched by ONE 4-D TMA
//   (cp.async.bulk.tensor.4d) whose out-of-bounds rows/cols/images are zero-filled by the hardware - that is the conv's
//   zero padding, so no im2col buffer, no halo logic and no predication exists anywhere.
// * B operand: weights fp16 packed [tap][CoutPad][Cin] (K-major), fetched by a 3-D TMA.
// * Both land in shared memory in the canonical K-major SWIZZLE_128B layout (64 fp16 = 128 B per row), consumed directly
//   by tcgen05.mma (kind::f16, M=128, N=BN, K=16) through shared-memory matrix descriptors; fp32 accumulators live in
//   TMEM (128 lanes x BN columns).
// * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//   warps 2-5 = epilogue (tcgen05.ld 32x32b -> registers -> bias / residual / NCHW / eps->x0 -> global).
// * STAGES-deep smem ring with full/empty mbarriers; MMA completion releases a stage via tcgen05.commit.
//   Two CTAs fit per SM (3 x 32 KB stages, 128 TMEM columns each) so one CTA's epilogue overlaps the other's main loop.
//
// Kernels in this file (all share the TMA / SWIZZLE_128B / tcgen05 scheme above; DESIGN.md section 4 has the measurements):
//   k_conv_tc<BN,STAGES>    one 128-pixel x BN-cout tile per CTA: 1x1 convs, 8x8 maps, the 128->6 output conv (eps->x0 epilogue)
//   k_conv3x3_v2            persistent halo-reuse 3x3 kernel, pixels on M (A/B variant, dpir_conv_tc_variant(2))
//   k_conv3x3_v3<0>         persistent swap-AB 3x3 kernel (couts on M, 256 pixels on N), TMA-store epilogue: the default for
//                           3x3 convs on >=16x16 maps in the fp16-stream pipeline
//   k_conv3x3_v3<1|2>       the same with the preceding GroupNorm(+FiLM)+SiLU applied to the activation operand by 8 producer
//                           warps ("v4", default for eligible ResBlocks); <3> is an unvalidated activate-once variant
#include "unet_kernels.cuh"
#include "tc_ptx.cuh"

#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>

namespace dpir {

// ------------------------------------------------------------------------------------------------ kernel
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                       // fp16 elements per K chunk = 128 bytes = one swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB
constexpr int TC_THREADS = 192;

struct ConvTcGeom {
  int B, H, W, Cin, Cout, taps, ksize;
  int a_wrap;                 // split-precision: K chunk ch reads activation chunk (ch >= a_wrap ? ch - a_wrap : ch); else INT_MAX
  int TW, TH, TN;             // M-tile box (TW*TH*TN == 128)
  int tiles_w, tiles_h, tiles_n;
};

// Per-quad (4-channel) sums over the 32 pixel rows held by a warp: reduce-scatter butterfly (8 -> 4 -> 2 -> 1 values per
// lane over xor 16/8/4, then plain xor 2/1: 18 shuffles per 32-column chunk), then one fp64 atomic pair per quad.
__device__ __forceinline__ void warp_quad_stats(const float *qs, const float *qq, int lane, bool ok, double *qstats, int n,
                                                int Cout, int co) {
  float a4[4], b4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float s_send = (lane & 16) ? qs[i] : qs[i + 4], s_keep = (lane & 16) ? qs[i + 4] : qs[i];
    float q_send = (lane & 16) ? qq[i] : qq[i + 4], q_keep = (lane & 16) ? qq[i + 4] : qq[i];
    a4[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, 16);
    b4[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, 16);
  }
  float a2[2], b2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float s_send = (lane & 8) ? a4[i] : a4[i + 2], s_keep = (lane & 8) ? a4[i + 2] : a4[i];
    float q_send = (lane & 8) ? b4[i] : b4[i + 2], q_keep = (lane & 8) ? b4[i + 2] : b4[i];
    a2[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, 8);
    b2[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, 8);
  }
  float s1 = ((lane & 4) ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, (lane & 4) ? a2[0] : a2[1], 4);
  float q1 = ((lane & 4) ? b2[1] : b2[0]) + __shfl_xor_sync(0xffffffffu, (lane & 4) ? b2[0] : b2[1], 4);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
  const int quad = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  if ((lane & 3) == 0 && ok && co + quad * 4 < Cout) {
    double *qp = qstats + ((int64_t)n * (Cout >> 2) + (co >> 2) + quad) * 2;
    atomicAdd(qp, (double)s1);
    atomicAdd(qp + 1, (double)q1);
  }
}

// NHWC epilogue of one 32-column chunk held by one warp (lane = pixel row): + bias (+ residual), store fp32 or fp16,
// and (optionally) accumulate the per-quad GroupNorm statistics of the stored tensor.
__device__ __forceinline__ void epi_nhwc_chunk32(const ConvEpi &epi, const ConvTcGeom &g, const uint32_t *r, bool valid, int n,
                                                 int yh, int xw, int co, int lane, bool warp_img_ok) {
  float qs[8], qq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { qs[j] = 0.f; qq[j] = 0.f; }
  if (valid) {
    const float sc = epi.acc_scale ? __ldg(epi.acc_scale) : 1.f;
    const int64_t pix = ((int64_t)n * g.H + yh) * g.W + xw;
    const float *rs = (epi.residual && !epi.res_half) ? epi.residual + pix * g.Cout + co : nullptr;
    const __half *rsh = (epi.residual && epi.res_half) ? reinterpret_cast<const __half *>(epi.residual) + pix * g.Cout + co : nullptr;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (co + j < g.Cout) {
        float4 b4 = __ldg(reinterpret_cast<const float4 *>(epi.bias + co + j));
        if (epi.bias2) {
          float4 c4 = __ldg(reinterpret_cast<const float4 *>(epi.bias2 + co + j));
          b4.x += c4.x; b4.y += c4.y; b4.z += c4.z; b4.w += c4.w;
        }
        v[j] = fmaf(__uint_as_float(r[j]), sc, b4.x); v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc, b4.y);
        v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc, b4.z); v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc, b4.w);
        if (rs) {
          float4 rv = *reinterpret_cast<const float4 *>(rs + j);
          v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
        } else if (rsh) {
          uint2 u = *reinterpret_cast<const uint2 *>(rsh + j);
          float2 f0 = __half22float2(*reinterpret_cast<__half2 *>(&u.x)), f1 = __half22float2(*reinterpret_cast<__half2 *>(&u.y));
          v[j] += f0.x; v[j + 1] += f0.y; v[j + 2] += f1.x; v[j + 3] += f1.y;
        }
        qs[j >> 2] = (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
        qq[j >> 2] = (v[j] * v[j] + v[j + 1] * v[j + 1]) + (v[j + 2] * v[j + 2] + v[j + 3] * v[j + 3]);
      } else {
        v[j] = v[j + 1] = v[j + 2] = v[j + 3] = 0.f;
      }
    }
    if (epi.out_half) {
      __half *o = reinterpret_cast<__half *>(epi.out) + pix * g.Cout + co;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (co + j < g.Cout) {
          __half2 h0 = __floats2half2_rn(v[j], v[j + 1]), h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
          __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]), h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t *>(&h0); u.y = *reinterpret_cast<uint32_t *>(&h1);
          u.z = *reinterpret_cast<uint32_t *>(&h2); u.w = *reinterpret_cast<uint32_t *>(&h3);
          *reinterpret_cast<uint4 *>(o + j) = u;
        }
      }
    } else {
      float *o = epi.out + pix * g.Cout + co;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        if (co + j < g.Cout) *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  if (epi.qstats) {
    if (g.TW * g.TH >= 32) {
      warp_quad_stats(qs, qq, lane, warp_img_ok, epi.qstats, n, g.Cout, co);      // the 32 rows of a warp lie in one image
    } else if (valid) {
      // maps smaller than 32 pixels (4x4, 2x2: the deepest levels of the 512x512 model family, or small inputs): a warp's rows
      // span several images, so every pixel row adds its own quads (a handful of atomics on tensors of a few KB)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (co + 4 * j < g.Cout) {
          double *qp = epi.qstats + ((int64_t)n * (g.Cout >> 2) + ((co + 4 * j) >> 2)) * 2;
          atomicAdd(qp, (double)qs[j]);
          atomicAdd(qp + 1, (double)qq[j]);
        }
      }
    }
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(TC_THREADS) k_conv_tc(const __grid_constant__ CUtensorMap tmA,
                                                         const __grid_constant__ CUtensorMap tmB, ConvTcGeom g, ConvEpi epi) {
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  // instruction descriptor: D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  uint64_t *tmem_full = empty + STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;
  const int w0 = (mt % g.tiles_w) * g.TW;
  const int h0 = ((mt / g.tiles_w) % g.tiles_h) * g.TH;
  const int n0 = (mt / (g.tiles_w * g.tiles_h)) * g.TN;
  const int co0 = blockIdx.y * BN;
  const int chunks = g.Cin / TC_BK;
  const int num_k = g.taps * chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int pad = g.ksize / 2;
      for (int k = 0; k < num_k; ++k) {
        const int s = k % STAGES;
        const uint32_t ph = (uint32_t)(k / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const int tap = k / chunks, ch = k - tap * chunks;
        const int dy = tap / g.ksize - pad, dx = tap % g.ksize - pad;
        uint8_t *sa = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        const int ach = ch >= g.a_wrap ? ch - g.a_wrap : ch;
        tma_load_4d(sa, &tmA, &full[s], ach * TC_BK, w0 + dx, h0 + dy, n0);
        tma_load_3d(sa + TC_A_BYTES, &tmB, &full[s], ch * TC_BK, co0, tap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int k = 0; k < num_k; ++k) {
        const int s = k % STAGES;
        const uint32_t ph = (uint32_t)(k / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t adesc = make_sw128_desc(sa);
        const uint64_t bdesc = make_sw128_desc(sa + TC_A_BYTES);
#pragma unroll
        for (int kk = 0; kk < TC_BK / 16; ++kk) {
          // advance 16 fp16 = 32 bytes along K inside the 128 B swizzle row: +2 in the (>>4) address field
          umma_f16(tmem_base, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), IDESC, (k | kk) ? 1u : 0u);
        }
        umma_commit(&empty[s]);   // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full);     // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 4 warps x 32 TMEM lanes
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // M index inside the tile
    const int xw = w0 + row % g.TW;
    const int yh = h0 + (row / g.TW) % g.TH;
    const int n = n0 + row / (g.TW * g.TH);
    const bool valid = (n < g.B) && (yh < g.H) && (xw < g.W);
    const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
    if (BN >= 32) {
      const bool warp_img_ok = n < g.B;       // the 32 rows of a warp lie in one image (TW*TH >= 32)
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tbase + (uint32_t)c, r);
        tmem_ld_wait();
        const int co = co0 + c;
        if (epi.mode == EPI_NHWC) {
          epi_nhwc_chunk32(epi, g, r, valid, n, yh, xw, co, lane, warp_img_ok);
        } else if (valid) {
          float acc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
          int ncols = g.Cout - co;
          if (ncols > 32) ncols = 32;
          if (ncols > 0) conv_epilogue_store(epi, n, yh, xw, g.H, g.W, g.Cout, co, acc, ncols);
        }
      }
    } else {
      uint32_t r[16];
      tmem_ld16(tbase, r);
      tmem_ld_wait();
      if (valid) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]);
        int ncols = g.Cout - co0;
        if (ncols > 16) ncols = 16;
        if (ncols > 0) conv_epilogue_store(epi, n, yh, xw, g.H, g.W, g.Cout, co0, acc, ncols);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ v2: persistent 3x3 kernel
// The v1 kernel above re-fetches the A tile once per tap (9x) and the weight tile once per CTA: ncu shows it bound by
// L2->SM traffic (8.5 TB/s through the crossbar, tensor pipe 19-27 % active), not by the tensor cores.  v2 cuts that
// traffic ~2.3x and hides the epilogue:
//   * M tile = 16 x 16 pixels = two 128-row sub-tiles (image rows 0-7 / 8-15) with separate TMEM accumulators that SHARE
//     every weight tile (B traffic per pixel halved);
//   * per (64-channel chunk, dx) ONE 4-D TMA fetches an 18-row x 16-col halo box (shifted by dx-1 in W, zero-filled
//     outside the image); the three vertical taps dy are three views into it, each starting dy*16 pixels = dy*2048 B
//     further - still 1024 B aligned, so the same SWIZZLE_128B descriptors apply (A traffic 3 x 18/16 = 3.4x instead of 9x);
//   * persistent CTAs (one per SM) walk a static tile list; TMEM holds 2 x (2 x 128) accumulator columns so the epilogue
//     of tile i (8 warps) overlaps the TMA/MMA main loop of tile i+1.
constexpr int V2_A_BYTES = 18 * 16 * 128;            // halo box: 288 pixel rows x 64 fp16
constexpr int V2_B_BYTES = 3 * 128 * 128;            // 3 vertical taps x (128 couts x 64 fp16)
constexpr int V2_STAGE_BYTES = V2_A_BYTES + V2_B_BYTES;   // 86,016 B
constexpr int V2_STAGES = 2;
constexpr int V2_THREADS = 320;                       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
// v4 (fused GroupNorm+SiLU operand): 5 warpgroups - {TMA, MMA, 2 idle warps}, 2 x epilogue (warps 4-11), 2 x operand
// producers (warps 12-19).  5 warps per SM sub-partition leave 96 registers per thread, fewer than the epilogue (117) and the
// double-buffered producers need, so the warpgroups re-balance the CTA's register allocation (640 x 96) with setmaxnreg:
// 32 / 120 / 104 per thread (128*32 + 256*120 + 256*104 = 640*96: the pool is what the launch allocated, no more).
constexpr int V4_THREADS = 640;
constexpr int V2_STG_BYTES = 8 * 4096;                // per-epilogue-warp 32 x 128 B transpose buffers
constexpr int V2_SMEM = V2_STAGES * V2_STAGE_BYTES + V2_STG_BYTES + 1024 + 256;
constexpr int V3_STG_BYTES = 8 * 6144;                // v3: per-epilogue-warp fp16 tiles OUT, RES0, RES1 (2 KB each)
// v3 pipeline: the activation operand and the weights live in SEPARATE rings.  An activation slot is one (64-channel chunk, dx)
// halo box (36 KB, read by the three dy taps); a weight slot is one tap tile (128 couts x 64 k, 16 KB).  Three activation slots
// hold all three dx copies of a chunk, so the activate-once producers (FUSE == 3) publish a whole chunk back to back and have
// the tensor core's entire chunk time (3 x 12 MMAs) to fetch and activate the next one; with the fused 2-stage ring of round 1
// the load latency of every chunk was exposed (profiles/r02_conv_ring_experiments.md).
constexpr int V3_A_SLOTS = 3;
constexpr int V3_W_SLOTS = 4;
constexpr int V3_W_BYTES = 128 * 128;                 // one tap: 128 couts x 64 fp16
constexpr int V3_RING_BYTES = V3_A_SLOTS * V2_A_BYTES + V3_W_SLOTS * V3_W_BYTES;     // 176,128 B
constexpr int V3_SMEM = V3_RING_BYTES + V3_STG_BYTES + 1024 + 1024;    // + alignment slack + barrier block

struct ConvV2Geom {
  int B, H, W, Cin, Cout;
  int a_wrap;                         // split-precision K-chunk -> activation-chunk wrap (see ConvTcGeom), INT_MAX otherwise
  int skip_chunks;                    // extra K chunks of a fused 1x1 convolution on a second operand (ResBlock skip), or 0
  int skip_chunks_a;                  // of which the first skip_chunks_a come from tmA2, the rest from tmA2b (concat skip input)
  int dbg;                            // timing diagnostics only (env DPIR_CONV_DEBUG): 1 = no epilogue memory traffic,
                                      // 2 = no MMA issue, 4 = no TMA loads; v4 producers: 32 = no global loads, 64 = no
                                      // arithmetic, 128 = no proxy fence; results are garbage when non-zero
  int tiles_w, tiles_h, n_tiles;      // 16x16 pixel tiles per image, Cout/128 column tiles
  int total_tiles;                    // B * tiles_h * tiles_w * n_tiles
  int a_slots, w_slots, w_taps;       // v3 rings: activation halo boxes (36 KB each) and weight slots of w_taps taps (16 KB per tap);
                                      // a_slots*36 KB + w_slots*w_taps*16 KB <= V3_RING_BYTES
  int pf_dist;                        // fused kernels: L2 prefetch distance of the raw operand in 64-channel chunks (0 = off)
  int stg_stride;                     // epilogue staging bytes per warp: 6144 (OUT + two residual boxes) or 2048 (OUT only, convs
                                      // without a residual: the 32 KB saved buy a deeper weight ring)
};

__device__ __forceinline__ float tanh_approx_f(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ __half2 tanh_approx_h2(__half2 x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(*reinterpret_cast<const uint32_t *>(&x)));
  return *reinterpret_cast<const __half2 *>(&y);
}

template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Optional fusion of the ResBlock's 1x1 skip_connection (unet.py:222, 256): out = conv3x3(a2) + conv1x1(x) is one
// accumulation - after the 3*chunks (chunk, dx) stages, `skip_chunks` more stages load the centre-aligned halo box of the
// second operand (tmA2) and the 1x1 weights (tmB2) and issue only the dy = 1 view.
__global__ void __launch_bounds__(V2_THREADS, 1) k_conv3x3_v2(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmA2,
                                                               const __grid_constant__ CUtensorMap tmB2, ConvV2Geom g,
                                                               ConvEpi epi) {
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *stg_base = smem + V2_STAGES * V2_STAGE_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(stg_base + V2_STG_BYTES);
  uint64_t *empty = full + V2_STAGES;
  uint64_t *tmem_full = empty + V2_STAGES;     // [2]
  uint64_t *tmem_empty = tmem_full + 2;        // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = g.Cin / TC_BK;
  const int num_main = 3 * chunks;
  const int num_k = num_main + g.skip_chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.skip_chunks) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    for (int s = 0; s < V2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile id -> (image, tile row, tile col, cout tile); the cout tile varies fastest so that concurrently running CTAs
  // share the same activation box in L2
  auto decode = [&](int tile, int &n, int &h0, int &w0, int &co0) {
    int nt = tile % g.n_tiles;
    int mt = tile / g.n_tiles;
    co0 = nt * 128;
    w0 = (mt % g.tiles_w) * 16;
    h0 = ((mt / g.tiles_w) % g.tiles_h) * 16;
    n = mt / (g.tiles_w * g.tiles_h);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        int n, h0, w0, co0;
        decode(tile, n, h0, w0, co0);
        for (int k = 0; k < num_k; ++k, ++it) {
          const int s = it % V2_STAGES;
          const uint32_t ph = (it / V2_STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t *sa = smem + s * V2_STAGE_BYTES;
          if (k < num_main) {
            const int ch = k / 3, dx = k - ch * 3;
            const int ach = ch >= g.a_wrap ? ch - g.a_wrap : ch;
            mbar_expect_tx(&full[s], V2_STAGE_BYTES);
            tma_load_4d(sa, &tmA, &full[s], ach * TC_BK, w0 + dx - 1, h0 - 1, n);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
              tma_load_3d(sa + V2_A_BYTES + dy * (128 * 128), &tmB, &full[s], ch * TC_BK, co0, dy * 3 + dx);
          } else {
            const int ch = k - num_main;
            mbar_expect_tx(&full[s], V2_A_BYTES + 128 * 128);
            tma_load_4d(sa, &tmA2, &full[s], ch * TC_BK, w0, h0 - 1, n);
            tma_load_3d(sa + V2_A_BYTES + 1 * (128 * 128), &tmB2, &full[s], ch * TC_BK, co0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tc) {
        const uint32_t ab = tc & 1u, aph = (tc >> 1) & 1u;
        mbar_wait(&tmem_empty[ab], aph ^ 1u);      // epilogue has drained this accumulator pair
        tc_fence_after();
        for (int k = 0; k < num_k; ++k, ++it) {
          const int s = it % V2_STAGES;
          const uint32_t ph = (it / V2_STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * V2_STAGE_BYTES);
          const int dy_lo = k < num_main ? 0 : 1, dy_hi = k < num_main ? 3 : 2;
          for (int dy = dy_lo; dy < dy_hi; ++dy) {
            const uint64_t bdesc = make_sw128_desc(sa + V2_A_BYTES + dy * (128 * 128));
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              // rows (sub*8 + dy) .. +7 of the 18-row halo box, 16 pixels each: 128 consecutive 128 B rows
              const uint64_t adesc = make_sw128_desc(sa + (uint32_t)((sub * 8 + dy) * 16 * 128));
              const uint32_t d = tmem_base + ab * 256u + (uint32_t)sub * 128u;
#pragma unroll
              for (int kk = 0; kk < TC_BK / 16; ++kk)
                umma_f16(d, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), IDESC, (k | dy | kk) ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[ab]);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 8 warps, coalesced through shared memory
    // A lane owns one pixel row of the accumulator (TMEM lane), but a pixel's channels are what is contiguous in NHWC:
    // writing row-per-lane touches 32 different 128 B lines per instruction.  Each warp therefore transposes 32 rows x 128 B
    // through a private XOR-swizzled buffer (16 B column c of row r at c ^ (r & 7): conflict-free both ways) so that global
    // loads/stores are issued with 8 consecutive lanes covering one full 128 B line (4 lines per instruction).
    const int ew = warp - 2;              // 0..7
    const int q = warp & 3;               // TMEM lane quarter accessible to this warp
    const int sub = ew >> 2;              // which 128-row sub-tile
    float4 *stg = reinterpret_cast<float4 *>(stg_base + ew * 4096);
    const int tr = lane >> 3, tcn = lane & 7;     // transposed role: row (4i + tr), 16-byte column tcn
    const float acc_sc = epi.acc_scale ? __ldg(epi.acc_scale) : 1.f;      // split-precision weight scale (fp32-output path only)
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tc) {
      const uint32_t ab = tc & 1u, aph = (tc >> 1) & 1u;
      int n, h0, w0, co0;
      decode(tile, n, h0, w0, co0);
      // pixel of sub-tile row rr (0..31 inside this warp's quarter)
      auto pix_of = [&](int rr, bool &ok) -> int64_t {
        const int row = q * 32 + rr;
        const int yh = h0 + sub * 8 + (row >> 4), xw = w0 + (row & 15);
        ok = (yh < g.H) && (xw < g.W);
        return ((int64_t)n * g.H + yh) * g.W + xw;
      };
      bool my_ok;
      (void)pix_of(lane, my_ok);
      mbar_wait(&tmem_full[ab], aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256u + (uint32_t)sub * 128u;
      if (!epi.out_half) {
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          const int co = co0 + c;
          float4 res[8];
          if (epi.residual) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + tr;
              bool ok;
              const int64_t pix = pix_of(rr, ok);
              float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok) v4 = *reinterpret_cast<const float4 *>(epi.residual + pix * g.Cout + co + tcn * 4);
              stg[rr * 8 + (tcn ^ (rr & 7))] = v4;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) res[j] = stg[lane * 8 + (j ^ (lane & 7))];
            __syncwarp();
          }
          uint32_t r[32];
          tmem_ld32(tbase + (uint32_t)c, r);
          tmem_ld_wait();
          float qs[8], qq[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b4 = __ldg(reinterpret_cast<const float4 *>(epi.bias + co + 4 * j));
            if (epi.bias2) {
              float4 c4 = __ldg(reinterpret_cast<const float4 *>(epi.bias2 + co + 4 * j));
              b4.x += c4.x; b4.y += c4.y; b4.z += c4.z; b4.w += c4.w;
            }
            float4 v4 = make_float4(fmaf(__uint_as_float(r[4 * j]), acc_sc, b4.x), fmaf(__uint_as_float(r[4 * j + 1]), acc_sc, b4.y),
                                    fmaf(__uint_as_float(r[4 * j + 2]), acc_sc, b4.z), fmaf(__uint_as_float(r[4 * j + 3]), acc_sc, b4.w));
            if (epi.residual) { v4.x += res[j].x; v4.y += res[j].y; v4.z += res[j].z; v4.w += res[j].w; }
            qs[j] = my_ok ? (v4.x + v4.y) + (v4.z + v4.w) : 0.f;
            qq[j] = my_ok ? (v4.x * v4.x + v4.y * v4.y) + (v4.z * v4.z + v4.w * v4.w) : 0.f;
            stg[lane * 8 + (j ^ (lane & 7))] = v4;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + tr;
            bool ok;
            const int64_t pix = pix_of(rr, ok);
            if (ok) *reinterpret_cast<float4 *>(epi.out + pix * g.Cout + co + tcn * 4) = stg[rr * 8 + (tcn ^ (rr & 7))];
          }
          __syncwarp();
          if (epi.qstats) warp_quad_stats(qs, qq, lane, true, epi.qstats, n, g.Cout, co);
        }
      } else {
        // fp16 output (conv1 -> GroupNorm -> conv2): 64 columns = 128 B per row per round
        __half *outh = reinterpret_cast<__half *>(epi.out);
        const __half *resh = reinterpret_cast<const __half *>(epi.residual);   // fp16 residual stream (res_half) or null
#pragma unroll 1
        for (int c = 0; c < 128; c += 64) {
          const int co = co0 + c;
          float4 resp[8];      // this lane's residual row: 8 x (8 halves)
          if (resh) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + tr;
              bool ok;
              const int64_t pix = pix_of(rr, ok);
              float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok) v4 = *reinterpret_cast<const float4 *>(resh + pix * g.Cout + co + tcn * 8);
              stg[rr * 8 + (tcn ^ (rr & 7))] = v4;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) resp[j] = stg[lane * 8 + (j ^ (lane & 7))];
            __syncwarp();
          }
#pragma unroll
          for (int hlf = 0; hlf < 2; ++hlf) {
            uint32_t r[32];
            tmem_ld32(tbase + (uint32_t)(c + 32 * hlf), r);
            tmem_ld_wait();
            float qs[8], qq[8];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float v[8];
              float rv[8];
              if (resh) {
                const float4 pk = resp[4 * hlf + (j >> 1)];
                uint32_t w0_ = __float_as_uint(pk.x), w1_ = __float_as_uint(pk.y), w2_ = __float_as_uint(pk.z), w3_ = __float_as_uint(pk.w);
                float2 f0 = __half22float2(*reinterpret_cast<__half2 *>(&w0_)), f1 = __half22float2(*reinterpret_cast<__half2 *>(&w1_));
                float2 f2 = __half22float2(*reinterpret_cast<__half2 *>(&w2_)), f3 = __half22float2(*reinterpret_cast<__half2 *>(&w3_));
                rv[0] = f0.x; rv[1] = f0.y; rv[2] = f1.x; rv[3] = f1.y; rv[4] = f2.x; rv[5] = f2.y; rv[6] = f3.x; rv[7] = f3.y;
              } else {
#pragma unroll
                for (int z = 0; z < 8; ++z) rv[z] = 0.f;
              }
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                float4 b4 = __ldg(reinterpret_cast<const float4 *>(epi.bias + co + 32 * hlf + 4 * (j + u)));
                if (epi.bias2) {
                  float4 c4 = __ldg(reinterpret_cast<const float4 *>(epi.bias2 + co + 32 * hlf + 4 * (j + u)));
                  b4.x += c4.x; b4.y += c4.y; b4.z += c4.z; b4.w += c4.w;
                }
                v[4 * u + 0] = __uint_as_float(r[4 * (j + u) + 0]) + b4.x + rv[4 * u + 0];
                v[4 * u + 1] = __uint_as_float(r[4 * (j + u) + 1]) + b4.y + rv[4 * u + 1];
                v[4 * u + 2] = __uint_as_float(r[4 * (j + u) + 2]) + b4.z + rv[4 * u + 2];
                v[4 * u + 3] = __uint_as_float(r[4 * (j + u) + 3]) + b4.w + rv[4 * u + 3];
                qs[j + u] = my_ok ? (v[4 * u] + v[4 * u + 1]) + (v[4 * u + 2] + v[4 * u + 3]) : 0.f;
                qq[j + u] = my_ok ? (v[4 * u] * v[4 * u] + v[4 * u + 1] * v[4 * u + 1]) +
                                        (v[4 * u + 2] * v[4 * u + 2] + v[4 * u + 3] * v[4 * u + 3]) : 0.f;
              }
              __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
              __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
              float4 pk;
              pk.x = __uint_as_float(*reinterpret_cast<uint32_t *>(&h0)); pk.y = __uint_as_float(*reinterpret_cast<uint32_t *>(&h1));
              pk.z = __uint_as_float(*reinterpret_cast<uint32_t *>(&h2)); pk.w = __uint_as_float(*reinterpret_cast<uint32_t *>(&h3));
              const int col16 = 4 * hlf + (j >> 1);      // 16-byte column (8 halves) inside the 128 B row
              stg[lane * 8 + (col16 ^ (lane & 7))] = pk;
            }
            if (epi.qstats) warp_quad_stats(qs, qq, lane, true, epi.qstats, n, g.Cout, co + 32 * hlf);
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + tr;
            bool ok;
            const int64_t pix = pix_of(rr, ok);
            if (ok) *reinterpret_cast<float4 *>(outh + pix * g.Cout + co + tcn * 8) = stg[rr * 8 + (tcn ^ (rr & 7))];
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ v3: swap-AB persistent kernel
// Same tiling / TMA / pipeline as v2, but the GEMM is issued transposed: D^T[cout][pixel] = W[cout][k] * A[pixel][k]^T with
// M = 128 output channels (weights are the A operand) and N = 256 pixels (the whole 16x16 tile is ONE B operand: rows
// dy*16 .. dy*16+255 of the halo box).  Per 64-channel chunk and tap this is 4 MMAs of 128x256x16 reading 4 KB (W) + 8 KB
// (pixels) of shared memory per 128 tensor cycles instead of 2 x (4 + 4) KB - 25 % less operand traffic on the
// shared-memory port, which is what bounds the SS-mode 128x128 MMAs of v2 (ncu: ~65 % tensor pipe on 128->128 layers).
// The accumulator then has CHANNELS on TMEM lanes and pixels on columns, so in the epilogue a warp instruction writes 32
// consecutive channels of one pixel = one contiguous 128 B (fp32) / 64 B (fp16) segment: coalesced without any transpose,
// bias is a per-lane constant, and the GroupNorm quad statistics reduce to a thread-local sum over pixels plus two shuffles.
//
// FUSE > 0 (v4): the GroupNorm (+FiLM) + SiLU that precedes the convolution in every ResBlock (unet.py:199-200, 205-212,
// nn.py:17-19) is applied on the fly to the activation operand, so the HBM-bound k_gn_apply pass (read + write of the whole
// tensor) and its output buffer disappear.  The normalisation is a per-(image, channel) affine map (k_gn_coef folds
// mean / rstd / gamma / beta / FiLM scale+shift into hs = 0.5*scale, ho = 0.5*offset), so 8 extra producer warps replace the
// activation TMA: they read the RAW fp16 tensor(s) - one or two sources = the virtual th.cat of unet.py:660 - with 16-byte
// global loads (prefetched one pipeline stage ahead into registers), compute h = x*hs + ho, SiLU(2h) = h + h*tanh(h)
// (one MUFU per element, fp32; FUSE == 2: packed-half arithmetic), write the result into the same K-major SWIZZLE_128B
// layout the TMA would have produced (16-byte chunk j of pixel row r lands at r*128 + ((j ^ (r & 7)) << 4)), zero outside
// the image (the conv pads the ACTIVATED tensor), and publish the stage through fence.proxy.async + the stage's mbarrier.
// The weight tiles, the fused 1x1-skip stages (raw operand, TMA; two tensor maps when the skip input is a concat), the
// MMA issue and the epilogue are unchanged.
template <int FUSE>
__global__ void __launch_bounds__(FUSE ? V4_THREADS : V2_THREADS, 1) k_conv3x3_v3(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmAb,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmA2,
                                                               const __grid_constant__ CUtensorMap tmA2b,
                                                               const __grid_constant__ CUtensorMap tmB2,
                                                               const __grid_constant__ CUtensorMap tmOut,
                                                               const __grid_constant__ CUtensorMap tmRes, ConvV2Geom g,
                                                               ConvEpi epi, GnFuse gf) {
  // M = 128 (couts), N = 256 (pixels)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // barriers first (fixed offset), then the rings, then the per-epilogue-warp staging (OUT, RES0, RES1: 2 KB each; OUT only when the
  // convolution has no residual) - ring and staging sizes are launch parameters, their sum never exceeds V3_RING_BYTES + V3_STG_BYTES
  uint64_t *a_full = reinterpret_cast<uint64_t *>(smem);
  const int A_SL = g.a_slots, W_SL = g.w_slots;              // ring depths (runtime: tuned per kernel flavour)
  uint8_t *a_ring = smem + 1024;                             // [A_SL][V2_A_BYTES]  activation halo boxes
  uint8_t *w_ring = a_ring + A_SL * V2_A_BYTES;              // [W_SL][w_taps * V3_W_BYTES]  weight tap tiles
  uint8_t *stg_base = w_ring + W_SL * g.w_taps * V3_W_BYTES;
  uint64_t *a_empty = a_full + 4;              // up to 4 activation slots
  uint64_t *w_full = a_empty + 4;
  uint64_t *w_empty = w_full + 8;              // up to 8 weight slots
  uint64_t *tmem_full = w_empty + 8;           // [2]
  uint64_t *tmem_empty = tmem_full + 2;        // [2]
  uint64_t *res_full = tmem_empty + 2;         // [8 warps][2 buffers]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(res_full + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = g.Cin / TC_BK;
  const int num_main = 3 * chunks;
  const int num_k = num_main + g.skip_chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.skip_chunks) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB2); }
    // an activation slot is full once the TMA thread has arrived (expect_tx + bytes when the operand comes by TMA: plain kernel
    // and fused 1x1-skip stages) AND, in the fused kernels, the 8 (FUSE 1|2) / 9 (FUSE 3) producer warps have published it
    for (int s = 0; s < A_SL; ++s) { mbar_init(&a_full[s], FUSE == 3 ? 10 : (FUSE ? 9 : 1)); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < W_SL; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    for (int a = 0; a < 16; ++a) mbar_init(&res_full[a], 1);
    tma_prefetch_desc(&tmOut);
    if (epi.residual) tma_prefetch_desc(&tmRes);
    fence_barrier_init();
    fence_proxy_async();
  }
  constexpr int MMA_WARP = FUSE == 3 ? 18 : 1;      // also owns the TMEM allocation
  if (warp == MMA_WARP) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int &n, int &h0, int &w0, int &co0) {
    int nt = tile % g.n_tiles;
    int mt = tile / g.n_tiles;
    co0 = nt * 128;
    w0 = (mt % g.tiles_w) * 16;
    h0 = ((mt / g.tiles_w) % g.tiles_h) * 16;
    n = mt / (g.tiles_w * g.tiles_h);
  };

  // warp roles as lambdas: the two thread layouts (v3: 10 warps; v4: 5 warpgroups with re-balanced registers) share them
  // Ring bookkeeping shared by all roles: every (chunk, dx) stage and every fused-skip chunk consumes ONE activation slot
  // (counter ia), every tap ONE weight slot (counter iw); all roles walk the same static tile list in the same order.
  // Ring bookkeeping shared by all roles: every (chunk, dx) stage and every fused-skip chunk consumes ONE activation slot
  // (counter ia); the weights of a stage take 3 / W_TP weight slots (counter iw; a fused-skip chunk takes one); all roles walk
  // the same static tile list in the same order.
  const int W_TP = g.w_taps;
  const int W_SLOT_BYTES = W_TP * V3_W_BYTES;
  // stage order of a tile, walked identically by every role: for each main chunk its three (chunk, dx) stages, then the fused-skip
  // chunks [sk, skip_hi(chunk)): all of them after the last main chunk.  (Spreading them evenly over the chunks - skip_hi(ch) =
  // (ch + 1) * skip_chunks / chunks, so that a skip stage's 36 KB box + 16 KB of weights for only 4 MMAs loads behind main stages -
  // was measured in round 2: the convolutions with a fused 1x1 skip did not change, profiles/r02_conv_ring_experiments.md 2f.)
  auto skip_hi = [&](const int ch) -> int { return ch == chunks - 1 ? g.skip_chunks : 0; };
  auto role_tma = [&]() {
    if (lane == 0) {
      // ring positions are (slot, phase) pairs advanced incrementally: the depths are launch parameters, and a runtime % and /
      // per stage costs ~20 instructions in every role (ncu source page of the first r02 build)
      int as = 0, ws_ = 0;
      uint32_t aph = 0, wph = 0;
      auto load_w = [&](const CUtensorMap *tm, int c0, int co0, int tap0, int ntap, int tap_stride) {
        const int s = ws_;
        mbar_wait(&w_empty[s], wph ^ 1u);
        mbar_expect_tx(&w_full[s], ntap * V3_W_BYTES);
        for (int j = 0; j < ntap; ++j)
          tma_load_3d(w_ring + s * W_SLOT_BYTES + j * V3_W_BYTES, tm, &w_full[s], c0, co0, tap0 + j * tap_stride);
        if (++ws_ == W_SL) { ws_ = 0; wph ^= 1u; }
      };
      // fused kernels: the producers read the raw operand with ordinary global loads whose latency sits on their critical
      // path (load -> activate -> publish); an L2 prefetch of the 18 x 18 halo box `pf_dist` chunks ahead turns it into an
      // L2 hit.  tmA / tmAb describe the one or two raw sources (virtual concat) with a {64 ch, 18 px, 18 rows} box.
      uint32_t tile_it = 0;
      auto prefetch_raw = [&](uint32_t lin) {
        const uint32_t ti = lin / (uint32_t)chunks, ch = lin - ti * (uint32_t)chunks;
        const int64_t tl = (int64_t)blockIdx.x + (int64_t)ti * gridDim.x;
        if (tl >= g.total_tiles) return;
        int n, h0, w0, co0;
        decode((int)tl, n, h0, w0, co0);
        const int c = (int)ch * TC_BK;
        if (gf.up) tma_prefetch_4d(&tmA, c, (w0 >> 1) - 1, (h0 >> 1) - 1, n);       // the 10 x 10 source box of the 18 x 18 halo
        else if (c < gf.C1) tma_prefetch_4d(&tmA, c, w0 - 1, h0 - 1, n);
        else tma_prefetch_4d(&tmAb, c - gf.C1, w0 - 1, h0 - 1, n);
      };
      if (FUSE && g.pf_dist > 0)
        for (int i = 1; i < g.pf_dist; ++i) prefetch_raw((uint32_t)i);
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tile_it) {
        int n, h0, w0, co0;
        decode(tile, n, h0, w0, co0);
        int sk = 0;                            // next fused-skip chunk
        for (int ch = 0; ch < chunks; ++ch) {
          for (int dx = 0; dx < 3; ++dx) {
            const int s = as;
            mbar_wait(&a_empty[s], aph ^ 1u);
            if (++as == A_SL) { as = 0; aph ^= 1u; }
            uint8_t *sa = a_ring + s * V2_A_BYTES;
            if (FUSE) {
              mbar_arrive(&a_full[s]);           // the operand is written by the producer warps
              if (dx == 0 && g.pf_dist > 0) prefetch_raw(tile_it * (uint32_t)chunks + (uint32_t)ch + (uint32_t)g.pf_dist);
            } else {
              mbar_expect_tx(&a_full[s], V2_A_BYTES);
              tma_load_4d(sa, &tmA, &a_full[s], ch * TC_BK, w0 + dx - 1, h0 - 1, n);
            }
            for (int dy = 0; dy < 3; dy += W_TP) load_w(&tmB, ch * TC_BK, co0, dy * 3 + dx, W_TP, 3);
          }
          for (const int sk_hi = skip_hi(ch); sk < sk_hi; ++sk) {
            const int s = as;
            mbar_wait(&a_empty[s], aph ^ 1u);
            if (++as == A_SL) { as = 0; aph ^= 1u; }
            uint8_t *sa = a_ring + s * V2_A_BYTES;
            mbar_expect_tx(&a_full[s], V2_A_BYTES);
            if (sk < g.skip_chunks_a) tma_load_4d(sa, &tmA2, &a_full[s], sk * TC_BK, w0, h0 - 1, n);
            else tma_load_4d(sa, &tmA2b, &a_full[s], (sk - g.skip_chunks_a) * TC_BK, w0, h0 - 1, n);
            load_w(&tmB2, sk * TC_BK, co0, 0, 1, 0);
          }
        }
      }
    }
  };
  auto role_mma = [&]() {
    if (lane == 0) {
      uint32_t tc = 0;
      int as = 0, ws_ = 0;
      uint32_t rph = 0, wph = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tc) {
        const uint32_t ab = tc & 1u, aph = (tc >> 1) & 1u;
        mbar_wait(&tmem_empty[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + ab * 256u;
        uint32_t acc = 0;                       // 0 for the tile's very first MMA (overwrite), 1 afterwards
        // one (chunk, dx) main stage: the three dy views of the halo box against 3 / W_TP weight slots; one fused-skip stage: the
        // dy = 1 view against one 1x1 weight tile
        auto stage = [&](const bool main_stage) {
          const int s = as;
          mbar_wait(&a_full[s], rph);
          if (++as == A_SL) { as = 0; rph ^= 1u; }
          tc_fence_after();
          const uint32_t sa = smem_u32(a_ring + s * V2_A_BYTES);
          const int dy_lo = main_stage ? 0 : 1, dy_hi = main_stage ? 3 : 2;
          for (int dy = dy_lo; dy < dy_hi;) {
            const int ws = ws_;
            mbar_wait(&w_full[ws], wph);
            if (++ws_ == W_SL) { ws_ = 0; wph ^= 1u; }
            tc_fence_after();
            const int nt = main_stage ? W_TP : 1;
            for (int j = 0; j < nt; ++j, ++dy) {
              if (g.dbg & 2) continue;
              const uint64_t wdesc = make_sw128_desc(smem_u32(w_ring + ws * W_SLOT_BYTES + j * V3_W_BYTES));   // A: 128 couts x 64 k
              const uint64_t pdesc = make_sw128_desc(sa + (uint32_t)(dy * 16 * 128));                          // B: 256 pixels x 64 k
#pragma unroll
              for (int kk = 0; kk < TC_BK / 16; ++kk) {
                umma_f16(d, wdesc + (uint64_t)(2 * kk), pdesc + (uint64_t)(2 * kk), IDESC, acc);
                acc = 1u;
              }
            }
            umma_commit(&w_empty[ws]);
          }
          umma_commit(&a_empty[s]);
        };
        int sk = 0;
        for (int ch = 0; ch < chunks; ++ch) {
          for (int dx = 0; dx < 3; ++dx) stage(true);
          for (const int sk_hi = skip_hi(ch); sk < sk_hi; ++sk) stage(false);
        }
        umma_commit(&tmem_full[ab]);
      }
    }
  };
  auto role_producer = [&]() {
    // ---------------------------------------------------------------- v4 operand producers (8 warps, 256 threads)
    // The 18 x 16 pixel halo box of a stage is 2304 16-byte units (pixel row r = yy*16 + xx, channel octet j); thread p owns
    // j = p & 7, xx = (p >> 3) & 15 and the 9 rows yy = 2i + (p >> 7): the channel octet (hence the 8 coefficient pairs), the
    // column validity and the swizzle term (r & 7 == xx & 7) are per-thread constants; a warp reads 4 pixels x 128 contiguous B.
    // Per stage: wait until the MMAs that read the slot have retired -> transform the 9 raw units held in registers -> st.shared
    // -> fence.proxy.async -> arrive -> issue the global loads of the next stage into the same registers.  The loads come
    // AFTER the fence on purpose: it compiles to MEMBAR.ALL.CTA, which waits for every load in flight.
    // The SM sub-partitions are issue-bound in this kernel (2 producer + 2 epilogue warps each, ncu: ~0.9 IPC), so the stage
    // is written for instruction count: branch-free (validity is a bit mask that predicates the loads and, only in tiles that
    // touch the image border - warp-uniform test - selects zeros), one register buffer, addresses advanced incrementally.
    const int p = (int)threadIdx.x - 12 * 32;
    const int j = p & 7, xx = (p >> 3) & 15, yh = p >> 7;
    const uint32_t dst_off = (uint32_t)(yh * 2048 + xx * 128 + ((j ^ (xx & 7)) << 4));
    const int Ctot = gf.C1 + gf.C2;
    uint4 buf[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) buf[i] = make_uint4(0u, 0u, 0u, 0u);
    uint32_t mask = 0;                    // validity bits of the 9 units in buf (outside the image -> exact zeros)
    float hs[8], ho[8];
    // state of the tile being LOADED (it runs one stage ahead of the tile being transformed)
    int ln = 0, lw0 = 0;
    uint32_t lym = 0;
    int64_t lbase = 0;                    // pixel index of (n, h0 - 1 + yh, w0 - 1 + xx)
    auto load_tile = [&](int tile) {
      int h0, co0;
      decode(tile, ln, h0, lw0, co0);
      lym = 0;
#pragma unroll
      for (int i = 0; i < 9; ++i) lym |= ((unsigned)(h0 - 1 + yh + 2 * i) < (unsigned)g.H) ? (1u << i) : 0u;
      lbase = ((int64_t)ln * g.H + (h0 - 1 + yh)) * g.W + (lw0 - 1 + xx);
    };
    // issue the (up to) 9 global loads of main stage k of the tile described by (lbase, lym, lw0)
    auto issue = [&](int k) {
      const int ch = k / 3, dx = k - ch * 3;
      const int c = ch * TC_BK + j * 8;
      mask = ((unsigned)(lw0 + dx - 1 + xx) < (unsigned)g.W) ? lym : 0u;
      if (g.dbg & 32) mask = 0u;          // timing diagnostic: no global loads
      const __half *ptr;
      int64_t rs;
      if (c < gf.C1) { ptr = gf.src1 + (lbase + dx) * gf.C1 + c; rs = (int64_t)2 * g.W * gf.C1; }
      else { ptr = gf.src2 + (lbase + dx) * gf.C2 + (c - gf.C1); rs = (int64_t)2 * g.W * gf.C2; }
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        if (mask & (1u << i)) buf[i] = __ldg(reinterpret_cast<const uint4 *>(ptr));
        ptr += rs;
      }
    };
    // SiLU(GroupNorm(x)) of one 16-byte unit (8 channels)
    auto act8 = [&](const uint4 &raw) {
      uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __half2 hv = *reinterpret_cast<const __half2 *>(&w[q]);
        __half2 r;
        if (FUSE == 2) {
          const __half2 h = __hfma2(hv, __floats2half2_rn(hs[2 * q], hs[2 * q + 1]), __floats2half2_rn(ho[2 * q], ho[2 * q + 1]));
          r = __hfma2(h, tanh_approx_h2(h), h);
        } else {
          const float2 f = __half22float2(hv);
          const float ha = fmaf(f.x, hs[2 * q], ho[2 * q]), hb = fmaf(f.y, hs[2 * q + 1], ho[2 * q + 1]);
          r = __floats2half2_rn(fmaf(ha, tanh_approx_f(ha), ha), fmaf(hb, tanh_approx_f(hb), hb));
        }
        w[q] = *reinterpret_cast<const uint32_t *>(&r);
      }
      return make_uint4(w[0], w[1], w[2], w[3]);
    };
    int tile = blockIdx.x;
    if (tile < g.total_tiles) {
      load_tile(tile);
      issue(0);
    }
    int rs = 0;
    uint32_t rph = 0;           // activation-ring position (slot, phase), advanced once per stage
    for (; tile < g.total_tiles; tile += gridDim.x) {
      const int n = ln;                   // image of the tile being transformed (the loader moves on during its last stage)
      for (int k = 0; k < num_k; ++k) {
        const int s = rs;
        const uint32_t ph = rph;
        if (++rs == A_SL) { rs = 0; rph ^= 1u; }
        if (k >= num_main) {               // fused 1x1-skip stage: operand comes by TMA, only keep the barrier count
          mbar_wait(&a_empty[s], ph ^ 1u);
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[s]);
          continue;
        }
        const int ch = k / 3;
        if (k == ch * 3) {                 // new 64-channel chunk: this thread's 8 coefficient pairs
          const float *cf = gf.coef + (int64_t)n * 2 * Ctot + ch * TC_BK + j * 8;
          const float4 a0 = __ldg(reinterpret_cast<const float4 *>(cf)), a1 = __ldg(reinterpret_cast<const float4 *>(cf + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4 *>(cf + Ctot)), b1 = __ldg(reinterpret_cast<const float4 *>(cf + Ctot + 4));
          hs[0] = a0.x; hs[1] = a0.y; hs[2] = a0.z; hs[3] = a0.w; hs[4] = a1.x; hs[5] = a1.y; hs[6] = a1.z; hs[7] = a1.w;
          ho[0] = b0.x; ho[1] = b0.y; ho[2] = b0.z; ho[3] = b0.w; ho[4] = b1.x; ho[5] = b1.y; ho[6] = b1.z; ho[7] = b1.w;
        }
        mbar_wait(&a_empty[s], ph ^ 1u);
        uint8_t *dst = a_ring + s * V2_A_BYTES + dst_off;
        if (g.dbg & 64) {                   // timing diagnostic: raw copy, no arithmetic
#pragma unroll
          for (int i = 0; i < 9; ++i) *reinterpret_cast<uint4 *>(dst + i * 4096) = buf[i];
        } else if (__all_sync(0xffffffffu, mask == 0x1FFu)) {   // interior: every unit of the warp is inside the image
#pragma unroll
          for (int i = 0; i < 9; ++i) *reinterpret_cast<uint4 *>(dst + i * 4096) = act8(buf[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            uint4 v = act8(buf[i]);
            if (!(mask & (1u << i))) v = make_uint4(0u, 0u, 0u, 0u);    // select, not a branch
            *reinterpret_cast<uint4 *>(dst + i * 4096) = v;
          }
        }
        if (!(g.dbg & 128)) fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[s]);
        // the next main stage (same tile, or the first stage of this CTA's next tile)
        if (k + 1 < num_main) {
          issue(k + 1);
        } else if (tile + (int)gridDim.x < g.total_tiles) {
          load_tile(tile + (int)gridDim.x);
          issue(0);
        }
      }
    }
  };
  auto role_producer_once = [&]() {
    // ---------------------------------------------------------------- v5 operand producers (FUSE == 3; 9 warps, 288 threads)
    // The default since round 2 (dpir_set_gn_fusion 1 / 3).
    // k_conv3x3_v3<1> activates every input element once per horizontal tap box: 3 x 18/16 = 3.4 times.  Here a thread owns a
    // column X of the 18-wide halo (instead of a column of one 16-wide box), a channel octet j and every second row (rg):
    // 18 x 8 x 2 = 288 threads x 9 units.  The 9 raw units of a 64-channel chunk are loaded once, activated once in place,
    // and kept in registers across the chunk's three stages: for tap column dx the thread stores them at box column
    // xx = X - dx (threads whose xx falls outside [0,16) skip that stage's stores).  The loads of the NEXT chunk are issued
    // after the dx = 2 stage has been published (never in flight at a proxy fence = MEMBAR.ALL.CTA); the producers then run
    // up to two cheap stages ahead of the tensor core, which leaves ~2 MMA stages for load latency + activation.
    const int t = (int)threadIdx.x - 8 * 32;      // 0..287
    const int j = t & 7, Xl = t >> 3;             // Xl 0..35
    const int X = Xl % 18, rg = Xl / 18;
    const int Ctot = gf.C1 + gf.C2;
    uint4 buf[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) buf[i] = make_uint4(0u, 0u, 0u, 0u);
    uint32_t mask = 0, lmask = 0;          // validity of the units in buf / of the tile being loaded
    int ln = 0;
    int64_t lbase = 0;                     // pixel index of (n, h0 - 1 + rg, w0 - 1 + X) of the tile being loaded
    auto load_tile = [&](int tile) {
      int h0, w0, co0;
      decode(tile, ln, h0, w0, co0);
      lmask = 0;
      if ((unsigned)(w0 - 1 + X) < (unsigned)g.W) {
#pragma unroll
        for (int i = 0; i < 9; ++i) lmask |= ((unsigned)(h0 - 1 + rg + 2 * i) < (unsigned)g.H) ? (1u << i) : 0u;
      }
      if (g.dbg & 32) lmask = 0u;
      if (gf.up) {
        // nearest-neighbour up-sampling folded into the operand: halo pixel (Y, X) of the OUTPUT grid reads source pixel
        // (Y >> 1, X >> 1) of the half-resolution tensor; this thread's rows Y = h0 - 1 + rg + 2i map to consecutive source rows
        // h0/2 - 1 + rg + i (h0 is a multiple of 16), validity stays an output-grid test (the conv pads the UP-SAMPLED tensor)
        lbase = ((int64_t)ln * (g.H >> 1) + ((h0 >> 1) - 1 + rg)) * (g.W >> 1) + ((w0 - 1 + X) >> 1);
      } else {
        lbase = ((int64_t)ln * g.H + (h0 - 1 + rg)) * g.W + (w0 - 1 + X);
      }
    };
    auto issue = [&](int ch) {             // raw units of chunk ch of the tile described by (lbase, lmask)
      const int c = ch * TC_BK + j * 8;
      const __half *ptr;
      int64_t rs;
      if (gf.up) { ptr = gf.src1 + lbase * gf.C1 + c; rs = (int64_t)(g.W >> 1) * gf.C1; }
      else if (c < gf.C1) { ptr = gf.src1 + lbase * gf.C1 + c; rs = (int64_t)2 * g.W * gf.C1; }
      else { ptr = gf.src2 + lbase * gf.C2 + (c - gf.C1); rs = (int64_t)2 * g.W * gf.C2; }
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        if (lmask & (1u << i)) buf[i] = __ldg(reinterpret_cast<const uint4 *>(ptr));
        ptr += rs;
      }
    };
    int tile = blockIdx.x;
    if (tile < g.total_tiles) {
      load_tile(tile);
      issue(0);
    }
    int rs = 0;
    uint32_t rph = 0;           // activation-ring position (slot, phase), advanced once per stage
    for (; tile < g.total_tiles; tile += gridDim.x) {
      const int n = ln;
      mask = lmask;
      int sk = 0;
      for (int ch = 0; ch < chunks; ++ch) {
        if (!(g.dbg & 64)) {    // new chunk: activate the 9 units in place (needs no pipeline slot)
          const float *cf = gf.coef + (int64_t)n * 2 * Ctot + ch * TC_BK + j * 8;
          const float4 a0 = __ldg(reinterpret_cast<const float4 *>(cf)), a1 = __ldg(reinterpret_cast<const float4 *>(cf + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4 *>(cf + Ctot)), b1 = __ldg(reinterpret_cast<const float4 *>(cf + Ctot + 4));
          const float hs[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float ho[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            uint32_t w[4] = {buf[i].x, buf[i].y, buf[i].z, buf[i].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
              const float ha = fmaf(f.x, hs[2 * q], ho[2 * q]), hb = fmaf(f.y, hs[2 * q + 1], ho[2 * q + 1]);
              const __half2 r = __floats2half2_rn(fmaf(ha, tanh_approx_f(ha), ha), fmaf(hb, tanh_approx_f(hb), hb));
              w[q] = *reinterpret_cast<const uint32_t *>(&r);
            }
            const bool ok = (mask >> i) & 1u;              // outside the image: the conv pads the ACTIVATED tensor with zeros
            buf[i] = make_uint4(ok ? w[0] : 0u, ok ? w[1] : 0u, ok ? w[2] : 0u, ok ? w[3] : 0u);
          }
        }
#pragma unroll 1
        for (int dx = 0; dx < 3; ++dx) {
          const int s = rs;
          const uint32_t ph = rph;
          if (++rs == A_SL) { rs = 0; rph ^= 1u; }
          mbar_wait(&a_empty[s], ph ^ 1u);
          const int xx = X - dx;
          if ((unsigned)xx < 16u) {
            uint8_t *dst = a_ring + s * V2_A_BYTES + (rg * 2048 + xx * 128 + ((j ^ (xx & 7)) << 4));
#pragma unroll
            for (int i = 0; i < 9; ++i) *reinterpret_cast<uint4 *>(dst + i * 4096) = buf[i];
          }
          if (!(g.dbg & 128)) fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[s]);
        }
        // the chunk is published three times: fetch the next one into the same registers (never in flight at a proxy fence)
        if (ch + 1 < chunks) {
          issue(ch + 1);
        } else if (tile + (int)gridDim.x < g.total_tiles) {
          load_tile(tile + (int)gridDim.x);
          issue(0);
        }
        for (const int sk_hi = skip_hi(ch); sk < sk_hi; ++sk) {   // fused 1x1-skip stages: operand comes by TMA, only keep the barrier count
          const int s = rs;
          const uint32_t ph = rph;
          if (++rs == A_SL) { rs = 0; rph ^= 1u; }
          mbar_wait(&a_empty[s], ph ^ 1u);
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[s]);
        }
      }
    }
  };
  auto role_epilogue = [&](const int ew) {
    // ---------------------------------------------------------------- epilogue: lane = output channel, column = pixel
    // fp16 in / fp16 out only (the fp16-stream pipeline).  The epilogue is bound by per-warp instruction latency (ncu: ~6
    // cycles per issued instruction, 2 warps per scheduler), so it is built for minimum instruction count: no global
    // load/store instructions, no bounds logic.  Per 32-pixel chunk a warp adds bias (+ residual) to its 32 ch x 32 px
    // accumulator slice, writes it as fp16 into a [pixel][channel] 2 KB tile of shared memory (32 x STS.U16, each one 64
    // contiguous bytes) and ONE thread issues a TMA tensor store of the {32 ch, 16 px, 2 rows} box; the hardware clips boxes
    // that stick out of the image.  The residual box arrives the same way (TMA load, double buffered, prefetched one chunk
    // ahead) and is read back column-wise (32 x LDS.U16).
    const int q = warp & 3;               // TMEM lane quarter -> channels co0 + q*32 + lane
    const int hv = ew >> 2;               // which half of the tile's 256 pixels (image rows hv*8 .. hv*8+7)
    __half *outb = reinterpret_cast<__half *>(stg_base + ew * g.stg_stride);
    __half *resb0 = outb + 1024, *resb1 = outb + 2048;
    uint64_t *rbar = res_full + ew * 2;
    uint32_t rcount = 0;                  // residual boxes consumed so far by this warp (buffer = rcount & 1, parity = (rcount >> 1) & 1)
    const bool has_res = epi.residual != nullptr;
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tc) {
      const uint32_t ab = tc & 1u, aph = (tc >> 1) & 1u;
      int n, h0, w0, co0;
      decode(tile, n, h0, w0, co0);
      const int cbase = co0 + q * 32;
      float bch = __ldg(epi.bias + cbase + lane);
      if (epi.bias2) bch += __ldg(epi.bias2 + cbase + lane);
      float st_s = 0.f, st_q = 0.f;
      const bool full_tile = (h0 + 16 <= g.H) && (w0 + 16 <= g.W);
      const int ybase = h0 + hv * 8;
      // residual boxes: {32 ch, 16 px, 2 rows} of the output grid, or - res_up - the {32 ch, 8 px, 1 row} source box of the
      // half-resolution tensor that the two output rows y0, y0 + 1 (y0 even) up-sample
      const int res_up = epi.res_up;
      const uint32_t res_bytes = res_up ? 512u : 2048u;
      if (has_res && lane == 0) {          // prefetch the residual box of chunk 0
        uint64_t *bar = &rbar[rcount & 1];
        mbar_expect_tx(bar, res_bytes);
        if (res_up) tma_load_4d((rcount & 1) ? resb1 : resb0, &tmRes, bar, cbase, w0 >> 1, ybase >> 1, n);
        else tma_load_4d((rcount & 1) ? resb1 : resb0, &tmRes, bar, cbase, w0, ybase, n);
      }
      mbar_wait_relaxed(&tmem_full[ab], aph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 256u + (uint32_t)hv * 128u;
#pragma unroll 1
      for (int c = 0; c < 4 && !(g.dbg & 1); ++c) {       // 32 pixels = image rows y0, y0+1 of the tile
        const int y0 = ybase + 2 * c;
        float v[32];
        if (has_res) {
          if (c < 3 && lane == 0) {        // prefetch the next chunk's box into the other buffer (freed one iteration ago)
            uint64_t *bar = &rbar[(rcount + 1) & 1];
            mbar_expect_tx(bar, res_bytes);
            if (res_up) tma_load_4d(((rcount + 1) & 1) ? resb1 : resb0, &tmRes, bar, cbase, w0 >> 1, (y0 + 2) >> 1, n);
            else tma_load_4d(((rcount + 1) & 1) ? resb1 : resb0, &tmRes, bar, cbase, w0, y0 + 2, n);
          }
          mbar_wait(&rbar[rcount & 1], (rcount >> 1) & 1u);
          const __half *rb = (rcount & 1) ? resb1 : resb0;
          if (res_up) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __half2float(rb[((j & 15) >> 1) * 32 + lane]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __half2float(rb[j * 32 + lane]);
          }
          ++rcount;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        uint32_t r[32];
        tmem_ld32(tbase + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c == 3) {                      // accumulator fully read: hand it back to the MMA warp before the tail work
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[ab]);
        }
        if (full_tile) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] += __uint_as_float(r[j]) + bch;
            st_s += v[j];
            st_q = fmaf(v[j], v[j], st_q);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] += __uint_as_float(r[j]) + bch;
            if ((y0 + (j >> 4) < g.H) && (w0 + (j & 15) < g.W)) { st_s += v[j]; st_q = fmaf(v[j], v[j], st_q); }
          }
        }
        // the previous chunk's TMA store must have finished reading the staging tile before it is overwritten
        if (lane == 0) tma_store_wait_read0();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) outb[j * 32 + lane] = __float2half_rn(v[j]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(g.dbg & 16)) {
          tma_store_4d(&tmOut, outb, cbase, w0, y0, n);
          tma_store_commit();
        }
      }
      if (g.dbg & 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[ab]);
      }
      if (epi.qstats && !(g.dbg & 8)) {
        st_s += __shfl_xor_sync(0xffffffffu, st_s, 1); st_q += __shfl_xor_sync(0xffffffffu, st_q, 1);
        st_s += __shfl_xor_sync(0xffffffffu, st_s, 2); st_q += __shfl_xor_sync(0xffffffffu, st_q, 2);
        if ((lane & 3) == 0) {
          double *qp = epi.qstats + ((int64_t)n * (g.Cout >> 2) + ((cbase + lane) >> 2)) * 2;
          atomicAdd(qp, (double)st_s);
          atomicAdd(qp + 1, (double)st_q);
        }
      }
    }
    if (lane == 0) tma_store_wait_all0();      // all tensor stores of this warp have landed before the CTA retires
  };
  if constexpr (FUSE == 3) {
    // v5 layout: warpgroups 0-1 = epilogue (120 registers), warpgroups 2-4 = 9 producer warps + TMA warp + MMA warp + one
    // idle warp (80 registers): 2*128*120 + 3*128*80 = 640*96
    if (warp < 8) {
      setmaxnreg_inc<120>();
      role_epilogue(warp);
    } else {
      setmaxnreg_dec<80>();
      if (warp < 17) role_producer_once();
      else if (warp == 17) role_tma();
      else if (warp == 18) role_mma();
    }
  } else if constexpr (FUSE != 0) {
    // setmaxnreg is executed once per warpgroup, at the top of the branch that dominates all of that group's code
    if (warp < 4) {
      setmaxnreg_dec<32>();
      if (warp == 0) role_tma();
      else if (warp == 1) role_mma();
    } else if (warp < 12) {
      setmaxnreg_inc<120>();
      role_epilogue(warp - 4);
    } else {
      setmaxnreg_inc<104>();
      role_producer();
    }
  } else {
    if (warp == 0) role_tma();
    else if (warp == 1) role_mma();
    else role_epilogue(warp - 2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ narrow-output 3x3 kernel
// The UNet's output layer (unet.py:617-621: conv 3x3, model_channels -> 6) has too few output channels for the swap-AB kernel (M =
// couts) and ran on k_conv_tc<16, 4>, which fetches the 128-pixel activation tile once per tap: 9x the tensor through L2 (ncu/r02:
// 0.42 ms for a 0.54 GB tensor = 11.6 TB/s of L2->SM traffic, the crossbar's limit).  k_conv3x3_dxn reads every halo box ONCE per
// 64-channel chunk by moving the horizontal taps into the N dimension of the MMA:
//   * M = 128 box pixels (8 image rows x the 16 columns of the box), N = 48 = (dx, cout padded to 16), K = 64 channels: for the
//     vertical tap dy the A operand is the view of the {64 ch, 16 px, 18 rows} halo box starting dy (+ 8 for the lower half-tile)
//     rows down (2048 B: still SWIZZLE_128B aligned), the B operand the 48 weight rows [dx][16 couts] of that dy - which are
//     CONTIGUOUS in the packed weight tensor [tap = dy*3 + dx][16][Cin], so a {64, 48, 1} TMA box over the same buffer delivers them;
//   * D[pixel p][dx*16 + o] then holds tap (dy summed, dx) applied AT box pixel p; the output at box column xb is
//     D[xb - 1][0*16 + o] + D[xb][1*16 + o] + D[xb + 1][2*16 + o]: two warp shuffles per channel in the epilogue (TMEM lane = box
//     pixel, 16 consecutive lanes = one image row).  Box columns 0 and 15 are halo only: a tile produces 14 x 16 output pixels;
//   * all weights (chunks x 3 x 6 KB) stay resident in shared memory; the activation ring holds 3-4 halo boxes; persistent CTAs,
//     two TMEM accumulator sets so the epilogue of tile i overlaps the loads / MMAs of tile i + 1.
// L2->SM traffic per output pixel: 36 KB x chunks / 224 px instead of 9 x 16 KB x chunks / 128 px (3.5x less).
constexpr int DXN_THREADS = 320;                      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int DXN_W_BYTES = 48 * 128;                 // one (chunk, dy) weight tile: 48 rows x 64 fp16
constexpr int DXN_MAX_CHUNKS = 6;
constexpr int DXN_ACT_WARPS = 8;                      // FUSE: warps 10-17 activate the halo boxes in place
struct ConvDxnGeom {
  int B, H, W, Cout, chunks, a_wrap;
  int tiles_w, tiles_h, total_tiles;                 // 14 x 16 output-pixel tiles
  int a_slots;
  const float *coef;                                 // FUSE: [B][2][C] GroupNorm(+FiLM) coefficients of launch_gn_coef (C = chunks * 64)
};

// FUSE: the operand is SiLU(GroupNorm(x)) of the RAW tensor the TMA fetched (the UNet's out.0 / out.1, unet.py:617-619): eight
// activator warps rewrite every halo box in place (thread = channel octet x box pixel mod 32: 9 x LDS.128 -> 8 x (fma, tanh, fma)
// -> STS.128 at the same swizzled address), write exact zeros outside the image (the convolution pads the ACTIVATED tensor; the
// TMA's zero fill would activate to SiLU(offset)), and hand the box to the MMA warp through fence.proxy.async + a second barrier.
template <bool FUSE>
__global__ void __launch_bounds__(FUSE ? DXN_THREADS + 32 * DXN_ACT_WARPS : DXN_THREADS, 1)
    k_conv3x3_dxn(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvDxnGeom g, ConvEpi epi) {
  // M = 128 (box pixels), N = 48 (dx x 16 couts)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t *a_full = reinterpret_cast<uint64_t *>(smem);
  uint64_t *a_empty = a_full + 4;
  uint64_t *a_ready = a_empty + 4;             // FUSE: box activated (one arrive per activator warp)
  uint64_t *w_full = a_ready + 4;
  uint64_t *tmem_full = w_full + 1;            // [2]
  uint64_t *tmem_empty = tmem_full + 2;        // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  uint8_t *w_sm = smem + 1024;                                   // [chunks][3 dy][48 x 128 B]
  uint8_t *a_ring = w_sm + g.chunks * 3 * DXN_W_BYTES;           // [a_slots][V2_A_BYTES]
  const int A_SL = g.a_slots;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < A_SL; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&a_ready[s], DXN_ACT_WARPS); }
    mbar_init(w_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int tile, int &n, int &h0, int &w0) {
    w0 = (tile % g.tiles_w) * 14;
    h0 = ((tile / g.tiles_w) % g.tiles_h) * 16;
    n = tile / (g.tiles_w * g.tiles_h);
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)(g.chunks * 3 * DXN_W_BYTES));
      for (int ch = 0; ch < g.chunks; ++ch)
        for (int dy = 0; dy < 3; ++dy) tma_load_3d(w_sm + (ch * 3 + dy) * DXN_W_BYTES, &tmW, w_full, ch * TC_BK, 0, dy);
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
        int n, h0, w0;
        decode(tile, n, h0, w0);
        for (int ch = 0; ch < g.chunks; ++ch) {
          const int s = as;
          mbar_wait(&a_empty[s], aph ^ 1u);
          if (++as == A_SL) { as = 0; aph ^= 1u; }
          mbar_expect_tx(&a_full[s], V2_A_BYTES);
          const int ach = ch >= g.a_wrap ? ch - g.a_wrap : ch;
          tma_load_4d(a_ring + s * V2_A_BYTES, &tmA, &a_full[s], ach * TC_BK, w0 - 1, h0 - 1, n);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(w_full, 0);
      tc_fence_after();
      uint32_t tc = 0;
      int as = 0;
      uint32_t rph = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tc) {
        const uint32_t ab = tc & 1u, aph = (tc >> 1) & 1u;
        mbar_wait(&tmem_empty[ab], aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + ab * 128u;
        for (int ch = 0; ch < g.chunks; ++ch) {
          const int s = as;
          mbar_wait(FUSE ? &a_ready[s] : &a_full[s], rph);
          if (++as == A_SL) { as = 0; rph ^= 1u; }
          tc_fence_after();
          const uint32_t sa = smem_u32(a_ring + s * V2_A_BYTES);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t wdesc = make_sw128_desc(smem_u32(w_sm + (ch * 3 + dy) * DXN_W_BYTES));      // B: 48 (dx, cout) rows x 64 k
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              const uint64_t pdesc = make_sw128_desc(sa + (uint32_t)((dy + mb * 8) * 16 * 128));       // A: 128 box pixels x 64 k
#pragma unroll
              for (int kk = 0; kk < TC_BK / 16; ++kk)
                umma_f16(d + (uint32_t)mb * 64u, pdesc + (uint64_t)(2 * kk), wdesc + (uint64_t)(2 * kk), IDESC, (ch | dy | kk) ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&tmem_full[ab]);
      }
    }
  } else if (FUSE && warp >= 10) {
    // ---------------------------------------------------------------- activators: 256 threads, 9 units of 16 B each per box
    const int t = (int)threadIdx.x - DXN_THREADS;
    const int j = t & 7, pix = t >> 3;                   // channel octet, box pixel mod 32 (box row 2i + (pix >> 4), column pix & 15)
    const int xb = pix & 15, yo = pix >> 4;
    const uint32_t off = (uint32_t)(pix * 128 + ((j ^ (pix & 7)) << 4));
    const int C = g.chunks * TC_BK;
    int as = 0;
    uint32_t rph = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      int n, h0, w0;
      decode(tile, n, h0, w0);
      uint32_t vmask = 0;                                // validity of this thread's 9 box rows (all-zero when its column is outside)
      if ((unsigned)(w0 - 1 + xb) < (unsigned)g.W) {
#pragma unroll
        for (int i = 0; i < 9; ++i) vmask |= ((unsigned)(h0 - 1 + yo + 2 * i) < (unsigned)g.H) ? (1u << i) : 0u;
      }
      for (int ch = 0; ch < g.chunks; ++ch) {
        const int s = as;
        const uint32_t ph = rph;
        if (++as == A_SL) { as = 0; rph ^= 1u; }
        const float *cf = g.coef + (int64_t)n * 2 * C + ch * TC_BK + j * 8;
        const float4 a0 = __ldg(reinterpret_cast<const float4 *>(cf)), a1 = __ldg(reinterpret_cast<const float4 *>(cf + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(cf + C)), b1 = __ldg(reinterpret_cast<const float4 *>(cf + C + 4));
        const float hs[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float ho[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        mbar_wait(&a_full[s], ph);
        uint8_t *base = a_ring + s * V2_A_BYTES + off;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          uint4 *up = reinterpret_cast<uint4 *>(base + i * 4096);
          const uint4 raw = *up;
          uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
            const float ha = fmaf(f.x, hs[2 * q], ho[2 * q]), hb = fmaf(f.y, hs[2 * q + 1], ho[2 * q + 1]);
            const __half2 r = __floats2half2_rn(fmaf(ha, tanh_approx_f(ha), ha), fmaf(hb, tanh_approx_f(hb), hb));
            w[q] = *reinterpret_cast<const uint32_t *>(&r);
          }
          const bool ok = (vmask >> i) & 1u;
          *up = make_uint4(ok ? w[0] : 0u, ok ? w[1] : 0u, ok ? w[2] : 0u, ok ? w[3] : 0u);
        }
        fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[s]);
      }
    }
  } else if (warp < 10) {
    // ---------------------------------------------------------------- epilogue: TMEM lane = box pixel (row yy = lane quarter*2 +
    // lane/16 of the half-tile mb, column xb = lane % 16)
    const int q = warp & 3, mb = (warp - 2) >> 2;
    const int xb = lane & 15, yy = q * 2 + (lane >> 4);
    const int ncols = g.Cout;
    uint32_t tc = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++tc) {
      const uint32_t ab = tc & 1u, aph = (tc >> 1) & 1u;
      int n, h0, w0;
      decode(tile, n, h0, w0);
      mbar_wait_relaxed(&tmem_full[ab], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 128u + (uint32_t)mb * 64u;
      uint32_t r0[16], r1[16], r2[16];
      tmem_ld16(taddr, r0);
      tmem_ld16(taddr + 16u, r1);
      tmem_ld16(taddr + 32u, r2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[ab]);
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[j]), 1);      // tap dx = 0 evaluated at box column xb - 1
        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 1);   // tap dx = 2 evaluated at box column xb + 1
        acc[j] = (left + __uint_as_float(r1[j])) + right;
      }
      const int x = w0 + xb - 1, y = h0 + mb * 8 + yy;
      if (xb >= 1 && xb <= 14 && x < g.W && y < g.H) conv_epilogue_store(epi, n, y, x, g.H, g.W, g.Cout, 0, acc, ncols);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host side
struct ConvTcPlan {
  CUtensorMap tmA, tmB;
  ConvTcGeom g;
  int BN, CoutPad;
  bool split = false;     // split-precision operands ([hi | lo] activations, [hi | hi | lo] weights)
  bool v2 = false;        // persistent 3x3 kernel (k_conv3x3_v2 / k_conv3x3_v3)
  bool use_v3 = true;     // swap-AB variant
  ConvV2Geom g2;
  CUtensorMap tmA2, tmB2; // fused 1x1 skip operand / weights (v2 only; copies of tmA/tmB when unused)
  CUtensorMap tmA2b;      // second source of a concatenated skip operand (v3/v4)
  CUtensorMap tmAb;       // fused kernels: raw source 2 of the virtual concat (L2 prefetch boxes); tmA then describes source 1
  GnFuse gf{};            // v4: fused GroupNorm+SiLU on the activation operand (gf.coef != null)
  int gf_math = 3;        // producer flavour of the fused kernel (k_conv3x3_v3<gf_math>): 3 activate-once, 1 per-tap fp32, 2 per-tap half
  CUtensorMap tmOut, tmRes;   // v3: fp16 output / residual tensors for the TMA-store epilogue
  bool io_half = false;       // tmOut (and tmRes when has_res) are valid
  bool has_res = false;
  bool res_up = false;        // tmRes describes the half-resolution source of an up-sampled residual
  const void *io_out = nullptr, *io_res = nullptr;
  bool dxn = false;           // narrow-output 3x3 kernel (k_conv3x3_dxn): tmA has the halo box, tmW3 the {64, 48, 1} weight box
  CUtensorMap tmW3;
  ConvDxnGeom gd{};
};

static int g_force_v1 = 0;   // dpir_conv_tc_variant: 0 = auto (persistent swap-AB v3), 1 = v1 everywhere, 2 = persistent v2
// per-device launch state: cudaFuncSetAttribute and the SM count belong to a device, and one process may drive several
struct DevState { bool attr_v1[3] = {false, false, false}; bool attr_v3 = false; int attr_dxn = 0; int num_sms = 0; };
static DevState g_dev[64];
static DevState *dev_state() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return &g_dev[dev];
}
static int g_gnfuse_math = 3; // fused-operand kernel: 3 (default) = activate-once-per-chunk producers (k_conv3x3_v3<3>, fp32 arithmetic),
                              // 1 = per-tap-box producers, fp32 arithmetic, 2 = the same with packed half (tanh.approx.f16x2)

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static int pow2_ge(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int conv_tc_plan_create(const void *act_half, const void *w_half, int B, int H, int W, int Cin_real, int CoutPad, int ksize,
                        ConvTcPlan **out, int split, int force_bn, int act_c) {
  DPIR_REQUIRE(Cin_real % TC_BK == 0, "conv_tc: Cin must be a multiple of 64 (got %d)", Cin_real);
  const int Cin = split ? 3 * Cin_real : Cin_real;          // K extent of the weights
  // channels per pixel of the activation tensor: [hi | lo] when split; act_c > 0: a NARROWER tensor than the K chunk (the stem's
  // 8-channel operand) - the TMA box still spans 64 channels and the hardware zero-fills the ones that do not exist
  DPIR_REQUIRE(act_c == 0 || (!split && act_c % 8 == 0 && act_c < Cin_real), "conv_tc: bad narrow-operand width %d", act_c);
  const int Cact = split ? 2 * Cin_real : (act_c ? act_c : Cin_real);
  DPIR_REQUIRE(ksize == 1 || ksize == 3, "conv_tc: ksize must be 1 or 3");
  DPIR_REQUIRE(CoutPad == 16 || CoutPad % 128 == 0, "conv_tc: padded Cout must be 16 or a multiple of 128 (got %d)", CoutPad);
  DPIR_REQUIRE((((uintptr_t)act_half | (uintptr_t)w_half) & 15) == 0, "conv_tc: operands must be 16-byte aligned");
  auto encode = get_encode_fn();
  if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return DPIR_E_CUDA; }
  ConvTcPlan *p = new ConvTcPlan();
  ConvTcGeom &g = p->g;
  g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = 0; g.ksize = ksize; g.taps = ksize * ksize;
  g.a_wrap = split ? Cact / TC_BK : 0x7fffffff;
  p->split = split != 0;
  g.TW = W >= 128 ? 128 : pow2_ge(W);
  int th = 128 / g.TW;
  g.TH = th < pow2_ge(H) ? th : pow2_ge(H);
  g.TN = 128 / (g.TW * g.TH);
  g.tiles_w = cdiv(W, g.TW); g.tiles_h = cdiv(H, g.TH); g.tiles_n = cdiv(B, g.TN);
  p->BN = CoutPad == 16 ? 16 : 128;
  p->CoutPad = CoutPad;
  p->v2 = g_force_v1 != 1 && ksize == 3 && H >= 16 && W >= 16 && CoutPad % 128 == 0;
  // one-tile-per-CTA kernel on small maps (8x8, 16x16 1x1): 128-wide cout tiles would leave most of the 148 SMs idle
  if (!p->v2 && CoutPad != 16 && g.tiles_w * g.tiles_h * g.tiles_n * (CoutPad / 128) < 100) p->BN = 64;
  if (force_bn) p->BN = force_bn;
  p->use_v3 = g_force_v1 == 0;
  p->dxn = !force_bn && ksize == 3 && conv_tc_narrow_eligible(H, W, Cin, CoutPad);
  if (p->dxn) {
    ConvDxnGeom &gd = p->gd;
    gd.B = B; gd.H = H; gd.W = W; gd.Cout = 0; gd.chunks = Cin / TC_BK; gd.a_wrap = g.a_wrap;
    gd.tiles_w = cdiv(W, 14); gd.tiles_h = cdiv(H, 16);
    gd.total_tiles = B * gd.tiles_w * gd.tiles_h;
    const int budget = 227 * 1024 - 2048 - gd.chunks * 3 * DXN_W_BYTES;
    gd.a_slots = budget / V2_A_BYTES > 4 ? 4 : budget / V2_A_BYTES;
  }
  if (p->v2) {
    ConvV2Geom &g2 = p->g2;
    g2.B = B; g2.H = H; g2.W = W; g2.Cin = Cin; g2.Cout = 0; g2.skip_chunks = 0; g2.skip_chunks_a = 0; g2.dbg = 0;
    g2.a_wrap = g.a_wrap;
    g2.tiles_w = cdiv(W, 16); g2.tiles_h = cdiv(H, 16); g2.n_tiles = CoutPad / 128;
    g2.total_tiles = B * g2.tiles_w * g2.tiles_h * g2.n_tiles;
    g2.a_slots = V3_A_SLOTS; g2.w_slots = V3_W_SLOTS; g2.w_taps = 1; g2.pf_dist = 0; g2.stg_stride = 6144;
    // (the stem's 8-channel operand: issuing only the first K = 16 step of its zero-filled chunk was measured in round 2 and changes
    // nothing - 0.236 vs 0.228 ms - the stem is bound by its epilogue, 4.2 us per tile)
  }
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cact, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)Cact * 2, (cuuint64_t)W * Cact * 2, (cuuint64_t)H * W * Cact * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.TN};
    if (p->v2 || p->dxn) { box[1] = 16; box[2] = 18; box[3] = 1; }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&p->tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(act_half), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(A) failed: %d", (int)r); delete p; return DPIR_E_CUDA; }
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Cin, (cuuint64_t)CoutPad, (cuuint64_t)g.taps};
    cuuint64_t gstr[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)CoutPad * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)p->BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&p->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(w_half), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r); delete p; return DPIR_E_CUDA; }
  }
  p->tmW3 = p->tmB;
  if (p->dxn) {
    // the packed weights [tap = dy*3 + dx][16][Cin] seen as [dy][dx*16 + cout][Cin]: one box = the 48 rows of a vertical tap
    cuuint64_t gdim[3] = {(cuuint64_t)Cin, 48, 3};
    cuuint64_t gstr[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)48 * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, 48, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&p->tmW3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(w_half), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(W3) failed: %d", (int)r); delete p; return DPIR_E_CUDA; }
  }
  p->tmA2 = p->tmA;
  p->tmA2b = p->tmA;
  p->tmAb = p->tmA;
  p->tmB2 = p->tmB;
  *out = p;
  return DPIR_OK;
}

bool conv_tc_plan_is_v2(const ConvTcPlan *p) { return p->v2; }

// whether a 3x3 convolution with K extent Cin (3x the channels for split-precision operands) and CoutPad padded weight rows runs on
// the narrow-output kernel k_conv3x3_dxn
bool conv_tc_narrow_eligible(int H, int W, int Cin, int CoutPad) {
  static int dxn_on = -1;                // DPIR_OUT_DXN=0: narrow-output convs stay on k_conv_tc<16, 4> (A/B timing)
  if (dxn_on < 0) { const char *e = getenv("DPIR_OUT_DXN"); dxn_on = (e && e[0] == '0') ? 0 : 1; }
  return dxn_on && g_force_v1 == 0 && CoutPad == 16 && H >= 16 && W >= 16 && Cin / TC_BK <= DXN_MAX_CHUNKS;
}

// register the (fp16) output / residual tensors of a persistent-kernel plan so that the swap-AB kernel can use its
// TMA-store epilogue; without it (or with fp32 tensors) the launch falls back to k_conv3x3_v2
int conv_tc_plan_set_io(ConvTcPlan *p, const void *out, int out_half, const void *residual, int res_half, int Cout, int res_up) {
  p->io_half = false;
  DPIR_REQUIRE(!res_up || (residual && res_half && p->g.H % 2 == 0 && p->g.W % 2 == 0), "conv_tc: up-sampled residual needs an fp16 tensor and even H, W");
  if (!p->v2 || !out_half || (residual && !res_half) || Cout % 32 != 0) return DPIR_OK;
  auto encode = get_encode_fn();
  if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled entry point unavailable"); return DPIR_E_CUDA; }
  const ConvTcGeom &g = p->g;
  auto make = [&](CUtensorMap *tm, const void *ptr, int up) -> int {
    const int Ht = up ? g.H / 2 : g.H, Wt = up ? g.W / 2 : g.W;        // the up-sampled residual is read from its half-resolution source
    cuuint64_t gdim[4] = {(cuuint64_t)Cout, (cuuint64_t)Wt, (cuuint64_t)Ht, (cuuint64_t)g.B};
    cuuint64_t gstr[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wt * Cout * 2, (cuuint64_t)Ht * Wt * Cout * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)(up ? 8 : 16), (cuuint32_t)(up ? 1 : 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(io) failed: %d", (int)r); return DPIR_E_CUDA; }
    return DPIR_OK;
  };
  int rc = make(&p->tmOut, out, 0);
  if (rc) return rc;
  p->tmRes = p->tmOut;
  if (residual) { rc = make(&p->tmRes, residual, res_up); if (rc) return rc; }
  p->res_up = res_up != 0;
  p->io_half = true;
  p->has_res = residual != nullptr;
  p->io_out = out;
  p->io_res = residual;
  return DPIR_OK;
}

// attach a fused 1x1 convolution on a second NHWC fp16 operand (same B,H,W; weights [1][CoutPad][Cin2a + Cin2b]); the
// operand may be the virtual concat of two tensors (act2b != null, Cin2b channels) - th.cat of unet.py:660 without a copy
int conv_tc_plan_add_skip(ConvTcPlan *p, const void *act2_half, const void *w2_half, int Cin2a, const void *act2b_half,
                          int Cin2b) {
  DPIR_REQUIRE(p->v2, "conv_tc: skip fusion needs the persistent 3x3 kernel");
  DPIR_REQUIRE(!p->split, "conv_tc: skip fusion is not available for split-precision operands");
  DPIR_REQUIRE(Cin2a % TC_BK == 0 && Cin2b % TC_BK == 0, "conv_tc: fused skip Cin must be a multiple of 64");
  DPIR_REQUIRE(!act2b_half || p->use_v3, "conv_tc: a concatenated skip operand needs the swap-AB kernel");
  auto encode = get_encode_fn();
  if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled entry point unavailable"); return DPIR_E_CUDA; }
  const ConvTcGeom &g = p->g;
  const int Cin2 = Cin2a + (act2b_half ? Cin2b : 0);
  auto make_a = [&](CUtensorMap *tm, const void *ptr, int C) -> int {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)g.W * C * 2, (cuuint64_t)g.H * g.W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, 16, 18, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(A2) failed: %d", (int)r); return DPIR_E_CUDA; }
    return DPIR_OK;
  };
  int rc = make_a(&p->tmA2, act2_half, Cin2a);
  if (rc) return rc;
  p->tmA2b = p->tmA2;
  if (act2b_half) { rc = make_a(&p->tmA2b, act2b_half, Cin2b); if (rc) return rc; }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Cin2, (cuuint64_t)p->CoutPad, 1};
    cuuint64_t gstr[2] = {(cuuint64_t)Cin2 * 2, (cuuint64_t)p->CoutPad * Cin2 * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&p->tmB2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(w2_half), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(B2) failed: %d", (int)r); return DPIR_E_CUDA; }
  }
  p->g2.skip_chunks = Cin2 / TC_BK;
  p->g2.skip_chunks_a = Cin2a / TC_BK;
  return DPIR_OK;
}

// v4: the activation operand is SiLU(GroupNorm(src)) computed inside the kernel from the raw fp16 source(s) and the
// per-(image, channel) coefficient table of launch_gn_coef; needs the swap-AB kernel with its fp16 TMA-store epilogue
int conv_tc_plan_set_gn_fuse(ConvTcPlan *p, const GnFuse &gf, int math) {
  if (p->dxn) {           // narrow-output kernel: the TMA fetches the raw tensor (tmA already describes it), activator warps apply the norm
    DPIR_REQUIRE(!p->split && gf.src1 && gf.coef && gf.C2 == 0 && !gf.up && gf.C1 == p->gd.chunks * TC_BK,
                 "conv_tc: the narrow-output kernel fuses the norm of ONE full-resolution source (C1 %d)", gf.C1);
    DPIR_REQUIRE(((uintptr_t)gf.coef & 15) == 0, "conv_tc: fused operands must be 16-byte aligned");
    p->gf = gf;
    return DPIR_OK;
  }
  p->gf_math = (math == 1 || math == 2) ? math : (math == 0 ? g_gnfuse_math : 3);
  DPIR_REQUIRE(!gf.up || (p->gf_math == 3 && gf.C2 == 0 && p->g.H % 2 == 0 && p->g.W % 2 == 0),
               "conv_tc: the up-sampling index map needs the activate-once producers, one source and even H, W");
  DPIR_REQUIRE(p->v2 && p->use_v3 && !p->split, "conv_tc: GroupNorm fusion needs the persistent swap-AB 3x3 kernel");
  DPIR_REQUIRE(gf.src1 && gf.coef && gf.C1 % TC_BK == 0 && gf.C2 % TC_BK == 0 && gf.C1 + gf.C2 == p->g2.Cin && (gf.C2 == 0 || gf.src2),
               "conv_tc: bad GroupNorm fusion operands (C1 %d, C2 %d, Cin %d)", gf.C1, gf.C2, p->g2.Cin);
  DPIR_REQUIRE((((uintptr_t)gf.src1 | (uintptr_t)gf.src2 | (uintptr_t)gf.coef) & 15) == 0, "conv_tc: fused operands must be 16-byte aligned");
  p->gf = gf;
  // tensor maps of the raw source(s) for the L2 prefetch of the producers' 18 x 18 pixel halo box (the kernel's activation TMA
  // map is unused when the operand is produced in-kernel)
  auto encode = get_encode_fn();
  if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled entry point unavailable"); return DPIR_E_CUDA; }
  const ConvTcGeom &g = p->g;
  auto make = [&](CUtensorMap *tm, const void *ptr, int C) -> int {
    const int Hs = gf.up ? g.H / 2 : g.H, Ws = gf.up ? g.W / 2 : g.W;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)g.B};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ws * C * 2, (cuuint64_t)Hs * Ws * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(gf.up ? 10 : 18), (cuuint32_t)(gf.up ? 10 : 18), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(raw operand) failed: %d", (int)r); return DPIR_E_CUDA; }
    return DPIR_OK;
  };
  int rc = make(&p->tmA, gf.src1, gf.C1);
  if (rc) return rc;
  p->tmAb = p->tmA;
  if (gf.C2) rc = make(&p->tmAb, gf.src2, gf.C2);
  return rc;
}

void conv_tc_plan_destroy(ConvTcPlan *p) { delete p; }

template <int BN, int STAGES>
static int launch_tc_impl(const ConvTcPlan *p, const ConvTcGeom &g, const ConvEpi &epi, cudaStream_t st) {
  constexpr int smem = STAGES * (TC_A_BYTES + BN * TC_BK * 2) + 1024 + 256;
  bool &attr_set = dev_state()->attr_v1[BN == 16 ? 0 : (BN == 64 ? 1 : 2)];
  if (!attr_set) {
    DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv_tc<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(g.tiles_w * g.tiles_h * g.tiles_n, p->CoutPad / BN);
  k_conv_tc<BN, STAGES><<<grid, TC_THREADS, smem, st>>>(p->tmA, p->tmB, g, epi);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

void conv_tc_force_v1(int on) { g_force_v1 = on; }
bool conv_tc_v2_enabled() { return g_force_v1 != 1; }
bool conv_tc_swapab_enabled() { return g_force_v1 == 0; }
void conv_tc_set_gnfuse_math(int m) { g_gnfuse_math = (m == 1 || m == 2) ? m : 3; }   // default of the per-kernel test entry

int launch_conv_tc(const ConvTcPlan *p, const ConvEpi &epi, int Cout, cudaStream_t st) {
  DPIR_REQUIRE(Cout <= p->CoutPad, "conv_tc: Cout exceeds padded weight rows");
  DPIR_REQUIRE(epi.mode != EPI_NHWC || Cout % 4 == 0, "conv_tc: NHWC epilogue needs Cout %% 4 == 0");
  if (p->v2 && epi.mode == EPI_NHWC) {
    DevState *ds = dev_state();
    bool &attr_set = ds->attr_v3;
    if (!attr_set) {
      DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM));
      DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_v3<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM));
      DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_v3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM));
      DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_v3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM));
      DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_v3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM));
      int dev = 0;
      DPIR_CUDA_CHECK(cudaGetDevice(&dev));
      DPIR_CUDA_CHECK(cudaDeviceGetAttribute(&ds->num_sms, cudaDevAttrMultiProcessorCount, dev));
      attr_set = true;
    }
    const int g_num_sms = ds->num_sms;
    ConvV2Geom g2 = p->g2;
    g2.Cout = Cout;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("DPIR_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
    g2.dbg = dbg;
    {
      // ring shapes: the TMA-fed plain kernel keeps round 1's two (halo box + 3 taps) stages; the activate-once producers want a
      // whole chunk (3 dx boxes) of activation slots, per-tap weight slots and an L2 prefetch of the raw operand two chunks ahead.
      // DPIR_CONV_RING="a,w,t,af,wf,tf,pf" overrides (plain: a boxes, w slots of t taps; fused likewise; pf = prefetch distance).
      static int ring[10] = {-1, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (ring[0] < 0) {
        // plain | fused with a residual | prefetch distance | fused without a residual (OUT-only staging -> 2 x 3-tap weight slots)
        int v[10] = {2, 2, 3, V3_A_SLOTS, V3_W_SLOTS, 1, 1, 3, 2, 3};
        const char *e = getenv("DPIR_CONV_RING");
        if (e) {
          int u[10];
          for (int i = 0; i < 10; ++i) u[i] = v[i];
          const int n = sscanf(e, "%d,%d,%d,%d,%d,%d,%d,%d,%d,%d", &u[0], &u[1], &u[2], &u[3], &u[4], &u[5], &u[6], &u[7], &u[8], &u[9]);
          auto ok = [](int a, int w, int t, int stg) {
            return a >= 2 && a <= 4 && w >= 2 && w <= 8 && (t == 1 || t == 3) &&
                   a * V2_A_BYTES + w * t * V3_W_BYTES + 8 * stg <= V3_RING_BYTES + V3_STG_BYTES;
          };
          if (n >= 3 && ok(u[0], u[1], u[2], 6144)) { v[0] = u[0]; v[1] = u[1]; v[2] = u[2]; }
          if (n >= 6 && ok(u[3], u[4], u[5], 6144)) { v[3] = u[3]; v[4] = u[4]; v[5] = u[5]; }
          if (n >= 7 && u[6] >= 0 && u[6] <= 8) v[6] = u[6];
          if (n >= 10 && ok(u[7], u[8], u[9], 2048)) { v[7] = u[7]; v[8] = u[8]; v[9] = u[9]; }
        }
        for (int i = 1; i < 10; ++i) ring[i] = v[i];
        ring[0] = v[0];
      }
      const bool fused = p->gf.coef != nullptr;
      static int plain_nores = -1;            // DPIR_CONV_PLAIN_NORES=0: residual-free plain convs keep the 2-box shape (A/B timing)
      if (plain_nores < 0) { const char *e = getenv("DPIR_CONV_PLAIN_NORES"); plain_nores = (e && e[0] == '0') ? 0 : 1; }
      const bool nores = !epi.residual && (fused || plain_nores);       // OUT-only staging -> room for 3 boxes + 2 three-tap slots
      g2.a_slots = nores ? ring[7] : (fused ? ring[3] : ring[0]);
      g2.w_slots = nores ? ring[8] : (fused ? ring[4] : ring[1]);
      g2.w_taps = nores ? ring[9] : (fused ? ring[5] : ring[2]);
      g2.pf_dist = fused ? ring[6] : 0;
      g2.stg_stride = nores ? 2048 : 6144;
      if (g2.a_slots * V2_A_BYTES + g2.w_slots * g2.w_taps * V3_W_BYTES + 8 * g2.stg_stride > V3_RING_BYTES + V3_STG_BYTES) {
        set_error("conv_tc: ring configuration exceeds the shared-memory budget");
        return DPIR_E_STATE;
      }
    }
    int grid = g2.total_tiles < g_num_sms ? g2.total_tiles : g_num_sms;
    DPIR_REQUIRE((epi.res_up != 0) == p->res_up, "conv_tc: up-sampled residual not planned for this convolution");
    const bool v3_ok = !p->split && p->use_v3 && p->io_half && epi.out_half && epi.out == p->io_out && epi.residual == p->io_res &&
                       (!epi.residual || epi.res_half);
    DPIR_REQUIRE(v3_ok || (!p->gf.coef && g2.skip_chunks_a == g2.skip_chunks),
                 "conv_tc: GroupNorm / concat-skip fusion planned for a conv the swap-AB kernel cannot run");
    if (v3_ok && p->gf.coef) {
      if (p->gf_math == 3)
        k_conv3x3_v3<3><<<grid, V4_THREADS, V3_SMEM, st>>>(p->tmA, p->tmAb, p->tmB, p->tmA2, p->tmA2b, p->tmB2, p->tmOut, p->tmRes, g2, epi, p->gf);
      else if (p->gf_math == 2)
        k_conv3x3_v3<2><<<grid, V4_THREADS, V3_SMEM, st>>>(p->tmA, p->tmAb, p->tmB, p->tmA2, p->tmA2b, p->tmB2, p->tmOut, p->tmRes, g2, epi, p->gf);
      else
        k_conv3x3_v3<1><<<grid, V4_THREADS, V3_SMEM, st>>>(p->tmA, p->tmAb, p->tmB, p->tmA2, p->tmA2b, p->tmB2, p->tmOut, p->tmRes, g2, epi, p->gf);
    } else if (v3_ok) {
      k_conv3x3_v3<0><<<grid, V2_THREADS, V3_SMEM, st>>>(p->tmA, p->tmAb, p->tmB, p->tmA2, p->tmA2b, p->tmB2, p->tmOut, p->tmRes, g2, epi, p->gf);
    } else {
      DPIR_REQUIRE(!p->split || !epi.out_half, "conv_tc: split-precision convolutions write fp32");
      DPIR_REQUIRE(!epi.residual || (epi.res_half != 0) == (epi.out_half != 0),
                   "conv_tc: the v2 persistent kernel needs residual and output of the same precision");
      k_conv3x3_v2<<<grid, V2_THREADS, V2_SMEM, st>>>(p->tmA, p->tmB, p->tmA2, p->tmB2, g2, epi);
    }
    DPIR_LAUNCH_CHECK();
    return DPIR_OK;
  }
  DPIR_REQUIRE(!p->v2, "conv_tc: the persistent 3x3 kernel only has the NHWC epilogue");
  if (p->dxn) {
    DPIR_REQUIRE(!epi.out_half && !epi.res_half && !epi.qstats && !epi.bias2, "conv_tc: the narrow-output kernel stores fp32 and has no statistics / second bias");
    DevState *ds = dev_state();
    const int smem = 2048 + p->gd.chunks * 3 * DXN_W_BYTES + p->gd.a_slots * V2_A_BYTES;
    if (ds->attr_dxn < smem) {
      DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_dxn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_conv3x3_dxn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      ds->attr_dxn = smem;
    }
    if (!ds->num_sms) {
      int dev = 0;
      DPIR_CUDA_CHECK(cudaGetDevice(&dev));
      DPIR_CUDA_CHECK(cudaDeviceGetAttribute(&ds->num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    ConvDxnGeom gd = p->gd;
    gd.Cout = Cout;
    const int grid = gd.total_tiles < ds->num_sms ? gd.total_tiles : ds->num_sms;
    gd.coef = p->gf.coef;
    if (gd.coef) k_conv3x3_dxn<true><<<grid, DXN_THREADS + 32 * DXN_ACT_WARPS, smem, st>>>(p->tmA, p->tmW3, gd, epi);
    else k_conv3x3_dxn<false><<<grid, DXN_THREADS, smem, st>>>(p->tmA, p->tmW3, gd, epi);
    DPIR_LAUNCH_CHECK();
    return DPIR_OK;
  }
  ConvTcGeom g = p->g;
  g.Cout = Cout;
  // (Measured in round 2, no effect: starting every CTA's K loop at a different (tap, chunk) so that the CTAs do not request the same
  // L2 lines in lockstep - 8x8 / 1x1 layers unchanged within noise, profiles/r02_conv_ring_experiments.md section 2f.)
  if (p->BN == 16) return launch_tc_impl<16, 4>(p, g, epi, st);
  // (Small maps - 8x8 / 16x16 - run at ~250 TFLOP/s here.  Deeper rings (9 stages, one CTA per SM) were measured in round 2
  // and change nothing: with 128 x 64 tiles the arithmetic intensity is 43 FLOP per L2 byte and 128 CTAs pull ~6 TB/s from L2,
  // i.e. the level is L2-bandwidth bound; the fix is operand multicast across a cluster)
  if (p->BN == 64) return launch_tc_impl<64, 4>(p, g, epi, st);
  return launch_tc_impl<128, 3>(p, g, epi, st);
}

}  // namespace dpir
