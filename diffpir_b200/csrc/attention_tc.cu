// attention_tc.cu - QKVAttentionLegacy / QKVAttention (guided_diffusion/unet.py:337-354, 371-389) on the 5th-gen tensor cores.
//
//   a[b, t, h*64 + c] = sum_s softmax_s( (q_t * s)(k_s * s) ) v_s[c],   s = 64^-1/4, one (sample, head) at a time, head dim 64
//
// Input: the fp16 NHWC output of the qkv 1x1 convolution, [B][T][3C].  One CTA = 128 queries of one (sample, head); every operand
// tile is fetched by ONE 3-D TMA ({64 channels, 128 tokens, 1 sample}, SWIZZLE_128B: rows beyond T are zero-filled by the
// hardware) straight into the canonical layout; both GEMMs run on tcgen05 with the accumulators in TMEM:
//   S[128 q x 128 keys] = Q K^T      A = Q  [128 x 64]  K-major,   B = K block [128 keys x 64]  K-major          (4 MMAs of K = 16)
//   O[128 q x 64 ch]   += P V        A = P  [128 x 128] K-major (two 64-key swizzle chunks, written by the softmax threads),
//                                    B = V block [128 keys x 64 ch] as it sits in memory = MN-major SWIZZLE_128B
//                                    (instruction-descriptor b_major = 1)                                        (8 MMAs of K = 16)
// The softmax is fp32 with thread = query row = TMEM lane (tcgen05.ld); the reference's scale s on q and on k is applied as
// s^2 = 1/8 inside the exponent (exp2 of a fused multiply-add), which is the same arithmetic up to the rounding of q s, k s.
// Nothing of size T x T is materialised.  The softmax is two-pass over the key blocks (pass A: row maxima, pass B: exp / sum /
// P V): Q K^T is recomputed - attention is 0.1-0.5 % of the UNet FLOPs - and in exchange the O accumulator never leaves TMEM for a
// rescale.  K blocks are double-buffered (the TMA of block kb+1 is issued before block kb's MMA) and V(kb) is fetched while
// S(kb) and its softmax run, so the load latency hides behind the tensor core and the softmax.  Any T works: key columns >= T are masked to -inf, query rows >= T are computed
// on zero operands and not stored (the round-1 kernels required T % 64 == 0, which broke utils_model.test_mode's
// quadrant splits, e.g. 288 x 288 -> T = 324).
#include "unet_kernels.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

namespace dpir {

constexpr int ATC_THREADS = 128;
constexpr int ATC_TILE = 128 * 128;     // bytes of one [128 rows][64 fp16] swizzled tile
// shared memory: Q | K0 | K1 | V | P0 | P1 (16 KB each) + barriers + tmem slot  -> 2 CTAs per SM
constexpr int ATC_SMEM = 6 * ATC_TILE + 1024 + 128;

struct AttnTcPlan {
  CUtensorMap tm;        // qkv fp16 as (3C, T, B), box {64, 128, 1}, SWIZZLE_128B
  const void *qkv;
  int B, T, C;
};

template <typename TO>
__global__ void __launch_bounds__(ATC_THREADS) k_attention_tc(const __grid_constant__ CUtensorMap tmQKV, TO *__restrict__ out, int T, int C,
                                                              int heads, int new_order) {
  // instruction descriptors: D = f32, A = B = f16; S: M 128, N 128, both K-major; PV: M 128, N 64, B MN-major (bit 16)
  constexpr uint32_t IDESC_S = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sQ = smem, *sK = smem + ATC_TILE, *sV = smem + 3 * ATC_TILE, *sP = smem + 4 * ATC_TILE;
  uint64_t *bar_mma = reinterpret_cast<uint64_t *>(smem + 6 * ATC_TILE);     // tcgen05.commit of the MMAs just issued
  uint64_t *bar_q = bar_mma + 1;                                             // Q tile landed
  uint64_t *bar_k = bar_mma + 2;                                             // [2] K block landed, per buffer
  uint64_t *bar_v = bar_mma + 4;                                             // V block landed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  int qoff, koff, voff;
  if (new_order) { qoff = h * 64; koff = C + h * 64; voff = 2 * C + h * 64; }
  else           { qoff = h * 192; koff = qoff + 64; voff = qoff + 128; }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(bar_mma, 1);
    mbar_init(bar_q, 1);
    mbar_init(&bar_k[0], 1);
    mbar_init(&bar_k[1], 1);
    mbar_init(bar_v, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128u;                       // columns [0,128) scores, [128,192) output
  const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);                   // this warp's TMEM lane quarter
  const int row = warp * 32 + lane;                                           // query row inside the tile = TMEM lane
  const int nkb = (T + 127) / 128;
  const int nblk = 2 * nkb;                                                   // pass A blocks, then pass B blocks
  uint32_t ph_mma = 0, ph_v = 0;
  uint32_t ph_k[2] = {0u, 0u};

  // block i of the sequence [pass A: kb = 0..nkb) | pass B: kb = 0..nkb) ] keeps its K tile in buffer i & 1
  auto load_block = [&](int i) {
    const int kb = i < nkb ? i : i - nkb, buf = i & 1;
    mbar_expect_tx(&bar_k[buf], ATC_TILE);
    tma_load_3d(sK + buf * ATC_TILE, &tmQKV, &bar_k[buf], koff, kb * 128, b);
  };
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_q, ATC_TILE);
    tma_load_3d(sQ, &tmQKV, bar_q, qoff, q0, b);
    load_block(0);
  }
  mbar_wait(bar_q, 0);

  const float sl2 = 0.125f * 1.4426950408889634f;                            // s^2 * log2(e)
  float m = -INFINITY, l = 0.f;
  for (int i = 0; i < nblk; ++i) {
    const bool pass_b = i >= nkb;
    const int kb = pass_b ? i - nkb : i, buf = i & 1;
    mbar_wait(&bar_k[buf], ph_k[buf]);
    ph_k[buf] ^= 1u;
    tc_fence_after();
    if (threadIdx.x == 0) {
      // the other K buffer's and the V buffer's last readers (block i-1's MMAs) have completed - their commit was waited for at
      // the end of the previous iteration - so refill them now: the loads overlap this block's MMA and softmax
      if (i + 1 < nblk) load_block(i + 1);
      if (pass_b) {
        mbar_expect_tx(bar_v, ATC_TILE);
        tma_load_3d(sV, &tmQKV, bar_v, voff, kb * 128, b);
      }
      const uint64_t ad = make_sw128_desc(smem_u32(sQ)), bd = make_sw128_desc(smem_u32(sK + buf * ATC_TILE));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16(tS, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), IDESC_S, kk ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1u;
    tc_fence_after();
    const int nvalid = T - kb * 128;                                          // key columns >= nvalid do not exist
    if (!pass_b) {
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t r[32];
        tmem_ld32(tS + lane_addr + (uint32_t)c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < nvalid) m = fmaxf(m, __uint_as_float(r[j]));
      }
      tc_fence_before();
      __syncthreads();                                                        // every thread has read S before the next MMA overwrites it
      continue;
    }
    const float mb = m * sl2;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t r[32];
      tmem_ld32(tS + lane_addr + (uint32_t)c, r);
      tmem_ld_wait();
      uint8_t *prow = sP + (c >> 6) * ATC_TILE + row * 128;                   // 64-key swizzle chunk, this thread's row
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c0 = c + j + 2 * q;
          const float e0 = c0 < nvalid ? exp2f(fmaf(__uint_as_float(r[j + 2 * q]), sl2, -mb)) : 0.f;
          const float e1 = c0 + 1 < nvalid ? exp2f(fmaf(__uint_as_float(r[j + 2 * q + 1]), sl2, -mb)) : 0.f;
          const __half2 hp = __floats2half2_rn(e0, e1);
          const float2 back = __half22float2(hp);          // sum what is actually multiplied into V: rows stay normalised
          l += back.x + back.y;
          w[q] = *reinterpret_cast<const uint32_t *>(&hp);
        }
        const int u = ((c & 63) + j) >> 3;                                     // 16-byte unit inside the 128-byte row
        *reinterpret_cast<uint4 *>(prow + ((u ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {                                                    // O += P V over the 128 keys of this block
      mbar_wait(bar_v, ph_v);
      ph_v ^= 1u;
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t ad = make_sw128_desc(smem_u32(sP + (kk >> 2) * ATC_TILE)) + (uint64_t)(2 * (kk & 3));
        const uint64_t bd = make_sw128_desc(smem_u32(sV + kk * 2048));                       // 16 keys further = 16 rows of 128 B
        umma_f16(tO, ad, bd, IDESC_PV, (kb | kk) ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);                                                // P, this K buffer, V and tS are free again
    ph_mma ^= 1u;
    tc_fence_after();
  }

  // ---------------------------------------------------------------- epilogue: O / l -> out[b, q0 + row, h*64 ..]
  {
    const float inv = 1.f / l;
    const bool ok = q0 + row < T;
    TO *dst = out + ((int64_t)b * T + q0 + row) * C + h * 64;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t r[32];
      tmem_ld32(tO + lane_addr + (uint32_t)c, r);
      tmem_ld_wait();
      if (ok) {
        if constexpr (sizeof(TO) == 2) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const __half2 hv = __floats2half2_rn(__uint_as_float(r[j + 2 * q]) * inv, __uint_as_float(r[j + 2 * q + 1]) * inv);
              w[q] = *reinterpret_cast<const uint32_t *>(&hv);
            }
            *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(dst) + c + j) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dst) + c + j) =
                make_float4(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv, __uint_as_float(r[j + 2]) * inv,
                            __uint_as_float(r[j + 3]) * inv);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

static PFN_cuTensorMapEncodeTiled_v12000 attn_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// qkv_half: fp16 [B][T][3C] (the NHWC output of the qkv convolution); the plan only holds the tensor map over it
int attn_tc_plan_create(const void *qkv_half, int B, int T, int C, AttnTcPlan **out) {
  DPIR_REQUIRE(qkv_half && B >= 1 && T >= 1 && C >= 64 && C % 64 == 0, "attention: bad shape");
  DPIR_REQUIRE(((uintptr_t)qkv_half & 15) == 0, "attention: qkv must be 16-byte aligned");
  auto encode = attn_encode_fn();
  if (!encode) { set_error("attention: cuTensorMapEncodeTiled entry point unavailable"); return DPIR_E_CUDA; }
  AttnTcPlan *p = new AttnTcPlan();
  p->qkv = qkv_half; p->B = B; p->T = T; p->C = C;
  cuuint64_t gdim[3] = {(cuuint64_t)(3 * C), (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)(3 * C) * 2, (cuuint64_t)T * 3 * C * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(&p->tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(qkv_half), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attention: cuTensorMapEncodeTiled failed: %d", (int)r); delete p; return DPIR_E_CUDA; }
  *out = p;
  return DPIR_OK;
}

void attn_tc_plan_destroy(AttnTcPlan *p) { delete p; }

int launch_attention_tc(const AttnTcPlan *p, void *out, int out_half, int heads, int new_order, cudaStream_t st) {
  DPIR_REQUIRE(p && out && p->C == heads * 64, "attention: head dim must be 64");
  static bool attr_set[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_attention_tc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_attention_tc<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    attr_set[dev] = true;
  }
  dim3 grid((p->T + 127) / 128, heads, p->B);
  if (out_half) k_attention_tc<__half><<<grid, ATC_THREADS, ATC_SMEM, st>>>(p->tm, (__half *)out, p->T, p->C, heads, new_order);
  else k_attention_tc<float><<<grid, ATC_THREADS, ATC_SMEM, st>>>(p->tm, (float *)out, p->T, p->C, heads, new_order);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

}  // namespace dpir
