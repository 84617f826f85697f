// attention_tc.cu - QKVAttentionLegacy / QKVAttention (guided_diffusion/unet.py:337-354, 371-389) on the 5th-gen tensor cores.
//
//   a[b, t, h*64 + c] = sum_s softmax_s( (q_t * s)(k_s * s) ) v_s[c],   s = 64^-1/4, one (sample, head) at a time, head dim 64
//
// One CTA = 128 queries of one (sample, head); both GEMMs run on tcgen05 with the accumulators in TMEM:
//   S[128 q x 128 keys] = Q K^T      A = Q  [128 x 64]  K-major SWIZZLE_128B,   B = K block [128 keys x 64]  K-major    (4 MMAs of K = 16)
//   O[128 q x 64 ch]   += P V        A = P  [128 x 128] K-major (two 64-key swizzle chunks),  B = V block [128 keys x 64 ch] as stored
//                                    = MN-major SWIZZLE_128B (channels contiguous), instruction-descriptor b_major = 1   (8 MMAs of K = 16)
// q, k (scaled) and v are rounded to fp16 while they are staged into shared memory in the canonical swizzled layout (the same
// operand rounding as every other GEMM of the fp16 pipelines, and what the reference's own use_fp16 mode does); the softmax is
// fp32: thread = query row = TMEM lane, scores read back with tcgen05.ld.  Nothing of size T x T is materialised.  The softmax is
// two-pass over the key blocks (pass A: row maxima, pass B: exp / sum / P V) - Q K^T is recomputed (attention is 0.1 % of the UNet
// FLOPs) and in exchange the O accumulator never leaves TMEM for a rescale.  Any T works: key columns >= T are masked to -inf,
// query rows >= T are computed on zero-filled operands and not stored (the round-1 kernels required T % 64 == 0,
// which broke utils_model.test_mode's quadrant splits, e.g. 288 x 288 -> T = 324).
#include "unet_kernels.cuh"
#include "tc_ptx.cuh"

namespace dpir {

constexpr int ATC_THREADS = 128;
constexpr int ATC_TILE = 128 * 128;     // bytes of one [128 rows][64 fp16] swizzled tile
// shared memory: Q | K | V | P0 | P1 (16 KB each) + barrier + tmem slot
constexpr int ATC_SMEM = 5 * ATC_TILE + 1024 + 64;

// stage rows [r0, r0+128) of a [T][3C] fp32 matrix (64 columns at `col`) into a swizzled [128][64] fp16 tile, scaled; rows >= T
// are zero.  8 threads per row (one 16-byte unit each), 16 rows per pass: global reads are 256 B contiguous per row.
__device__ __forceinline__ void stage_tile(uint8_t *dst, const float *__restrict__ base, int C3, int col, int r0, int T, float scale) {
  const int u = threadIdx.x & 7, rr = threadIdx.x >> 3;
#pragma unroll 2
  for (int p = 0; p < 8; ++p) {
    const int r = p * 16 + rr;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < T) {
      const float4 *src = reinterpret_cast<const float4 *>(base + (int64_t)(r0 + r) * C3 + col + u * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      const __half2 h0 = __floats2half2_rn(a.x * scale, a.y * scale), h1 = __floats2half2_rn(a.z * scale, a.w * scale);
      const __half2 h2 = __floats2half2_rn(b.x * scale, b.y * scale), h3 = __floats2half2_rn(b.z * scale, b.w * scale);
      o.x = *reinterpret_cast<const uint32_t *>(&h0); o.y = *reinterpret_cast<const uint32_t *>(&h1);
      o.z = *reinterpret_cast<const uint32_t *>(&h2); o.w = *reinterpret_cast<const uint32_t *>(&h3);
    }
    *reinterpret_cast<uint4 *>(dst + r * 128 + ((u ^ (r & 7)) << 4)) = o;
  }
}

template <typename TO>
__global__ void __launch_bounds__(ATC_THREADS) k_attention_tc(const float *__restrict__ qkv, TO *__restrict__ out, int T, int C, int heads,
                                                              int new_order) {
  // instruction descriptors: D = f32, A = B = f16; S: M 128, N 128, both K-major; PV: M 128, N 64, B MN-major (bit 16)
  constexpr uint32_t IDESC_S = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *sQ = smem, *sK = smem + ATC_TILE, *sV = smem + 2 * ATC_TILE, *sP = smem + 3 * ATC_TILE;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 5 * ATC_TILE);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int C3 = 3 * C;
  int qoff, koff, voff;
  if (new_order) { qoff = h * 64; koff = C + h * 64; voff = 2 * C + h * 64; }
  else           { qoff = h * 192; koff = qoff + 64; voff = qoff + 128; }
  const float *base = qkv + (int64_t)b * T * C3;
  const float sc = 0.35355339059327373f;     // 64^-1/4 on q and on k, like the reference (unet.py:345-347)

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  stage_tile(sQ, base, C3, qoff, q0, T, sc);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + 128u;                       // columns [0,128) scores, [128,192) output
  const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);                   // this warp's TMEM lane quarter
  const int row = warp * 32 + lane;                                           // query row inside the tile = TMEM lane
  const int nkb = (T + 127) / 128;
  uint32_t phase = 0;

  auto issue_s = [&]() {                                                      // S = Q K^T  (one elected thread)
    const uint64_t ad = make_sw128_desc(smem_u32(sQ)), bd = make_sw128_desc(smem_u32(sK));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_f16(tS, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), IDESC_S, kk ? 1u : 0u);
    umma_commit(bar);
  };
  auto wait_mma = [&]() {
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
  };

  // ---------------------------------------------------------------- pass A: row maxima over all keys
  float m = -INFINITY;
  for (int kb = 0; kb < nkb; ++kb) {
    stage_tile(sK, base, C3, koff, kb * 128, T, sc);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) issue_s();
    wait_mma();
    const int nvalid = T - kb * 128;                                          // key columns >= nvalid do not exist
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
    __syncthreads();                                                          // every thread has read S before sK / tS are reused
  }

  // ---------------------------------------------------------------- pass B: p = exp(s - m), l = sum p, O += P V
  float l = 0.f;
  for (int kb = 0; kb < nkb; ++kb) {
    stage_tile(sK, base, C3, koff, kb * 128, T, sc);
    stage_tile(sV, base, C3, voff, kb * 128, T, 1.0f);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) issue_s();
    wait_mma();
    const int nvalid = T - kb * 128;
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
          const float e0 = c0 < nvalid ? __expf(__uint_as_float(r[j + 2 * q]) - m) : 0.f;
          const float e1 = c0 + 1 < nvalid ? __expf(__uint_as_float(r[j + 2 * q + 1]) - m) : 0.f;
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
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t ad = make_sw128_desc(smem_u32(sP + (kk >> 2) * ATC_TILE)) + (uint64_t)(2 * (kk & 3));
        const uint64_t bd = make_sw128_desc(smem_u32(sV + kk * 2048));          // 16 keys further = 16 rows of 128 B
        umma_f16(tO, ad, bd, IDESC_PV, (kb | kk) ? 1u : 0u);
      }
      umma_commit(bar);
    }
    wait_mma();                                                                // P / V / K tiles and tS are free again
    tc_fence_before();
    __syncthreads();
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

int launch_attention_tc(const float *qkv, void *out, int out_half, int B, int T, int C, int heads, int new_order, cudaStream_t st) {
  DPIR_REQUIRE(C == heads * 64, "attention: head dim must be 64 (C=%d heads=%d)", C, heads);
  DPIR_REQUIRE(T >= 1 && C % 8 == 0, "attention: bad shape");
  static bool attr_set[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_set[dev]) {
    DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_attention_tc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    DPIR_CUDA_CHECK(cudaFuncSetAttribute(k_attention_tc<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    attr_set[dev] = true;
  }
  dim3 grid((T + 127) / 128, heads, B);
  if (out_half) k_attention_tc<__half><<<grid, ATC_THREADS, ATC_SMEM, st>>>(qkv, (__half *)out, T, C, heads, new_order);
  else k_attention_tc<float><<<grid, ATC_THREADS, ATC_SMEM, st>>>(qkv, (float *)out, T, C, heads, new_order);
  DPIR_LAUNCH_CHECK();
  return DPIR_OK;
}

}  // namespace dpir
