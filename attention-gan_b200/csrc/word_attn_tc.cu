// Generator word-context attention on tcgen05 for 16-bit I/O (bf16 / fp16 feature maps).
//
// Replaces AttentionModule.forward (reference networks/attention.py:25-79).  At C = 32, T = 18 the
// CUDA-core formulation needs 1152 FMA per pixel forward against 164 B of HBM traffic, which puts
// its ceiling (fp32 FMA rate) below the HBM roofline; the two tiny contractions therefore run on
// the tensor cores and the SM only does the softmax and the stores:
//
//   tile       = 128 consecutive pixels of one sample (NCHW, HW contiguous)
//   TMA        h tile [C rows x 128 px] -> smem, 128B swizzle; used as the MN-major A operand
//   GEMM1      S[px, t]   = sum_c h[c, px] * (W.e)[c, t] * scale*log2(e)      (B operand: hi + lo split,
//                            so W.e enters with fp32 accuracy)                  M=128, N=NT, K=C
//   epilogue 1 thread = pixel: mask, softmax over the T words in registers (exp2), attn -> global,
//              P = attn as fp16 written back into TMEM (tcgen05.st) over the scores
//   GEMM2      ctx[px, c] = sum_t P[px, t] * (W.e)[c, t]        A operand from TMEM    M=128, N=C, K=NT
//   epilogue 2 thread = pixel: ctx -> global (lanes = consecutive pixels: coalesced)
//
// TMEM: 2 tile buffers x (NT score/P columns + C context columns) = 128 columns per CTA at
// T <= 32, C <= 32 (256 otherwise), so up to 4 CTAs share an SM and each keeps two tiles in flight.  Warps: 0 = TMA, 1 = MMA, 2-5 = epilogue.
#include <algorithm>
#include <type_traits>

#include "tc_common.cuh"

namespace agb {
namespace tc {

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc_sw128_mn_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <typename T16> __device__ __forceinline__ T16 f2h(float v);
template <> __device__ __forceinline__ __half f2h<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 f2h<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct AttnFwdParams {
  const float* we;        // [B, C, T] fp32 projected words
  const int64_t* mask;    // [B, T]
  void* ctx;              // [B, C, HW] io dtype, batch stride ctx_bs
  void* attn;             // [B, T, HW] io dtype or null
  int64_t ctx_bs;
  int B, C, HW, T;
  float qscale;           // scale * log2(e)
  int tiles, ctas_per_sample;
};

constexpr int kAttnStages = 3;

// load TL (multiple of 8, <= 64) consecutive TMEM columns of this thread's lane
template <int TL>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
  if constexpr (TL >= 32) {
    tmem_ld32(taddr, v);
    if constexpr (TL > 32) tmem_ld_cols<TL - 32>(taddr + 32, v + 32);
  } else if constexpr (TL >= 16) {
    tmem_ld16(taddr, v);
    if constexpr (TL > 16) tmem_ld_cols<TL - 16>(taddr + 16, v + 16);
  } else {
    tmem_ld8(taddr, v);
  }
}

// TL = number of word columns the epilogue touches (T rounded up to 8); NT = MMA N (32 or 64)
template <typename IO, int NT, int TL>
__global__ void __launch_bounds__(192)
word_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapH, const AttnFwdParams p) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  const int C = p.C;
  const int box_bytes = C * 128;                 // one [C x 64 px] box
  const int stage_bytes = 2 * box_bytes;
  unsigned char* sH = smem;
  unsigned char* sB1hi = smem + kAttnStages * stage_bytes;   // [NT rows t][64 k=c]  io type
  unsigned char* sB1lo = sB1hi + NT * 128;
  unsigned char* sB2hi = sB1lo + NT * 128;                   // [C rows c][64 k=t]   fp16
  unsigned char* sB2lo = sB2hi + 64 * 128;
  __shared__ uint64_t h_full[kAttnStages], h_empty[kAttnStages], s_full[2], p_ready[2], c_full[2], c_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int bufc = NT + C;                       // TMEM columns of one tile buffer: scores/P + context
  const uint32_t tmem_cols = 2 * bufc <= 128 ? 128u : 256u;
  const int ntile = (p.tiles - (int)blockIdx.x + p.ctas_per_sample - 1) / p.ctas_per_sample;   // tiles of this CTA

  if (threadIdx.x == 0) {
    for (int i = 0; i < kAttnStages; ++i) {
      mbar_init(&h_full[i], 1);
      mbar_init(&h_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&c_full[i], 1);
      mbar_init(&c_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  // B operands of both GEMMs from W.e of this sample: hi + lo 16-bit split, K-major swizzled rows
  {
    const float* we = p.we + (size_t)b * C * p.T;
    for (int i = threadIdx.x; i < NT * 64; i += blockDim.x) {       // B1[t][c]
      const int t = i >> 6, c = i & 63;
      const float x = (t < p.T && c < C) ? we[c * p.T + t] * p.qscale : 0.f;
      const IO hi = f2h<IO>(x);
      const IO lo = f2h<IO>(x - to_f32(hi));
      *reinterpret_cast<IO*>(sB1hi + sw128_off(t, c)) = hi;
      *reinterpret_cast<IO*>(sB1lo + sw128_off(t, c)) = lo;
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {       // B2[c][t]
      const int c = i >> 6, t = i & 63;
      const float x = (t < p.T && c < C) ? we[c * p.T + t] : 0.f;
      const __half hi = __float2half_rn(x);
      const __half lo = __float2half_rn(x - __half2float(hi));
      *reinterpret_cast<__half*>(sB2hi + sw128_off(c, t)) = hi;
      *reinterpret_cast<__half*>(sB2lo + sw128_off(c, t)) = lo;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  constexpr int fmt_io = std::is_same<IO, __nv_bfloat16>::value ? 1 : 0;

  if (warp == 0) {
    // ===================== TMA producer: h tiles =====================
    if (elect_one()) {
      for (int it = 0; it < ntile; ++it) {
        const int s = it % kAttnStages, use = it / kAttnStages;
        const int tile = blockIdx.x + it * p.ctas_per_sample;
        mbar_wait(&h_empty[s], (use & 1) ^ 1);
        mbar_expect_tx(&h_full[s], (uint32_t)stage_bytes);
        tma_load_2d(sH + s * stage_bytes, &mapH, &h_full[s], tile * 128, b * C);
        tma_load_2d(sH + s * stage_bytes + box_bytes, &mapH, &h_full[s], tile * 128 + 64, b * C);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc1 = make_idesc(128, NT, fmt_io) | (1u << 15);   // A MN-major (pixels contiguous)
      const uint32_t idesc2 = make_idesc(128, C, 0);                       // P (TMEM) x W.e, fp16
      const int ks1 = C >> 4, ks2 = NT >> 4;
      auto gemm1 = [&](int it) {
        const int s = it % kAttnStages, use = it / kAttnStages, u = it & 1;
        mbar_wait(&h_full[s], use & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sH + s * stage_bytes);
        for (int kk = 0; kk < ks1; ++kk) {
          const uint64_t da = make_desc_sw128_mn_lbo(a0 + kk * 2048, (uint32_t)box_bytes);
          umma_f16(tmem + u * bufc, da, make_desc_sw128(smem_u32(sB1hi)) + 2 * kk, idesc1, kk ? 1u : 0u);
          umma_f16(tmem + u * bufc, da, make_desc_sw128(smem_u32(sB1lo)) + 2 * kk, idesc1, 1u);
        }
        umma_commit(&h_empty[s]);
        umma_commit(&s_full[u]);
      };
      if (ntile > 0) gemm1(0);
      if (ntile > 1) gemm1(1);
      for (int it = 0; it < ntile; ++it) {
        const int u = it & 1, k = it >> 1;
        mbar_wait(&p_ready[u], k & 1);
        mbar_wait(&c_empty[u], (k & 1) ^ 1);
        tc_fence_after();
        for (int kk = 0; kk < ks2; ++kk) {
          umma_f16_ts(tmem + u * bufc + NT, tmem + u * bufc + kk * 8, make_desc_sw128(smem_u32(sB2hi)) + 2 * kk, idesc2,
                      kk ? 1u : 0u);
          umma_f16_ts(tmem + u * bufc + NT, tmem + u * bufc + kk * 8, make_desc_sw128(smem_u32(sB2lo)) + 2 * kk, idesc2, 1u);
        }
        umma_commit(&c_full[u]);
        if (it + 2 < ntile) gemm1(it + 2);
      }
    }
  } else {
    // ===================== epilogue: thread = pixel =====================
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int px = q * 32 + lane;
    uint64_t valid = 0;
    for (int t = 0; t < p.T; ++t)
      if (p.mask[(size_t)b * p.T + t] != 0) valid |= 1ull << t;
    IO* ctx = (IO*)p.ctx + (size_t)b * p.ctx_bs;
    IO* attn = p.attn ? (IO*)p.attn + (size_t)b * p.T * p.HW : nullptr;

    auto softmax_phase = [&](int it) {
      const int u = it & 1, k = it >> 1;
      const int pix = (blockIdx.x + it * p.ctas_per_sample) * 128 + px;
      mbar_wait(&s_full[u], k & 1);
      tc_fence_after();
      float s[TL];
      tmem_ld_cols<TL>(tmem + lane_addr + u * bufc, s);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int t = 0; t < TL; ++t) {
        if (!((valid >> t) & 1)) s[t] = -INFINITY;
        m4[t & 3] = fmaxf(m4[t & 3], s[t]);
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < TL; ++t) {
        s[t] = exp2f(s[t] - mx);           // all-masked sample: (-inf) - (-inf) = NaN, like the reference
        s4[t & 3] += s[t];
      }
      const float inv = 1.f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
      uint32_t pk[NT / 2];
#pragma unroll
      for (int t = 0; t < NT; t += 2) {
        if (t < TL) {
          const float a0 = s[t] * inv, a1 = s[t + 1] * inv;
          s[t] = a0;
          s[t + 1] = a1;
          const __half2 h2 = __floats2half2_rn(a0, a1);
          pk[t / 2] = *reinterpret_cast<const uint32_t*>(&h2);
        } else {
          pk[t / 2] = 0u;
        }
      }
      tmem_st16(tmem + lane_addr + u * bufc, pk);
      if constexpr (NT == 64) tmem_st16(tmem + lane_addr + u * bufc + 16, pk + 16);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[u]);
      if (attn != nullptr && pix < p.HW) {
#pragma unroll
        for (int t = 0; t < TL; ++t)
          if (t < p.T) attn[(size_t)t * p.HW + pix] = f2h<IO>(s[t]);
      }
    };
    auto context_phase = [&](int it) {
      const int u = it & 1, k = it >> 1;
      const int pix = (blockIdx.x + it * p.ctas_per_sample) * 128 + px;
      mbar_wait(&c_full[u], k & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < C; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + lane_addr + u * bufc + NT + c0, v);
        tmem_ld_wait();
        if (pix < p.HW) {
#pragma unroll
          for (int j = 0; j < 16; ++j) ctx[(size_t)(c0 + j) * p.HW + pix] = f2h<IO>(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&c_empty[u]);
    };
    for (int it = 0; it < ntile; ++it) {
      softmax_phase(it);
      if (it > 0) context_phase(it - 1);
    }
    if (ntile > 0) context_phase(ntile - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <typename IO>
static int launch_attn_fwd_tc(const void* images, const AttnFwdParams& p, cudaStream_t st) {
  CUtensorMap mapH;
  if (int rc = make_tmap_2d(&mapH, images, (uint64_t)p.B * p.C, (uint64_t)p.HW, (uint32_t)p.C,
                            std::is_same<IO, __nv_bfloat16>::value))
    return rc;
  const int NT = p.T <= 32 ? 32 : 64;
  const int smem = kAttnStages * 2 * p.C * 128 + 2 * NT * 128 + 2 * 64 * 128 + 1024;
  dim3 grid(p.ctas_per_sample, p.B);
  const int slot = prof_begin(PROF_ATTN_FWD, st);
#define AGB_ATTN_FWD_CASE(NTV, TLV)                                                              \
  {                                                                                               \
    auto kern = word_attn_fwd_tc_kernel<IO, NTV, TLV>;                                            \
    AGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));      \
    kern<<<grid, 192, smem, st>>>(mapH, p);                                                       \
  }
  switch ((p.T + 7) / 8) {
    case 1: AGB_ATTN_FWD_CASE(32, 8) break;
    case 2: AGB_ATTN_FWD_CASE(32, 16) break;
    case 3: AGB_ATTN_FWD_CASE(32, 24) break;
    case 4: AGB_ATTN_FWD_CASE(32, 32) break;
    case 5: AGB_ATTN_FWD_CASE(64, 40) break;
    case 6: AGB_ATTN_FWD_CASE(64, 48) break;
    case 7: AGB_ATTN_FWD_CASE(64, 56) break;
    default: AGB_ATTN_FWD_CASE(64, 64) break;
  }
#undef AGB_ATTN_FWD_CASE
  prof_end(slot, st);
  return check_launch("word_attn_fwd_tc_kernel");
}

// =============================================================================================
// backward                                                   (autograd of a3)
//
//   per tile of 128 pixels (thread = pixel in the epilogues):
//   GEMM1a  S[px,t] = h^T (W.e) scale log2e          GEMM1b  G[px,t] = dctx^T (W.e)
//   epilogue 1: a = softmax_t(S), g = G (+ dattn), ds = a (g - sum_t a g);
//               ds -> TMEM as bf16 hi + lo (A operand of GEMM3); [a | ds] -> smem, transposed
//               (B operand of GEMM4)
//   GEMM3   dh[px,c] = sum_t ds[px,t] (W.e)[c,t] scale                        (A from TMEM)
//   GEMM4   acc[(dctx rows; h rows), (a cols | ds cols)] += [dctx; h][., px] [a | ds][px, .]
//           accumulated in TMEM over all tiles of the CTA; d(W.e) = acc[dctx, a] + scale acc[h, ds]
//   epilogue 2: dh -> global
// TMEM (256 columns): two tile buffers of 96 (S|P, G, dh) + 64 accumulator columns.
// =============================================================================================
struct AttnBwdParams {
  const float* we;
  const int64_t* mask;
  const void* dattn;      // [B, T, HW] io dtype or null
  void* dh;               // [B, C, HW]
  float* part;            // [B, ctas_per_sample, C, T] per-CTA partial d(W.e)
  int B, C, HW, T;
  float scale;
  int tiles, ctas_per_sample;
};

constexpr int kBwdNT = 32;

constexpr int kBwdStages = 2;

template <typename IO, int TL>
__global__ void __launch_bounds__(192)
word_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapD,
                        const AttnBwdParams p) {
  constexpr int NT = kBwdNT;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  const int C = p.C;
  const int box = C * 128;                       // one [C x 64 px] box
  const int stage_bytes = 4 * box;               // dctx_lo | h_lo | dctx_hi | h_hi
  unsigned char* sIn = smem;
  unsigned char* sB1s = smem + kBwdStages * stage_bytes;       // [t][c] scaled*log2e   hi, lo (io type)
  unsigned char* sB1u = sB1s + 2 * NT * 128;                   // [t][c] unscaled       hi, lo (io type)
  unsigned char* sB2 = sB1u + 2 * NT * 128;                    // [c][t] * scale        hi, lo (bf16), 32 rows each
  unsigned char* sBt = sB2 + 2 * 32 * 128;                     // 2 buffers x 2 px-chunks x [2NT rows][64 px]
  float* sAcc = reinterpret_cast<float*>(sBt + 2 * 2 * (2 * NT) * 128);   // [2][C][NT]
  __shared__ uint64_t in_full[kBwdStages], in_empty[kBwdStages], s_full[2], p_ready[2], dh_full[2], dh_empty[2], acc_done;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int ntile = (p.tiles - (int)blockIdx.x + p.ctas_per_sample - 1) / p.ctas_per_sample;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdStages; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&in_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&dh_full[i], 1);
      mbar_init(&dh_empty[i], 4);
    }
    mbar_init(&acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  {
    const float* we = p.we + (size_t)b * C * p.T;
    const float qs = p.scale * kLog2e;
    for (int i = threadIdx.x; i < NT * 64; i += blockDim.x) {
      const int t = i >> 6, c = i & 63;
      const float w = (t < p.T && c < C) ? we[c * p.T + t] : 0.f;
      const float xs = w * qs;
      IO hi = f2h<IO>(xs);
      *reinterpret_cast<IO*>(sB1s + sw128_off(t, c)) = hi;
      *reinterpret_cast<IO*>(sB1s + NT * 128 + sw128_off(t, c)) = f2h<IO>(xs - to_f32(hi));
      hi = f2h<IO>(w);
      *reinterpret_cast<IO*>(sB1u + sw128_off(t, c)) = hi;
      *reinterpret_cast<IO*>(sB1u + NT * 128 + sw128_off(t, c)) = f2h<IO>(w - to_f32(hi));
    }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
      const int c = i >> 6, t = i & 63;
      const float x = (t < p.T && c < C) ? we[c * p.T + t] * p.scale : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(x);
      *reinterpret_cast<__nv_bfloat16*>(sB2 + sw128_off(c, t)) = hi;
      *reinterpret_cast<__nv_bfloat16*>(sB2 + 32 * 128 + sw128_off(c, t)) = __float2bfloat16_rn(x - __bfloat162float(hi));
    }
    // rows of the transposed [a | ds] operand that no thread writes (t >= T) must be zero
    for (int i = threadIdx.x; i < 2 * 2 * (2 * NT) * 128 / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(sBt)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  constexpr int fmt_io = std::is_same<IO, __nv_bfloat16>::value ? 1 : 0;
  constexpr uint32_t kAccCol = 192;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < ntile; ++it) {
        const int s = it % kBwdStages, use = it / kBwdStages;
        const int px0 = (blockIdx.x + it * p.ctas_per_sample) * 128;
        mbar_wait(&in_empty[s], (use & 1) ^ 1);
        mbar_expect_tx(&in_full[s], (uint32_t)stage_bytes);
        unsigned char* st = sIn + s * stage_bytes;
        tma_load_2d(st, &mapD, &in_full[s], px0, b * C);
        tma_load_2d(st + box, &mapH, &in_full[s], px0, b * C);
        tma_load_2d(st + 2 * box, &mapD, &in_full[s], px0 + 64, b * C);
        tma_load_2d(st + 3 * box, &mapH, &in_full[s], px0 + 64, b * C);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc1 = make_idesc(128, NT, fmt_io) | (1u << 15);
      const uint32_t idesc3 = make_idesc(128, C, 1);               // ds (bf16, TMEM) x W.e (bf16)
      const uint32_t idesc4 = make_idesc(128, 2 * NT, fmt_io);     // [dctx; h] x [a | ds], both K-major over pixels
      const int ks1 = C >> 4;
      auto gemm1 = [&](int it) {
        const int s = it % kBwdStages, use = it / kBwdStages, u = it & 1;
        mbar_wait(&in_full[s], use & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(sIn + s * stage_bytes);
        for (int kk = 0; kk < ks1; ++kk) {
          const uint64_t dh_ = make_desc_sw128_mn_lbo(st + box + kk * 2048, (uint32_t)(2 * box));    // h
          const uint64_t dd_ = make_desc_sw128_mn_lbo(st + kk * 2048, (uint32_t)(2 * box));          // dctx
          umma_f16(tmem + u * 96, dh_, make_desc_sw128(smem_u32(sB1s)) + 2 * kk, idesc1, kk ? 1u : 0u);
          umma_f16(tmem + u * 96, dh_, make_desc_sw128(smem_u32(sB1s + NT * 128)) + 2 * kk, idesc1, 1u);
          umma_f16(tmem + u * 96 + NT, dd_, make_desc_sw128(smem_u32(sB1u)) + 2 * kk, idesc1, kk ? 1u : 0u);
          umma_f16(tmem + u * 96 + NT, dd_, make_desc_sw128(smem_u32(sB1u + NT * 128)) + 2 * kk, idesc1, 1u);
        }
        umma_commit(&s_full[u]);
      };
      if (ntile > 0) gemm1(0);
      if (ntile > 1) gemm1(1);
      for (int it = 0; it < ntile; ++it) {
        const int u = it & 1, k = it >> 1, s = it % kBwdStages;
        mbar_wait(&p_ready[u], k & 1);
        mbar_wait(&dh_empty[u], (k & 1) ^ 1);
        tc_fence_after();
        for (int kk = 0; kk < NT / 16; ++kk) {      // dh = ds (hi + lo) x W.e (hi + lo), lo*lo dropped
          const uint32_t a_hi = tmem + u * 96 + kk * 8, a_lo = tmem + u * 96 + 16 + kk * 8;
          const uint64_t b_hi = make_desc_sw128(smem_u32(sB2)) + 2 * kk, b_lo = make_desc_sw128(smem_u32(sB2 + 32 * 128)) + 2 * kk;
          umma_f16_ts(tmem + u * 96 + 2 * NT, a_hi, b_hi, idesc3, kk ? 1u : 0u);
          umma_f16_ts(tmem + u * 96 + 2 * NT, a_lo, b_hi, idesc3, 1u);
          umma_f16_ts(tmem + u * 96 + 2 * NT, a_hi, b_lo, idesc3, 1u);
        }
        umma_commit(&dh_full[u]);
        const uint32_t st = smem_u32(sIn + s * stage_bytes);
        const uint32_t bt = smem_u32(sBt + u * (2 * (2 * NT) * 128));
        for (int j = 0; j < 2; ++j)                  // two chunks of 64 pixels
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem + kAccCol, make_desc_sw128(st + j * 2 * box) + 2 * kk,
                     make_desc_sw128(bt + j * (2 * NT) * 128) + 2 * kk, idesc4, (it | j | kk) ? 1u : 0u);
        umma_commit(&in_empty[s]);
        if (it + 2 < ntile) gemm1(it + 2);
      }
      umma_commit(&acc_done);
    }
  } else {
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int px = q * 32 + lane;
    uint64_t valid = 0;
    for (int t = 0; t < p.T; ++t)
      if (p.mask[(size_t)b * p.T + t] != 0) valid |= 1ull << t;
    IO* dh = (IO*)p.dh + (size_t)b * C * p.HW;
    const IO* dattn = p.dattn ? (const IO*)p.dattn + (size_t)b * p.T * p.HW : nullptr;

    auto softmax_phase = [&](int it) {
      const int u = it & 1, k = it >> 1;
      const int pix = (blockIdx.x + it * p.ctas_per_sample) * 128 + px;
      mbar_wait(&s_full[u], k & 1);
      tc_fence_after();
      float s[TL], g[TL];
      tmem_ld_cols<TL>(tmem + lane_addr + u * 96, s);
      tmem_ld_cols<TL>(tmem + lane_addr + u * 96 + NT, g);
      tmem_ld_wait();
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int t = 0; t < TL; ++t) {
        if (!((valid >> t) & 1)) s[t] = -INFINITY;
        m4[t & 3] = fmaxf(m4[t & 3], s[t]);
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < TL; ++t) {
        s[t] = exp2f(s[t] - mx);
        s4[t & 3] += s[t];
      }
      const float inv = 1.f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < TL; ++t) {
        s[t] *= inv;
        if (dattn != nullptr && t < p.T && pix < p.HW) g[t] += to_f32(dattn[(size_t)t * p.HW + pix]);
        if (t >= p.T) g[t] = 0.f;
        d4[t & 3] = fmaf(s[t], g[t], d4[t & 3]);
      }
      const float dot = (d4[0] + d4[1]) + (d4[2] + d4[3]);
      uint32_t pk[NT];                               // [0,16): ds hi pairs, [16,32): ds lo pairs (bf16)
      unsigned char* btu = sBt + u * (2 * (2 * NT) * 128) + (px >> 6) * ((2 * NT) * 128);
      const int pc = px & 63;
      const bool live = pix < p.HW;                  // pixels past the end of the map contribute nothing
#pragma unroll
      for (int t = 0; t < NT; t += 2) {
        if (t < TL) {
          float d0 = s[t] * (g[t] - dot), d1 = s[t + 1] * (g[t + 1] - dot);
          if (!live) { d0 = 0.f; d1 = 0.f; }
          const __nv_bfloat16 h0 = __float2bfloat16_rn(d0), h1 = __float2bfloat16_rn(d1);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(d0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(d1 - __bfloat162float(h1));
          pk[t / 2] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          pk[16 + t / 2] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          if (t < p.T) {
            *reinterpret_cast<IO*>(btu + sw128_off(t, pc)) = f2h<IO>(live ? s[t] : 0.f);
            *reinterpret_cast<IO*>(btu + sw128_off(NT + t, pc)) = f2h<IO>(d0);
          }
          if (t + 1 < p.T) {
            *reinterpret_cast<IO*>(btu + sw128_off(t + 1, pc)) = f2h<IO>(live ? s[t + 1] : 0.f);
            *reinterpret_cast<IO*>(btu + sw128_off(NT + t + 1, pc)) = f2h<IO>(d1);
          }
        } else {
          pk[t / 2] = 0u;
          pk[16 + t / 2] = 0u;
        }
      }
      tmem_st16(tmem + lane_addr + u * 96, pk);
      tmem_st16(tmem + lane_addr + u * 96 + 16, pk + 16);
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[u]);
    };
    auto dh_phase = [&](int it) {
      const int u = it & 1, k = it >> 1;
      const int pix = (blockIdx.x + it * p.ctas_per_sample) * 128 + px;
      mbar_wait(&dh_full[u], k & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < C; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + lane_addr + u * 96 + 2 * NT + c0, v);
        tmem_ld_wait();
        if (pix < p.HW) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dh[(size_t)(c0 + j) * p.HW + pix] = f2h<IO>(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dh_empty[u]);
    };
    for (int it = 0; it < ntile; ++it) {
      softmax_phase(it);
      if (it > 0) dh_phase(it - 1);
    }
    if (ntile > 0) dh_phase(ntile - 1);
    // ---- d(W.e) partial of this CTA ----
    float* part = p.part + ((size_t)b * p.ctas_per_sample + blockIdx.x) * C * p.T;
    if (ntile > 0) {
      mbar_wait(&acc_done, 0);
      tc_fence_after();
      float v[2 * NT];
      tmem_ld32(tmem + lane_addr + kAccCol, v);
      tmem_ld32(tmem + lane_addr + kAccCol + 32, v + 32);
      tmem_ld_wait();
      if (px < C) {
        for (int t = 0; t < NT; ++t) sAcc[px * NT + t] = v[t];                       // rows of dctx x columns of a
      } else if (px < 2 * C) {
        for (int t = 0; t < NT; ++t) sAcc[C * NT + (px - C) * NT + t] = v[NT + t];   // rows of h x columns of ds
      }
      named_bar_sync(1, 128);
      for (int i = threadIdx.x - 64; i < C * p.T; i += 128) {
        const int c = i / p.T, t = i - c * p.T;
        part[i] = sAcc[c * NT + t] + p.scale * sAcc[C * NT + c * NT + t];
      }
    } else {
      for (int i = threadIdx.x - 64; i < C * p.T; i += 128) part[i] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int word_attn_bwd_tc_supported(const void* images, const void* dctx, int64_t dctx_bs, const void* dattn, int C, int HW,
                               int T, int io_dtype) {
  if (io_dtype != AGB_BF16 && io_dtype != AGB_F16) return 0;
  if ((C != 16 && C != 32) || T > kBwdNT || HW % 8 != 0) return 0;
  if (dctx_bs != (int64_t)C * HW) return 0;
  if ((((uintptr_t)images | (uintptr_t)dctx) & 15) != 0) return 0;
  return 1;
}

int word_attn_bwd_tc_ctas(int B, int HW) {
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = cdiv(HW, 128);
  return std::max(1, std::min(cdiv(tiles, 2), std::max(1, (sms * 2) / B)));
}

int word_attn_bwd_tc(const void* images, const float* we, const int64_t* mask, const void* dctx, const void* dattn,
                     void* dimages, float* part, int ctas_per_sample, int B, int C, int HW, int T, int io_dtype,
                     float scale, cudaStream_t st) {
  const bool bf = io_dtype == AGB_BF16;
  CUtensorMap mapH, mapD;
  if (int rc = make_tmap_2d(&mapH, images, (uint64_t)B * C, (uint64_t)HW, (uint32_t)C, bf)) return rc;
  if (int rc = make_tmap_2d(&mapD, dctx, (uint64_t)B * C, (uint64_t)HW, (uint32_t)C, bf)) return rc;
  AttnBwdParams p;
  p.we = we; p.mask = mask; p.dattn = dattn; p.dh = dimages; p.part = part;
  p.B = B; p.C = C; p.HW = HW; p.T = T; p.scale = scale;
  p.tiles = cdiv(HW, 128);
  p.ctas_per_sample = ctas_per_sample;
  const int NT = kBwdNT;
  const int smem = kBwdStages * 4 * C * 128 + 4 * NT * 128 + 2 * 32 * 128 + 2 * 2 * (2 * NT) * 128 + 2 * C * NT * 4 + 1024;
  dim3 grid(ctas_per_sample, B);
  const int slot = prof_begin(PROF_ATTN_BWD, st);
#define AGB_ATTN_BWD_CASE(TLV)                                                                    \
  if (bf) {                                                                                       \
    auto kern = word_attn_bwd_tc_kernel<__nv_bfloat16, TLV>;                                      \
    AGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));      \
    kern<<<grid, 192, smem, st>>>(mapH, mapD, p);                                                 \
  } else {                                                                                        \
    auto kern = word_attn_bwd_tc_kernel<__half, TLV>;                                             \
    AGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));      \
    kern<<<grid, 192, smem, st>>>(mapH, mapD, p);                                                 \
  }
  switch ((T + 7) / 8) {
    case 1: AGB_ATTN_BWD_CASE(8) break;
    case 2: AGB_ATTN_BWD_CASE(16) break;
    case 3: AGB_ATTN_BWD_CASE(24) break;
    default: AGB_ATTN_BWD_CASE(32) break;
  }
#undef AGB_ATTN_BWD_CASE
  prof_end(slot, st);
  return check_launch("word_attn_bwd_tc_kernel");
}

// 1 when the tensor-core kernel can take this problem
int word_attn_tc_supported(const void* images, int C, int HW, int T, int io_dtype) {
  if (io_dtype != AGB_BF16 && io_dtype != AGB_F16) return 0;
  if (C % 16 != 0 || C < 16 || C > 64 || T > 64 || HW % 8 != 0) return 0;
  if (((uintptr_t)images & 15) != 0) return 0;
  return 1;
}

int word_attn_fwd_tc(const void* images, const float* we, const int64_t* mask, void* ctx, int64_t ctx_bs, void* attn,
                     int B, int C, int HW, int T, int io_dtype, float qscale, cudaStream_t st) {
  AttnFwdParams p;
  p.we = we; p.mask = mask; p.ctx = ctx; p.attn = attn; p.ctx_bs = ctx_bs;
  p.B = B; p.C = C; p.HW = HW; p.T = T; p.qscale = qscale;
  p.tiles = cdiv(HW, 128);
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // whole waves of resident CTAs (3 per SM by registers / shared memory), at least 2 tiles each
  p.ctas_per_sample = std::max(1, std::min(cdiv(p.tiles, 2), std::max(1, (sms * 3) / B)));
  if (io_dtype == AGB_BF16) return launch_attn_fwd_tc<__nv_bfloat16>(images, p, st);
  return launch_attn_fwd_tc<__half>(images, p, st);
}

}  // namespace tc
}  // namespace agb
