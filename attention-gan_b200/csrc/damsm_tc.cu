// DAMSM word-region similarity on the 5th-generation tensor cores (AGB_MATH_TC_F16 / _BF16).
//
// Replaces the loop body of WordsLoss.get_loss (reference losses/words_loss.py:43-86) and the
// func_attention it calls (networks/attention.py:82-121) with ONE persistent kernel:
//
//   work item  = (image b, word tile c): up to 128 caption words (whole captions, packed by
//                their real lengths, so padded words cost nothing)
//   GEMM1      S[r, n]  = sum_d C_b[d, r] W[n, d]          3 M-tiles of 128 regions, N = 128, K = 256
//   epilogue 1 alpha = softmax over the words of each caption (thread = region: thread-local),
//              e = exp(gamma1 * alpha)  -> 16-bit, written K-major / 128B-swizzled to shared memory
//   GEMM2      V[n, d]  = sum_r e[r, n] C_b[d, r]          M = 128 words, N = 256, K = 320
//   epilogue 2 cos_n = <w_n, V_n> / (|w_n| |V_n|)   (thread = word: thread-local; the region-softmax
//              normaliser cancels in the cosine), m[b, i] = log sum_t exp(gamma2 cos)
//
// Operands reach the tensor core by TMA (128B swizzle, K-major); accumulators live in TMEM:
// columns [0,256) = two S buffers (MMA of M-tile j+1 overlaps epilogue 1 of M-tile j),
// columns [256,512) = V.  Warp roles: 0..11 = three epilogue warpgroups, 12 = TMA producer and
// TMEM allocator, 13 = MMA issuer (448 threads = 14 warps; with 4 warps on an SM sub-partition its 16K registers cap
// every thread at 128).
//
// HBM layout of the pre-packed operands (written by the pack kernels below, 16-bit):
//   Wh [tiles*128, 256]   packed caption words, K-major rows; unused rows are zero
//   Ct [Bi*384, 256]      regions as rows (M operand of GEMM1), rows >= R zero, except that rows 320.. repeat the
//                         regions 256..R-1 of the remainder tile (work sharing between the TMEM lane quarters)
//   Ck [Bi*256, 320]      features as rows (N operand of GEMM2), columns >= R zero
#include <algorithm>
#include <type_traits>

#include "tc_common.cuh"

namespace agb {

int sent_cos_fwd_launch(const float* cnn, const float* rnn, int Bi, int Bc, int D, float eps, float* scos_out,
                        cudaStream_t st);
int damsm_diag_att_maps(const float* img, const float* words, int64_t ws_b, int64_t ws_d, int64_t ws_t,
                        const int32_t* cap_lens, int Bi, int T, int D, int R, float gamma1, int row_offset,
                        float* att_out, float* S, float* Bt, cudaStream_t st);
size_t damsm_fp32_workspace_bytes(int Bi, int Bc, int T, int D, int R);
int damsm_fp32_bwd(const float* img, const float* words, int64_t ws_b, int64_t ws_d, int64_t ws_t,
                   const int32_t* cap_lens, int Bi, int Bc, int T, int D, int R, float gamma1, float gamma2,
                   float eps, const float* dm, const float* gscale, float* dimg, float* dwords, void* workspace,
                   size_t workspace_bytes, cudaStream_t st);

namespace tc {

constexpr int kD = 256;                   // feature dim: K of GEMM1, N of GEMM2
constexpr int kTileN = 128;               // word rows per tile
constexpr int kRRows = 384;               // region rows per image in Ct (3 M-tiles)
constexpr int kRCols = 320;               // region columns per image in Ck / e (5 chunks of 64)
constexpr int kChunk = 128 * 128;         // one [128 x 64] 16-bit K-major tile: 16 KB
// per-item cost model of the pair kernels, in units of one caption word slot (fitted to per-CTA times of the
// streaming backward, scripts/bwd3_timeline.py): fixed part per (word tile, image) item and per caption
constexpr int kItemCost0 = 900, kItemCostCap = 32;
constexpr int kThreads = 448;             // warps 0-11 epilogue, 12 TMA producer (+ TMEM alloc), 13 MMA issuer
constexpr int kSmemW = 0;                 // [4][128 x 64]  resident word tile       64 KB
constexpr int kSmemE = 4 * kChunk;        // [5][128 x 64]  e = exp(gamma1 alpha)    80 KB

template <typename T16> __device__ __forceinline__ T16 cvt16(float v);
template <> __device__ __forceinline__ __half cvt16<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T16> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// ---------------------------------------------------------------------------------------------
// pack kernels
// ---------------------------------------------------------------------------------------------
// greedy packing of whole captions into units of <= U word rows (U = 128: one unit per tile; U = 64 in split
// precision: two units per tile, so that no caption straddles a 64-row boundary) and <= U captions.
// tile_cpre[t] = exclusive prefix of the per-item cost estimate of word tile t (see cta_items_balanced).
// unit_first / unit_ncap / nunits (U = 64 only): the caption range of every half tile; nunits = 2 * ntiles
// (a trailing half tile without captions is emitted when the count is odd).
__global__ void tile_pack_kernel(const int32_t* __restrict__ cap_lens, int Bc, int T, int U, int32_t* __restrict__ cap_row,
                                 int32_t* __restrict__ tile_first, int32_t* __restrict__ tile_ncap,
                                 int32_t* __restrict__ ntiles, int32_t* __restrict__ tile_cpre,
                                 int32_t* __restrict__ unit_first, int32_t* __restrict__ unit_ncap,
                                 int32_t* __restrict__ nunits) {
  __shared__ int lens_s[1024];
  const int upt = kTileN / U;                 // units per tile
  int row = 0, unit = 0, first = 0, tfirst = 0;
  int cost = kItemCost0, cpre = 0;
  auto close_unit = [&](int next_first) {
    if (unit_first) {
      unit_first[unit] = first;
      unit_ncap[unit] = next_first - first;
    }
    if ((unit + 1) % upt == 0) {              // the tile is complete
      const int tile = unit / upt;
      tile_first[tile] = tfirst;
      tile_ncap[tile] = next_first - tfirst;
      tile_cpre[tile] = cpre;
      cpre += cost;
      cost = kItemCost0;
      tfirst = next_first;
    }
    ++unit;
    first = next_first;
    row = 0;
  };
  for (int base = 0; base < Bc; base += 1024) {
    const int n = min(1024, Bc - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) lens_s[i] = min(max(cap_lens[base + i], 0), T);
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 0; k < n; ++k) {
        const int i = base + k, L = lens_s[k];
        // the epilogues read a caption's TMEM columns in windows of 4: keep the whole window inside the
        // unit (the last accumulator buffer ends at TMEM column 512)
        if (row + ((L + 3) & ~3) > U || i - first == U) close_unit(i);
        cap_row[i] = unit * U + row;
        row += L;
        if (L > 0) cost += ((L + 3) & ~3) + kItemCostCap;
      }
    }
  }
  if (threadIdx.x == 0) {
    close_unit(Bc);
    while (unit % upt != 0) close_unit(Bc);   // pad the last tile with empty units
    const int nt = unit / upt;
    tile_cpre[nt] = cpre;
    ntiles[0] = nt;
    if (nunits) nunits[0] = unit;
  }
}

template <typename T16>
__global__ void pack_words_kernel_tc(const float* __restrict__ words, int64_t ws_b, int64_t ws_d, int64_t ws_t,
                                     const int32_t* __restrict__ cap_lens, const int32_t* __restrict__ cap_row,
                                     T16* __restrict__ Wh, T16* __restrict__ Wl, float* __restrict__ pn, int T) {
  const int i = blockIdx.x, t = blockIdx.y;
  const int L = min(max(cap_lens[i], 0), T);
  if (t >= L) return;
  const size_t row = (size_t)cap_row[i] + t;
  const float* src = words + (int64_t)i * ws_b + (int64_t)t * ws_t;
  float ss = 0.f;
  for (int d = threadIdx.x; d < kD; d += blockDim.x) {
    const float v = src[(int64_t)d * ws_d];
    const T16 h = cvt16<T16>(v);
    Wh[row * kD + d] = h;
    if (Wl) Wl[row * kD + d] = cvt16<T16>(v - to_f32(h));       // split precision: w = hi + lo
    ss = fmaf(v, v, ss);
  }
  __shared__ float red[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    pn[row] = sqrtf(tot);
  }
}

// img [Bi,256,R] fp32 -> Ck [Bi*256, 320] (same orientation, padded) and Ct [Bi*384, 256] (transposed)
template <typename T16>
__global__ void pack_img_kernel_tc(const float* __restrict__ img, T16* __restrict__ Ck, T16* __restrict__ Ct,
                                   T16* __restrict__ Ckl, T16* __restrict__ Ctl, int R) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, d0 = blockIdx.y * 32, r0 = blockIdx.x * 32;  // r0 < 384
  const int tx = threadIdx.x, ty = threadIdx.y;                            // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int d = d0 + k, r = r0 + tx;
    const float v = (r < R) ? img[((size_t)b * kD + d) * R + r] : 0.f;
    tile[k][tx] = v;
    if (r < kRCols) {
      const T16 h = cvt16<T16>(v);
      Ck[((size_t)b * kD + d) * kRCols + r] = h;
      if (Ckl) Ckl[((size_t)b * kD + d) * kRCols + r] = cvt16<T16>(v - to_f32(h));
    }
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, d = d0 + tx;
    const float v = tile[tx][k];
    const T16 h = cvt16<T16>(v);
    const T16 l = cvt16<T16>(v - to_f32(h));
    const bool dup_target = r >= 320 && r - 64 < R;       // written below by the block that owns region r - 64
    if (!dup_target) {
      Ct[((size_t)b * kRRows + r) * kD + d] = h;
      if (Ctl) Ctl[((size_t)b * kRRows + r) * kD + d] = l;
    }
    // the regions of the third (remainder) tile once more in its lanes 64..127 (rows 320..): the pair kernels let
    // lane quarters 2, 3 work on them with the other half of the captions (see `dup` in the epilogues)
    if (r >= 256 && r < R && r + 64 < kRRows) {
      Ct[((size_t)b * kRRows + r + 64) * kD + d] = h;
      if (Ctl) Ctl[((size_t)b * kRRows + r + 64) * kD + d] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// epilogue 1 for one caption and one region (thread): NL = caption length rounded up to 4.
// Branch-free: slots t >= L are set to -inf (their exponentials are 0) and only their stores are
// predicated; slots below NL-4 are always live, so they carry no predicate at all.
// ---------------------------------------------------------------------------------------------
template <typename T16> __device__ __forceinline__ uint16_t bits16(float v) {
  const T16 h = cvt16<T16>(v);
  return *reinterpret_cast<const uint16_t*>(&h);
}

// K region tiles at once (K = 2: the two score buffers of region tiles 0 and 1 are processed together,
// which gives every thread two independent dependency chains to overlap MUFU / TMEM latencies)
template <typename T16, int NL, int K>
__device__ __forceinline__ void caption_softmax(const uint32_t (&taddr)[K], int L, int off, const int (&r)[K], int R,
                                                uint32_t sE32, float scale_log2, float g1_log2) {
  float s[K][NL];
#pragma unroll
  for (int k = 0; k < K; ++k) tmem_ld_n<NL>(taddr[k], s[k]);
  tmem_ld_wait();
  uint32_t colbase[K], x[K];
  bool store[K], live[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // element (row n, col r) of the K-major 128B-swizzled e tile: chunk (r>>6), 16-byte unit ((r&63)>>3) ^ (n&7)
    colbase[k] = sE32 + (uint32_t)(r[k] >> 6) * kChunk + ((r[k] & 7) << 1) + (uint32_t)off * 128;
    x[k] = ((r[k] & 63) >> 3) << 4;
    store[k] = r[k] < kRCols;
    live[k] = r[k] < R;          // padding regions (R <= r < 320) contribute zeros to V
  }
  float ksc[K], mxs[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int t = NL - 4; t < NL; ++t)
      if (t >= L) s[k][t] = -INFINITY;
    float m4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m4[i] = s[k][i];
#pragma unroll
    for (int t = 4; t < NL; ++t) m4[t & 3] = fmaxf(m4[t & 3], s[k][t]);
    mxs[k] = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale_log2;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < NL; ++t) {
      s[k][t] = exp2f(fmaf(s[k][t], scale_log2, -mxs[k]));  // exp((s - max)/sqrt(D))
      s4[t & 3] += s[k][t];
    }
    ksc[k] = g1_log2 / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int t = 0; t < NL; ++t) {
      const uint16_t e = live[k] ? bits16<T16>(exp2f(s[k][t] * ksc[k])) : (uint16_t)0;  // exp(gamma1 alpha)
      if (store[k] && (t < NL - 4 || t < L)) st_shared_u16(colbase[k] + t * 128 + ((((off + t) << 4) & 0x70) ^ x[k]), e);
    }
  }
}

#define AGB_CAPTION_SWITCH(L, CALL)                       \
  switch (((L) + 3) >> 2) {                               \
    case 1: { constexpr int NL = 4; CALL; } break;        \
    case 2: { constexpr int NL = 8; CALL; } break;        \
    case 3: { constexpr int NL = 12; CALL; } break;       \
    case 4: { constexpr int NL = 16; CALL; } break;       \
    case 5: { constexpr int NL = 20; CALL; } break;       \
    case 6: { constexpr int NL = 24; CALL; } break;       \
    case 7: { constexpr int NL = 28; CALL; } break;       \
    default: { constexpr int NL = 32; CALL; } break;      \
  }

struct FwdParams {
  const int32_t* tile_first;
  const int32_t* tile_ncap;
  const int32_t* ntiles;
  const int32_t* cap_row;
  const int32_t* cap_lens;
  const int32_t* tile_cpre;   // [ntiles + 1] exclusive prefix of the per-item cost of each word tile
  int uniform_split;          // != 0: equal item counts per CTA instead of equal cost (option damsm_uniform_split, for A/B timing)
  int dup_rem;                // != 0: region tile 2 carries its regions twice (lanes 0..63 and 64..127, see pack_img_kernel_tc)
  int img_block;              // > 0: item order (image block, word tile, image) with this many images per block (item_pos)
  const float* pn;
  float* m_out;
  int Bi, Bc, T, R;
  float scale_log2, g1_log2, gamma2;
  // training forward (kSave) / streaming backward: saved per (image, word row), Ntot = nt_max * 128 rows per image
  void* v16;             // [Bi, Ntot, 256] 16-bit normalised context vectors
  float4* rowst;         // [Bi, Ntot] {<w,v>, |v|, 1/Z, -}
  int Ntot;
};

#include "damsm_tc_fwd2.inc"
#include "damsm_tc_fwd2x.inc"

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TcPlan {
  int nt_max;
  size_t off_Wh, off_pn, off_caprow, off_tfirst, off_tncap, off_tcpre, off_ntiles, off_Ct, off_Ck, off_attS, off_attB;
  // fused backward staging, per chunk of ct word tiles (N = ct*128 word rows) and all Bi images
  int ct;        // tiles per chunk
  int splits;    // slices of the image range in the d words GEMM
  size_t off_E16, off_dV16, off_A116, off_dpp, off_dwp, off_m, total;
  // what the training forward saves for the streaming backward (damsm_bwd3_kernel); save == 0: the saved
  // context vectors would not fit the budget (or AGB_DAMSM_BWD=2) and the backward recomputes (damsm_bwd2_kernel)
  int save;
  size_t off_v16, off_rowst;
  // split precision (AGB_MATH_TC_F16X2): captions packed into 64-row half tiles, lo parts of every operand
  int split;
  size_t off_Wl, off_Ctl, off_Ckl, off_ufirst, off_uncap, off_nunits;
};

static TcPlan make_tc_plan(int Bi, int Bc, int T, int D, int R, bool split) {
  const Options& opt = options();
  TcPlan p;
  p.split = split ? 1 : 0;
  const int window = (T + 3) & ~3;              // a caption's TMEM columns are read in windows of 4
  if (split) {
    const int per_unit = 64 / window;           // captions that always fit in one 64-row half tile
    p.nt_max = ((Bc + per_unit - 1) / per_unit + 1) / 2;
  } else {
    const int per_tile = 128 / window;          // captions that always fit in one tile
    p.nt_max = (Bc + per_tile - 1) / per_tile;
  }
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 1024); return at; };
  p.off_Wh = take((size_t)p.nt_max * kTileN * kD * 2);
  p.off_pn = take((size_t)p.nt_max * kTileN * 4);
  p.off_caprow = take((size_t)Bc * 4);
  p.off_tfirst = take((size_t)p.nt_max * 4);
  p.off_tncap = take((size_t)p.nt_max * 4);
  p.off_tcpre = take((size_t)(p.nt_max + 1) * 4);
  p.off_ntiles = take(256);
  p.off_Ct = take((size_t)Bi * kRRows * kD * 2);
  p.off_Ck = take((size_t)Bi * kD * kRCols * 2);
  p.off_attS = take((size_t)Bi * T * R * 4);
  p.off_attB = take((size_t)Bi * T * R * 4);
  p.off_Wl = p.off_Ctl = p.off_Ckl = p.off_ufirst = p.off_uncap = p.off_nunits = 0;
  if (split) {
    p.off_Wl = take((size_t)p.nt_max * kTileN * kD * 2);
    p.off_Ctl = take((size_t)Bi * kRRows * kD * 2);
    p.off_Ckl = take((size_t)Bi * kD * kRCols * 2);
    p.off_ufirst = take((size_t)2 * p.nt_max * 4);
    p.off_uncap = take((size_t)2 * p.nt_max * 4);
    p.off_nunits = take(256);
  }
  const size_t all_rows = (size_t)Bi * p.nt_max * kTileN;
  p.save = (all_rows * (kD * 2 + 16) <= (size_t)std::max<long long>(0, opt.damsm_save_bytes) && opt.damsm_bwd != 2) ? 1 : 0;
  // staging of one backward chunk: E16 + A116 (+ dV16 for the recomputing backward) + dpp per (image, word row)
  const size_t tile_bytes = (size_t)Bi * kTileN * (kRCols * 2 * 2 + (p.save ? 0 : kD * 2) + 4);
  size_t ct = (size_t)std::max<long long>(0, opt.damsm_chunk_bytes) / tile_bytes;
  if (ct < 1) ct = 1;
  if (ct > (size_t)p.nt_max) ct = p.nt_max;
  p.ct = (int)ct;
  p.splits = std::max(1, std::min(Bi, opt.damsm_dw_splits));
  const size_t rows = (size_t)Bi * ct * kTileN;
  p.off_E16 = take(rows * kRCols * 2);
  p.off_dV16 = p.save ? p.off_E16 : take(rows * kD * 2);
  p.off_A116 = take(rows * kRCols * 2);
  p.off_dpp = take(rows * 4);
  p.off_dwp = take((size_t)p.splits * ct * kTileN * kD * 4);
  p.off_m = take((size_t)Bi * Bc * 4);
  p.off_v16 = p.off_rowst = o;
  if (p.save) {
    p.off_v16 = take(all_rows * kD * 2);
    p.off_rowst = take(all_rows * 16);
  }
  p.total = o;
  return p;
}

static int num_sms() { return device_sms(); }

struct Packed {
  void* Wh; float* pn; int32_t *cap_row, *tfirst, *tncap, *tcpre, *ntiles; void* Ct; void* Ck;
  void *Wl, *Ctl, *Ckl; int32_t *ufirst, *uncap, *nunits;     // split precision only
};

// pack the operands of one call: caption tiles, 16-bit words, 16-bit region features (both layouts)
template <typename T16>
static int run_pack(const float* img, const float* words, int64_t ws_b, int64_t ws_d, int64_t ws_t,
                    const int32_t* cap_lens, int Bi, int Bc, int T, int R, char* ws, const TcPlan& pl, Packed* out,
                    cudaStream_t st, bool already_packed = false) {
  T16* Wh = (T16*)(ws + pl.off_Wh);
  float* pn = (float*)(ws + pl.off_pn);
  int32_t* cap_row = (int32_t*)(ws + pl.off_caprow);
  int32_t* tfirst = (int32_t*)(ws + pl.off_tfirst);
  int32_t* tncap = (int32_t*)(ws + pl.off_tncap);
  int32_t* tcpre = (int32_t*)(ws + pl.off_tcpre);
  int32_t* ntiles = (int32_t*)(ws + pl.off_ntiles);
  T16* Ct = (T16*)(ws + pl.off_Ct);
  T16* Ck = (T16*)(ws + pl.off_Ck);
  T16* Wl = pl.split ? (T16*)(ws + pl.off_Wl) : nullptr;
  T16* Ctl = pl.split ? (T16*)(ws + pl.off_Ctl) : nullptr;
  T16* Ckl = pl.split ? (T16*)(ws + pl.off_Ckl) : nullptr;
  int32_t* ufirst = pl.split ? (int32_t*)(ws + pl.off_ufirst) : nullptr;
  int32_t* uncap = pl.split ? (int32_t*)(ws + pl.off_uncap) : nullptr;
  int32_t* nunits = pl.split ? (int32_t*)(ws + pl.off_nunits) : nullptr;
  if (!already_packed) {
    // Wh and pn are neighbours in the workspace: one memset for both (unused word rows must read as zero)
    AGB_CUDA(cudaMemsetAsync(Wh, 0, (pl.off_pn - pl.off_Wh) + (size_t)pl.nt_max * kTileN * 4, st));
    if (Wl) AGB_CUDA(cudaMemsetAsync(Wl, 0, (size_t)pl.nt_max * kTileN * kD * 2, st));
    tile_pack_kernel<<<1, 256, 0, st>>>(cap_lens, Bc, T, pl.split ? 64 : kTileN, cap_row, tfirst, tncap, ntiles, tcpre,
                                        ufirst, uncap, nunits);
    if (int rc = check_launch("tile_pack_kernel")) return rc;
    pack_words_kernel_tc<T16><<<dim3(Bc, T), 128, 0, st>>>(words, ws_b, ws_d, ws_t, cap_lens, cap_row, Wh, Wl, pn, T);
    if (int rc = check_launch("pack_words_kernel_tc")) return rc;
    pack_img_kernel_tc<T16><<<dim3(kRRows / 32, kD / 32, Bi), dim3(32, 8), 0, st>>>(img, Ck, Ct, Ckl, Ctl, R);
    if (int rc = check_launch("pack_img_kernel_tc")) return rc;
  }
  out->Wl = Wl; out->Ctl = Ctl; out->Ckl = Ckl; out->ufirst = ufirst; out->uncap = uncap; out->nunits = nunits;
  out->Wh = Wh; out->pn = pn; out->cap_row = cap_row; out->tfirst = tfirst; out->tncap = tncap; out->tcpre = tcpre; out->ntiles = ntiles;
  out->Ct = Ct; out->Ck = Ck;
  return 0;
}

template <typename T16>
static int launch_fwd(const Packed& pk, const int32_t* cap_lens, int Bi, int Bc, int T, int R, float gamma1,
                      float gamma2, float* m_out, char* ws, const TcPlan& pl, bool save, cudaStream_t st);

template <typename T16>
static int run_fwd(const float* img, const float* words, int64_t ws_b, int64_t ws_d, int64_t ws_t,
                   const int32_t* cap_lens, int Bi, int Bc, int T, int R, float gamma1, float gamma2, float* m_out,
                   char* ws, const TcPlan& pl, bool save, cudaStream_t st) {
  Packed pk;
  if (int rc = run_pack<T16>(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, Bc, T, R, ws, pl, &pk, st)) return rc;
  return launch_fwd<T16>(pk, cap_lens, Bi, Bc, T, R, gamma1, gamma2, m_out, ws, pl, save, st);
}

template <typename T16>
static int launch_fwd(const Packed& pk, const int32_t* cap_lens, int Bi, int Bc, int T, int R, float gamma1,
                      float gamma2, float* m_out, char* ws, const TcPlan& pl, bool save, cudaStream_t st) {
  save = save && pl.save;
  const bool bf = std::is_same<T16, __nv_bfloat16>::value;
  CUtensorMap mapW, mapCt, mapCk;
  if (int rc = make_tmap_2d(&mapW, pk.Wh, (uint64_t)pl.nt_max * kTileN, kD, 128, bf)) return rc;
  if (int rc = make_tmap_2d(&mapCt, pk.Ct, (uint64_t)Bi * kRRows, kD, 128, bf)) return rc;
  if (int rc = make_tmap_2d(&mapCk, pk.Ck, (uint64_t)Bi * kD, kRCols, 128, bf)) return rc;
  FwdParams p;
  p.tile_first = pk.tfirst; p.tile_ncap = pk.tncap; p.ntiles = pk.ntiles; p.cap_row = pk.cap_row; p.cap_lens = cap_lens;
  p.tile_cpre = pk.tcpre;
  p.uniform_split = options().damsm_uniform_split;
  p.img_block = options().damsm_img_block;
  p.dup_rem = R > 256 ? 1 : 0;
  p.pn = pk.pn; p.m_out = m_out; p.Bi = Bi; p.Bc = Bc; p.T = T; p.R = R;
  p.scale_log2 = kLog2e / sqrtf((float)kD);
  p.g1_log2 = gamma1 * kLog2e;
  p.gamma2 = gamma2;
  p.v16 = ws + pl.off_v16; p.rowst = (float4*)(ws + pl.off_rowst); p.Ntot = pl.nt_max * kTileN;
  if (pl.split) {
    if constexpr (std::is_same<T16, __half>::value) {
      CUtensorMap mapWh, mapWl, mapCtl, mapCkl;
      if (int rc = make_tmap_2d(&mapWh, pk.Wh, (uint64_t)pl.nt_max * kTileN, kD, 64, false)) return rc;
      if (int rc = make_tmap_2d(&mapWl, pk.Wl, (uint64_t)pl.nt_max * kTileN, kD, 64, false)) return rc;
      if (int rc = make_tmap_2d(&mapCtl, pk.Ctl, (uint64_t)Bi * kRRows, kD, 128, false)) return rc;
      if (int rc = make_tmap_2d(&mapCkl, pk.Ckl, (uint64_t)Bi * kD, kRCols, 128, false)) return rc;
      FwdXParams xp;
      xp.f = p; xp.unit_first = pk.ufirst; xp.unit_ncap = pk.uncap; xp.nunits = pk.nunits;
      auto kx = save ? damsm_fwd2x_kernel<true> : damsm_fwd2x_kernel<false>;
      AGB_CUDA(cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
      const long long max_items_x = (long long)Bi * pl.nt_max * 2;
      const int grid_x = (int)std::min<long long>(num_sms(), max_items_x);
      const int slot_x = prof_begin(PROF_DAMSM_TC_FWD, st);
      kx<<<grid_x, kThreads, kSmemBytes2, st>>>(mapWh, mapWl, mapCt, mapCtl, mapCk, mapCkl, xp);
      prof_end(slot_x, st);
      return check_launch("damsm_fwd2x_kernel");
    } else {
      return fail_unsupported("split precision needs fp16 operands");
    }
  }
  auto kern = save ? damsm_fwd2_kernel<T16, true> : damsm_fwd2_kernel<T16, false>;
  AGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
  const long long max_items = (long long)Bi * pl.nt_max;
  const int grid = (int)std::min<long long>(num_sms(), max_items);
  const int slot = prof_begin(PROF_DAMSM_TC_FWD, st);
  kern<<<grid, kThreads, kSmemBytes2, st>>>(mapW, mapCt, mapCk, p);
  prof_end(slot, st);
  return check_launch("damsm_fwd_kernel");
}

#include "damsm_tc_bwd2.inc"
#include "damsm_tc_bwd3.inc"
#include "damsm_tc_bwd2_host.inc"

}  // namespace tc

int damsm_tc_supported(int T, int D, int R) { return (D == tc::kD && T >= 1 && T <= 32 && R >= 1 && R <= tc::kRCols) ? 1 : 0; }

size_t damsm_tc_workspace_bytes(int Bi, int Bc, int T, int D, int R, int math) {
  return tc::make_tc_plan(Bi, Bc, T, D, R, math == AGB_MATH_TC_F16X2).total;
}

int damsm_tc_fwd(const float* img, const float* words, int64_t ws_b, int64_t ws_d, int64_t ws_t,
                 const int32_t* cap_lens, int Bi, int Bc, int T, int D, int R, float gamma1, float gamma2,
                 float eps, int row_offset, float* m_out, float* att_out, const float* cnn, const float* rnn,
                 float* scos_out, void* workspace, size_t workspace_bytes, int math, int save, cudaStream_t st) {
  if (Bi <= 0 || Bc <= 0) return fail_arg("non-positive batch");
  if (Bi > 65535) return fail_unsupported("Bi=%d > 65535", Bi);
  if (math != AGB_MATH_TC_BF16 && gamma1 > 11.f) return fail_unsupported("gamma1=%g overflows fp16 (use bf16 or fp32 math)", gamma1);
  const tc::TcPlan pl = tc::make_tc_plan(Bi, Bc, T, D, R, math == AGB_MATH_TC_F16X2);
  if (workspace_bytes < pl.total) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, pl.total);
    return AGB_E_WORKSPACE;
  }
  char* ws = (char*)workspace;
  int rc;
  // the matched-pair attention maps and the sentence cosine only read the caller's inputs: they run on a
  // forked stream beside the pack kernels and the pair kernel and join before this call returns
  tc::SideStream* side = nullptr;
  int rc_side = 0;
  if (att_out || cnn) {
    if (att_out && (row_offset < 0 || row_offset + Bi > Bc)) return fail_arg("row_offset=%d out of range", row_offset);
    if ((rc = tc::side_stream(&side))) return rc;
    if ((rc = tc::side_fork(side, st))) return rc;
    if (att_out)
      rc_side = damsm_diag_att_maps(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, T, D, R, gamma1, row_offset, att_out,
                                    (float*)(ws + pl.off_attS), (float*)(ws + pl.off_attB), side->stream);
    if (cnn && rc_side == 0) rc_side = sent_cos_fwd_launch(cnn, rnn, Bi, Bc, D, eps, scos_out, side->stream);
  }
  if (rc_side == 0) {
    if (math == AGB_MATH_TC_BF16)
      rc = tc::run_fwd<__nv_bfloat16>(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, Bc, T, R, gamma1, gamma2, m_out, ws, pl, save != 0, st);
    else
      rc = tc::run_fwd<__half>(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, Bc, T, R, gamma1, gamma2, m_out, ws, pl, save != 0, st);
  } else {
    rc = rc_side;
  }
  // joined on every path after the fork (an enclosing graph capture must not be left forked)
  if (side)
    if (int rj = tc::side_join(side, st)) return rc ? rc : rj;
  return rc;
}

int damsm_tc_bwd(const float* img, const float* words, int64_t ws_b, int64_t ws_d, int64_t ws_t,
                 const int32_t* cap_lens, int Bi, int Bc, int T, int D, int R, float gamma1, float gamma2,
                 float eps, const float* dm, const float* m_fwd, const float* gscale, float* dimg, float* dwords,
                 void* workspace, size_t workspace_bytes, int ws_from_fwd, int math, cudaStream_t st) {
  if (Bi <= 0 || Bc <= 0) return fail_arg("non-positive batch");
  if (Bi > 65535) return fail_unsupported("Bi=%d > 65535", Bi);
  const tc::TcPlan pl = tc::make_tc_plan(Bi, Bc, T, D, R, math == AGB_MATH_TC_F16X2);
  if (workspace_bytes < pl.total) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, pl.total);
    return AGB_E_WORKSPACE;
  }
  if (pl.save) {
    if (math == AGB_MATH_TC_BF16)
      return tc::run_bwd3<__nv_bfloat16>(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, Bc, T, R, gamma1, gamma2, dm, m_fwd,
                                         gscale, dimg, dwords, (char*)workspace, pl, ws_from_fwd, st);
    return tc::run_bwd3<__half>(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, Bc, T, R, gamma1, gamma2, dm, m_fwd, gscale,
                                dimg, dwords, (char*)workspace, pl, ws_from_fwd, st);
  }
  if (math == AGB_MATH_TC_BF16)
    return tc::run_bwd2<__nv_bfloat16>(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, Bc, T, R, gamma1, gamma2, dm, m_fwd,
                                       gscale, dimg, dwords, (char*)workspace, pl, ws_from_fwd, st);
  return tc::run_bwd2<__half>(img, words, ws_b, ws_d, ws_t, cap_lens, Bi, Bc, T, R, gamma1, gamma2, dm, m_fwd, gscale,
                              dimg, dwords, (char*)workspace, pl, ws_from_fwd, st);
}

}  // namespace agb

#ifdef AGB_TIMELINE
// debug build only (-DAGB_TIMELINE): the clock64 timeline CTA 0 of damsm_bwd3_kernel recorded, and per-CTA run times
extern "C" int agb_damsm_debug_timeline(long long* out, int n) {
  const size_t bytes = std::min((size_t)n * 8, sizeof(agb::tc::g_tl));
  return (int)cudaMemcpyFromSymbol(out, agb::tc::g_tl, bytes);
}
extern "C" int agb_damsm_debug_cta_times(long long* out, int n) {
  const size_t bytes = std::min((size_t)n * 8, sizeof(agb::tc::g_tl_cta));
  return (int)cudaMemcpyFromSymbol(out, agb::tc::g_tl_cta, bytes);
}
#endif
