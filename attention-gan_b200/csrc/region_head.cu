// Region-feature head of the image encoder on the 5th-generation tensor cores.
//
// Replaces  CNNEncoder.emb_features = conv1x1(768 -> 256)  applied to the Mixed_6e feature map
// (reference networks/cnn_encoder.py:56,101; utilities/layers.py:46-48: kernel 1, no bias) and its autograd:
//
//   forward   feat[b][o, r] = sum_c W[o, c] x[b][c, r]                       one [Cout x Cin] x [Cin x R] GEMM per image
//   backward  dW[o, c]      = sum_b sum_r dfeat[b][o, r] x[b][c, r]          (the main trainable CNN weight of the
//                                                                             DAMSM pretraining, pretrain_damsm.py:70-74)
//             dx[b][c, r]   = sum_o W[o, c] dfeat[b][o, r]                   (optional: the Inception trunk is frozen)
//
// PyTorch runs this convolution through cuDNN, by default with TF32 operands (10 mantissa bits; exact fp32 only with
// torch.backends.cudnn.allow_tf32 = False, then on CUDA cores): 29 GFLOP per direction at batch 256.
// Here the operands are cast once to 16 bit (rows padded from R = 289 to 320 columns: a 578-byte row pitch is not
// TMA-addressable) and the GEMMs run on the batched tcgen05 kernel of tc_gemm.cu:
//   * forward in SPLIT precision: x = x_hi + x_lo, W = W_hi + W_lo (bf16 pairs, 16 significant bits); W_hi x_hi +
//     W_lo x_hi (one launch, two operand pairs) + W_hi x_lo (accumulating launch): features within ~1e-5 of fp32
//     arithmetic (PyTorch's default for this convolution on the GPU is TF32: 5e-4), because they feed the loss that is
//     held to 1e-4;
//   * backward with plain bf16 operands (the gradients can be ~1e-8: bf16 has fp32's range; their precision needs
//     are those of a gradient), reusing x_hi / W_hi of the forward call when the caller hands the workspace back.
// The features leave as fp32 [B, Cout, R] (the layout the reference hands to WordsLoss); packing them into the 16-bit
// operand layouts of the pair kernels stays in pack_img_kernel_tc (0.4 % of the step at batch 2048).
#include <algorithm>

#include "tc_common.cuh"

namespace agb {
namespace {

constexpr int kPad = 64;
constexpr int kRowsPerBlock = 8;
static unsigned cast_grid(size_t rows) { return (unsigned)((rows + kRowsPerBlock - 1) / kRowsPerBlock); }

// src fp32 [rows, R] -> hi (and lo) 16-bit [rows, Rp], columns >= R zero
template <typename T16>
__global__ void cast_pad_kernel(const float* __restrict__ src, T16* __restrict__ hi, T16* __restrict__ lo, int R,
                                int Rp, size_t rows) {
  // kRowsPerBlock consecutive rows per CTA: the source rows are contiguous, so the block streams one span
  const size_t row0 = (size_t)blockIdx.x * kRowsPerBlock;
  const int nrow = (int)min((size_t)kRowsPerBlock, rows > row0 ? rows - row0 : 0);
  for (int i = threadIdx.x; i < nrow * Rp; i += blockDim.x) {
    const int k = i / Rp, r = i - k * Rp;
    const size_t row = row0 + k;
    const float v = r < R ? src[row * R + r] : 0.f;
    const T16 a = from_f32<T16>(v);
    hi[row * Rp + r] = a;
    if (lo) lo[row * Rp + r] = from_f32<T16>(v - to_f32(a));
  }
}

// out[i] = sum_s part[s][i]   (fixed order: deterministic)
__global__ void sum_slices_kernel(const float* __restrict__ part, int slices, size_t n, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int s = 0; s < slices; ++s) acc += part[(size_t)s * n + i];
  out[i] = acc;
}

struct HeadPlan {
  int Rp, splits;
  size_t off_xh, off_xl, off_wh, off_wl, off_db, off_part, total;
};
HeadPlan make_plan(int B, int Cin, int Cout, int R) {
  HeadPlan p;
  p.Rp = (R + kPad - 1) / kPad * kPad;
  p.splits = std::max(1, std::min(B, 48));
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 1024); return at; };
  const size_t x16 = (size_t)B * Cin * p.Rp * 2, w16 = (size_t)Cout * Cin * 2;
  p.off_xh = take(x16);
  p.off_xl = take(x16);
  p.off_wh = take(w16);
  p.off_wl = take(w16);
  p.off_db = take((size_t)B * Cout * p.Rp * 2);               // bf16 d feat
  p.off_part = take((size_t)p.splits * Cout * Cin * 4);
  p.total = o;
  return p;
}

int check(int B, int Cin, int Cout, int R) {
  if (B <= 0 || Cin <= 0 || Cout <= 0 || R <= 0) return fail_arg("non-positive size B=%d Cin=%d Cout=%d R=%d", B, Cin, Cout, R);
  if (Cin % 64 != 0 || Cout % 128 != 0) return fail_unsupported("region head: Cin=%d must be a multiple of 64, Cout=%d of 128", Cin, Cout);
  if (B > 65535) return fail_unsupported("B=%d > 65535", B);
  return 0;
}

}  // namespace
}  // namespace agb

using namespace agb;

extern "C" size_t agb_region_head_workspace_bytes(int B, int Cin, int Cout, int R) {
  if (B <= 0 || Cin <= 0 || Cout <= 0 || R <= 0 || Cin % 64 || Cout % 128) return 0;
  return make_plan(B, Cin, Cout, R).total;
}

extern "C" int agb_region_head_fwd(const float* x, const float* w, float* feat, void* workspace, size_t workspace_bytes,
                                   int B, int Cin, int Cout, int R, void* stream) {
  if (int rc = check(B, Cin, Cout, R)) return rc;
  if (!x || !w || !feat || !workspace) return fail_arg("null pointer");
  const HeadPlan pl = make_plan(B, Cin, Cout, R);
  if (workspace_bytes < pl.total) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, pl.total);
    return AGB_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  __nv_bfloat16* xh = (__nv_bfloat16*)(ws + pl.off_xh);
  __nv_bfloat16* xl = (__nv_bfloat16*)(ws + pl.off_xl);
  __nv_bfloat16* wh = (__nv_bfloat16*)(ws + pl.off_wh);
  __nv_bfloat16* wl = (__nv_bfloat16*)(ws + pl.off_wl);
  cast_pad_kernel<__nv_bfloat16><<<cast_grid((size_t)B * Cin), 256, 0, st>>>(x, xh, xl, R, pl.Rp, (size_t)B * Cin);
  if (int rc = check_launch("cast_pad_kernel")) return rc;
  cast_pad_kernel<__nv_bfloat16><<<cast_grid(Cout), 256, 0, st>>>(w, wh, wl, Cin, Cin, (size_t)Cout);
  if (int rc = check_launch("cast_pad_kernel")) return rc;

  CUtensorMap mWh, mWl, mXh, mXl;
  if (int rc = tc::make_tmap_2d(&mWh, wh, Cout, Cin, 128, true)) return rc;
  if (int rc = tc::make_tmap_2d(&mWl, wl, Cout, Cin, 128, true)) return rc;
  if (int rc = tc::make_tmap_2d(&mXh, xh, (uint64_t)B * Cin, pl.Rp, 64, true)) return rc;
  if (int rc = tc::make_tmap_2d(&mXl, xl, (uint64_t)B * Cin, pl.Rp, 64, true)) return rc;
  // feat[z][o, r] = sum_c W[o, c] x[z][c, r]:  A = W K-major (rows o, cols c), B = x[z] MN-major (rows c, cols r)
  tc::TcGemmArgs g{};
  g.a_mn = 0; g.b_mn = 1; g.bf16 = 1; g.M = Cout; g.N = R; g.K = Cin; g.KB = 1;
  g.NT = 128; g.NT0 = (pl.Rp == 320 && R > 192) ? 192 : 0; g.MT = 2;
  g.b_zrow = Cin; g.b2_zrow = Cin;
  g.C = feat; g.c_z = (int64_t)Cout * R; g.c_m = R; g.c_n = 1; g.alpha = 1.f;
  g.nsrc = 2; g.accumulate = 0; g.prof_tag = PROF_DAMSM_PACK;
  if (int rc = tc::tc_gemm2(g, mWh, mXh, mWl, mXh, B, st)) return rc;          // W_hi x_hi + W_lo x_hi
  g.nsrc = 1; g.accumulate = 1;
  return tc::tc_gemm(g, mWh, mXl, B, st);                                      // + W_hi x_lo
}

extern "C" int agb_region_head_bwd(const float* x, const float* w, const float* dfeat, float* dw, float* dx,
                                   void* workspace, size_t workspace_bytes, int ws_from_fwd, int B, int Cin, int Cout,
                                   int R, void* stream) {
  if (int rc = check(B, Cin, Cout, R)) return rc;
  if (!x || !w || !dfeat || !workspace || (!dw && !dx)) return fail_arg("null pointer");
  const HeadPlan pl = make_plan(B, Cin, Cout, R);
  if (workspace_bytes < pl.total) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, pl.total);
    return AGB_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  __nv_bfloat16* xb = (__nv_bfloat16*)(ws + pl.off_xh);       // the hi parts double as the plain bf16 operands
  __nv_bfloat16* wb = (__nv_bfloat16*)(ws + pl.off_wh);
  __nv_bfloat16* db = (__nv_bfloat16*)(ws + pl.off_db);
  float* part = (float*)(ws + pl.off_part);
  cast_pad_kernel<__nv_bfloat16><<<cast_grid((size_t)B * Cout), 256, 0, st>>>(dfeat, db, nullptr, R, pl.Rp, (size_t)B * Cout);
  if (int rc = check_launch("cast_pad_kernel")) return rc;
  CUtensorMap mD_k, mD_mn, mX_k, mW_mn;
  if (dw) {
    if (!ws_from_fwd) {
      cast_pad_kernel<__nv_bfloat16><<<cast_grid((size_t)B * Cin), 256, 0, st>>>(x, xb, nullptr, R, pl.Rp, (size_t)B * Cin);
      if (int rc = check_launch("cast_pad_kernel")) return rc;
    }
    // dW[o, c] = sum_b sum_r dfeat[b][o, r] x[b][c, r]: both operands K-major over the (zero-padded) regions,
    // the images are the reduction blocks, cut into `splits` slices whose partial sums are added in a fixed order
    if (int rc = tc::make_tmap_2d(&mD_k, db, (uint64_t)B * Cout, pl.Rp, 128, true)) return rc;
    if (int rc = tc::make_tmap_2d(&mX_k, xb, (uint64_t)B * Cin, pl.Rp, 256, true)) return rc;
    tc::TcGemmArgs h{};
    h.a_mn = 0; h.b_mn = 0; h.bf16 = 1; h.M = Cout; h.N = Cin; h.K = pl.Rp; h.NT = 256; h.MT = 2;
    h.split_kb = 1; h.kb_total = B; h.KB = (B + pl.splits - 1) / pl.splits;
    h.a_kbrow = Cout; h.b_kbrow = Cin;
    h.C = part; h.c_z = (int64_t)Cout * Cin; h.c_m = Cin; h.c_n = 1; h.alpha = 1.f; h.accumulate = 0;
    h.prof_tag = PROF_DAMSM_PACK;
    if (int rc = tc::tc_gemm(h, mD_k, mX_k, pl.splits, st)) return rc;
    const size_t n = (size_t)Cout * Cin;
    sum_slices_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, pl.splits, n, dw);
    if (int rc = check_launch("sum_slices_kernel")) return rc;
  }
  if (dx) {
    if (!ws_from_fwd) {
      cast_pad_kernel<__nv_bfloat16><<<cast_grid(Cout), 256, 0, st>>>(w, wb, nullptr, Cin, Cin, (size_t)Cout);
      if (int rc = check_launch("cast_pad_kernel")) return rc;
    }
    // dx[z][c, r] = sum_o W[o, c] dfeat[z][o, r]: A = W MN-major (rows o = k, cols c = m), B = dfeat[z] MN-major
    if (int rc = tc::make_tmap_2d(&mW_mn, wb, Cout, Cin, 64, true)) return rc;
    if (int rc = tc::make_tmap_2d(&mD_mn, db, (uint64_t)B * Cout, pl.Rp, 64, true)) return rc;
    tc::TcGemmArgs g{};
    g.a_mn = 1; g.b_mn = 1; g.bf16 = 1; g.M = Cin; g.N = R; g.K = Cout; g.KB = 1;
    g.NT = 128; g.NT0 = (pl.Rp == 320 && R > 192) ? 192 : 0; g.MT = 2;
    g.b_zrow = Cout;
    g.C = dx; g.c_z = (int64_t)Cin * R; g.c_m = R; g.c_n = 1; g.alpha = 1.f; g.accumulate = 0;
    g.prof_tag = PROF_DAMSM_PACK;
    if (int rc = tc::tc_gemm(g, mW_mn, mD_mn, B, st)) return rc;
  }
  return 0;
}
