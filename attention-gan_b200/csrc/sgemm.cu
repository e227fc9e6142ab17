// Strided, batched fp32 GEMM on the CUDA cores: the arithmetic of the AGB_MATH_FP32 DAMSM path.
//
//   C[z][m,n] (+)= alpha * sum_{kb<KB} sum_{k<K} A[z][kb][m,k] * B[z][kb][k,n]
//
// Every operand is addressed by explicit element strides, so the transposes the reference
// materialises with .transpose().contiguous() (attention.py:94,109,115; words_loss.py:65-66) cost
// nothing here.  The second reduction level (kb) lets one launch contract over (image, region)
// pairs, which is what the gradient w.r.t. the word embeddings needs (sum over all images).
// Summation order is fixed -> deterministic, no atomics.
#include "agb_common.cuh"

namespace agb {

constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 256;

__global__ void __launch_bounds__(GEMM_THREADS)
sgemm_strided_kernel(SgemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* A = g.A + (int64_t)blockIdx.z * g.a_batch;
  const float* B = g.B + (int64_t)blockIdx.z * g.b_batch;
  float* C = g.C + (int64_t)blockIdx.z * g.c_batch;
  const int tx = tid & 15, ty = tid >> 4;

  // load mappings: keep the unit-stride index on consecutive threads
  const bool a_mfast = (g.a_m == 1);
  const bool b_nfast = (g.b_n == 1);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int kb = 0; kb < g.KB; ++kb) {
    const float* Ak = A + (int64_t)kb * g.a_kb;
    const float* Bk = B + (int64_t)kb * g.b_kb;
    for (int k0 = 0; k0 < g.K; k0 += BK) {
      float ra[4], rb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int m, k;
        if (a_mfast) { m = tid & 63; k = (tid >> 6) + 4 * j; }
        else { k = tid & 15; m = (tid >> 4) + 16 * j; }
        const int gm = m0 + m, gk = k0 + k;
        ra[j] = (gm < g.M && gk < g.K) ? Ak[(int64_t)gm * g.a_m + (int64_t)gk * g.a_k] : 0.f;
        int n, kk;
        if (b_nfast) { n = tid & 63; kk = (tid >> 6) + 4 * j; }
        else { kk = tid & 15; n = (tid >> 4) + 16 * j; }
        const int gn = n0 + n, gk2 = k0 + kk;
        rb[j] = (gn < g.N && gk2 < g.K) ? Bk[(int64_t)gk2 * g.b_k + (int64_t)gn * g.b_n] : 0.f;
      }
      __syncthreads();  // previous tile fully consumed
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int m, k;
        if (a_mfast) { m = tid & 63; k = (tid >> 6) + 4 * j; }
        else { k = tid & 15; m = (tid >> 4) + 16 * j; }
        As[k][m] = ra[j];
        int n, kk;
        if (b_nfast) { n = tid & 63; kk = (tid >> 6) + 4 * j; }
        else { kk = tid & 15; n = (tid >> 4) + 16 * j; }
        Bs[kk][n] = rb[j];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float* p = C + (int64_t)gm * g.c_m + (int64_t)gn * g.c_n;
      const float v = g.alpha * acc[i][j];
      *p = g.accumulate ? (*p + v) : v;
    }
  }
}

int sgemm_strided(const SgemmArgs& g, int batch, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || batch <= 0) return 0;
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), batch);
  if (grid.y > 65535 || grid.z > 65535) return fail_unsupported("sgemm grid too large");
  const int slot = prof_begin(PROF_SGEMM, st);
  sgemm_strided_kernel<<<grid, GEMM_THREADS, 0, st>>>(g);
  prof_end(slot, st);
  return check_launch("sgemm_strided_kernel");
}

}  // namespace agb
