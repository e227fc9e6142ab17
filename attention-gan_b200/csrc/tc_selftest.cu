// Self-test of the tcgen05 / TMEM / TMA building blocks (tc_common.cuh): one CTA computes
//   C[128, N] = A[128, K] * B[N, K]^T      (16-bit inputs, fp32 accumulate in TMEM)
// with B brought in by TMA (128B swizzle) and A either by TMA or written to shared memory by the
// threads through sw128_off() -- the two ways operands reach the tensor core in damsm_tc.cu.
// Exposed as agb_tc_selftest so the GPU tests can pin the descriptor encodings independently of
// the fused kernels.
#include "tc_common.cuh"

namespace agb {
namespace tc {

static EncodeTiledFn g_encode = nullptr;

int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                 bool bf16) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr) {
      set_error("cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
      return e != cudaSuccess ? (int)e : (int)cudaErrorUnknown;
    }
    g_encode = (EncodeTiledFn)fn;
  }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box_rows=%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return (int)cudaErrorInvalidValue;
  }
  return 0;
}

// smem: A [KC][128 rows x 128 B] then B [KC][N rows x 128 B]
__global__ void __launch_bounds__(128)
selftest_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const uint16_t* __restrict__ A_raw, float* __restrict__ C, int N, int K, int fmt, int manual_a) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int KC = K / 64;
  unsigned char* sA = smem;
  unsigned char* sB = smem + (size_t)KC * 128 * 128;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  if (manual_a) {  // A[row, k] -> swizzled K-major tile, chunk by chunk
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
      const int row = i / K, k = i - row * K;
      *reinterpret_cast<uint16_t*>(sA + (size_t)(k >> 6) * 128 * 128 + sw128_off(row, k & 63)) = A_raw[i];
    }
    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && elect_one()) {
    const uint32_t bytes = (uint32_t)KC * (uint32_t)(N * 128) + (manual_a ? 0u : (uint32_t)KC * 128u * 128u);
    mbar_expect_tx(&bar_full, bytes);
    for (int kc = 0; kc < KC; ++kc) {
      if (!manual_a) tma_load_2d(sA + (size_t)kc * 128 * 128, &mapA, &bar_full, kc * 64, 0);
      tma_load_2d(sB + (size_t)kc * N * 128, &mapB, &bar_full, kc * 64, 0);
    }
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(128, N, fmt);
    for (int kc = 0; kc < KC; ++kc) {
      const uint64_t da = make_desc_sw128(smem_u32(sA + (size_t)kc * 128 * 128));
      const uint64_t db = make_desc_sw128(smem_u32(sB + (size_t)kc * N * 128));
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, da + 2 * kk, db + 2 * kk, idesc, (kc | kk) ? 1u : 0u);
    }
    umma_commit(&bar_done);
  }
  __syncwarp();
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  // thread = output row (TMEM lane), 32 columns at a time
  const int row = threadIdx.x;
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace tc
}  // namespace agb

using namespace agb;

// A [128,K], B [N,K] 16-bit row-major device buffers; C [128,N] fp32.  K % 64 == 0, K <= 256,
// N in {128, 256}.  bf16 != 0 selects bfloat16 operands.  manual_a != 0 stages A through the
// thread-written swizzled path.
extern "C" int agb_tc_selftest(const void* A, const void* B, float* C, int N, int K, int bf16, int manual_a,
                               void* stream) {
  if (!A || !B || !C) return fail_arg("null pointer");
  if (K <= 0 || K % 64 || K > 256 || (N != 128 && N != 256)) return fail_unsupported("selftest shape N=%d K=%d", N, K);
  CUtensorMap mapA, mapB;
  if (int rc = tc::make_tmap_2d(&mapA, A, 128, K, 128, bf16 != 0)) return rc;
  if (int rc = tc::make_tmap_2d(&mapB, B, N, K, N, bf16 != 0)) return rc;
  const size_t smem = (size_t)(K / 64) * (128 + N) * 128 + 1024;
  AGB_CUDA(cudaFuncSetAttribute(tc::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mapA, mapB, (const uint16_t*)A, C, N, K, bf16 ? 1 : 0,
                                                            manual_a);
  return check_launch("selftest_kernel");
}
