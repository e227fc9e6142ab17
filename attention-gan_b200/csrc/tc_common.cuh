// sm_100a primitives used by the tensor-core kernels: tcgen05 (UMMA) issue / commit / TMEM
// allocation and loads, 2-D TMA tensor-map loads, shared-memory matrix descriptors.
// Layout conventions (K-major operands, 128-byte swizzle) follow the canonical UMMA layouts:
//   a [rows x 64] tile of 16-bit elements is rows of 128 B; 8-row groups are 1024 B apart (SBO);
//   inside a group the 16-byte chunk index is XORed with (row & 7).  TMA with
//   CU_TENSOR_MAP_SWIZZLE_128B produces exactly this layout when the tile base is 1024-B aligned.
#pragma once
#include <cuda.h>

#include "agb_common.cuh"

namespace agb {
namespace tc {

// ---- tensor maps (host) ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

// 2-D row-major 16-bit tensor [rows, cols] (cols contiguous), box = [box_rows, 64 cols], 128B swizzle
int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                 bool bf16);

constexpr int kChunkBytes16 = 128 * 128;   // one [128 x 64] 16-bit swizzled tile

// batched GEMM on tcgen05 (tc_gemm.cu):  C[z][m,n] (+)= alpha * sum_kb sum_k A[z][kb][m,k] B[z][kb][n,k]
struct TcGemmArgs {
  int a_mn, b_mn;              // 0: K-major array (rows = m|n, cols = k), 1: MN-major (rows = k, cols = m|n)
  int bf16;                    // operand type: 0 = fp16, 1 = bf16
  int M, N, K, KB, NT;         // K % 64 == 0; NT = 64, 128, 192 or 256 output columns per CTA (MT * NT <= 512)
  int64_t a_zrow, a_zcol, a_kbrow, a_kbcol;   // offsets into map A per batch z / reduction block kb
  int64_t b_zrow, b_zcol, b_kbrow, b_kbcol;
  float* C;
  int64_t c_z, c_m, c_n;       // fp32 output strides
  float alpha;
  int accumulate;
  // data-dependent extent: valid = clamp((*dyn_tiles - dyn_t0) * 128, 0, extent) rows of dimension
  // dyn_dim (0 = none, 1 = M, 3 = K); CTAs / K-chunks beyond it do no work
  const int32_t* dyn_tiles;
  int dyn_t0, dyn_dim;
  // split_kb != 0: blockIdx.z enumerates slices of the kb range (kb = z*KB + i < kb_total) instead of
  // independent batches; slice z writes its partial sum to C + z*c_z
  int split_kb, kb_total;
  int stages;                  // ring depth, set by tc_gemm()
  int MT;                      // 128-row output tiles per CTA (1 or 2), sharing the B operand
  int nsrc;                    // 1, or 2: a second (A2, B2) operand pair continues the same reduction
  int64_t a2_zrow, b2_zrow;    // batch row offsets of the second pair
  int64_t a2_zcol, b2_zcol;    // batch column offsets of the second pair
  int prof_tag;                // agb::ProfTag the launch is timed under (0 = PROF_DAMSM_TC_BWD)
  int NT0;                     // != 0 (MN-major B only): CTA column 0 covers NT0 output columns, the others NT each
};
int tc_gemm(const TcGemmArgs& g, const CUtensorMap& mapA, const CUtensorMap& mapB, int batch, cudaStream_t st);
int tc_gemm2(const TcGemmArgs& g, const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapA2,
             const CUtensorMap& mapB2, int batch, cudaStream_t st);

// ---- PTX wrappers (device) --------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], one elected thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// K-major, SWIZZLE_128B operand descriptor for a tile whose rows are 128 B (64 x 16-bit)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                  // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, both operands K-major; fmt 0 = f16, 1 = bf16
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x NC consecutive 32-bit columns (thread i <- lane base+i)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// NL (multiple of 4, <= 32) consecutive columns
template <int NL>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, float* v) {
  if constexpr (NL >= 32) tmem_ld32(taddr, v);
  else if constexpr (NL >= 16) { tmem_ld16(taddr, v); if constexpr (NL > 16) tmem_ld_n<NL - 16>(taddr + 16, v + 16); }
  else if constexpr (NL >= 8) { tmem_ld8(taddr, v); if constexpr (NL > 8) tmem_ld_n<NL - 8>(taddr + 8, v + 8); }
  else tmem_ld4(taddr, v);
}
// ---- fragment-layout TMEM access (16 lanes per instruction) -------------------------------------
// tcgen05.ld.16x256b.xN: thread i of the warp receives, for column block k < N (8 fp32 columns
// each), rows (i/4) and (i/4 + 8) of the 16 lanes at `taddr`, columns 8k + 2(i%4) + {0,1}:
//   v[4k + 0], v[4k + 1] = row i/4,     cols 8k + 2(i%4), +1
//   v[4k + 2], v[4k + 3] = row i/4 + 8, same columns
// i.e. the mma accumulator fragment layout, so a packed 16-bit pair per (row, column pair) can go
// straight to stmatrix.trans (transposed 16-byte rows in shared memory) or back into TMEM as a
// 16-bit A operand with tcgen05.st.16x128b (32-bit column 4k + i%4 of the same rows).
__device__ __forceinline__ void tmem_ldf1(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ldf2(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ldf4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// KB column blocks of 8 (KB <= 8) -> v[4*KB]
template <int KB>
__device__ __forceinline__ void tmem_ld_frag(uint32_t taddr, float* v) {
  if constexpr (KB >= 4) { tmem_ldf4(taddr, v); if constexpr (KB > 4) tmem_ld_frag<KB - 4>(taddr + 32, v + 16); }
  else if constexpr (KB >= 2) { tmem_ldf2(taddr, v); if constexpr (KB > 2) tmem_ld_frag<KB - 2>(taddr + 16, v + 8); }
  else tmem_ldf1(taddr, v);
}
// tcgen05.st.16x128b.xN: r[2k + j] = 32-bit column 4k + i%4 of row i/4 + 8j (16 lanes at taddr)
__device__ __forceinline__ void tmem_stf2(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_stf4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_stf8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// four 8x8 16-bit matrices, transposed: thread i supplies the address of the 16-byte row (i%8) of
// matrix (i/8); register m holds elements [2(i%4)+{0,1}][i/4] of stored matrix m
__device__ __forceinline__ void stsm_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};"
               ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}
// PB blocks of 4 32-bit columns (PB in {2, 4, 6, 8}) from r[2 * PB]
template <int PB>
__device__ __forceinline__ void tmem_st_frag(uint32_t taddr, const uint32_t* r) {
  if constexpr (PB >= 8) tmem_stf8(taddr, r);
  else if constexpr (PB >= 4) { tmem_stf4(taddr, r); if constexpr (PB > 4) tmem_st_frag<PB - 4>(taddr + 16, r + 8); }
  else tmem_stf2(taddr, r);
}
// two matrices: threads 0-15 supply the row addresses (matrix i/8, row i%8)
__device__ __forceinline__ void stsm_x2_trans(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1,%2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
template <typename T16> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// ---- thread = lane TMEM stores (32x32b) and packed fragment loads (16x128b) ----------------------
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// N (multiple of 4, <= 32) consecutive 32-bit columns of this thread's lane
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r) {
  if constexpr (N >= 16) { tmem_st16(taddr, r); if constexpr (N > 16) tmem_st_n<N - 16>(taddr + 16, r + 16); }
  else if constexpr (N >= 8) { tmem_st8(taddr, r); if constexpr (N > 8) tmem_st_n<N - 8>(taddr + 8, r + 8); }
  else tmem_st4(taddr, r);
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.ld.16x128b.xN: r[2k + j] = 32-bit column 4k + i%4 of row i/4 + 8j of the 16 lanes at taddr.
// When the columns hold packed 16-bit pairs (t = 2c, 2c+1) this is the stmatrix fragment of the
// pixel x word matrix: the transposition of a thread-per-row result goes TMEM -> registers -> smem.
__device__ __forceinline__ void tmem_ldp1(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ldp2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ldp4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
// KB blocks of 8 words (= 4 packed columns) -> r[2 * KB]
template <int KB>
__device__ __forceinline__ void tmem_ld_packed(uint32_t taddr, uint32_t* r) {
  if constexpr (KB >= 4) { tmem_ldp4(taddr, r); if constexpr (KB > 4) tmem_ld_packed<KB - 4>(taddr + 16, r + 8); }
  else if constexpr (KB >= 2) { tmem_ldp2(taddr, r); if constexpr (KB > 2) tmem_ld_packed<KB - 2>(taddr + 8, r + 4); }
  else tmem_ldp1(taddr, r);
}
// TL (multiple of 8, <= 64) consecutive fp32 columns of this thread's lane
template <int TL>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
  if constexpr (TL >= 32) { tmem_ld32(taddr, v); if constexpr (TL > 32) tmem_ld_cols<TL - 32>(taddr + 32, v + 32); }
  else if constexpr (TL >= 16) { tmem_ld16(taddr, v); if constexpr (TL > 16) tmem_ld_cols<TL - 16>(taddr + 16, v + 16); }
  else tmem_ld8(taddr, v);
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// byte offset of element (row, col) inside a [rows x 64] 16-bit K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + ((col & 7) << 1));
}

}  // namespace tc
}  // namespace agb
