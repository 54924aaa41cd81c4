// tcgen05 / TMEM / TMA / mbarrier building blocks shared by the tensor-core kernels
// (ops_gemm_tcgen05.cu, ops_gather_gemm.cu). Raw PTX only - no CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "group.cuh"
#include "../adapm/log.h"

namespace adapm {
namespace cudaops {
namespace tc {

constexpr int BM = 128, BN = 128;
constexpr int ROW_BYTES = 128;                // one K-block row = 128 B = one swizzle-128B row
constexpr int MMA_K_BYTES = 32;               // one tcgen05.mma consumes 32 B of K per row (16 bf16 / 32 fp8)
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * ROW_BYTES;       // 16 KiB
constexpr int B_BYTES = BN * ROW_BYTES;       // 16 KiB
enum Kind : int { KIND_BF16 = 0, KIND_E4M3 = 1 };
constexpr int kGemmThreads = 256;
constexpr int TMEM_COLS = 128;                // fp32 accumulator: one column per output column

struct SmemLayout {
  alignas(1024) unsigned char a[STAGES][A_BYTES];
  alignas(1024) unsigned char b[STAGES][B_BYTES];
  alignas(8) unsigned long long full_bar[STAGES];
  alignas(8) unsigned long long empty_bar[STAGES];
  alignas(8) unsigned long long tmem_full_bar;
  unsigned int tmem_base;
  float epi[4][32 * 33];   // per epilogue warp: 32x32 transpose tile (padded) for coalesced C stores
};

// Epilogue helper: lane i holds 32 consecutive columns of tile row i (what tcgen05.ld 32x32b gives); write them
// through a padded shared-memory tile so that every global store instruction covers one 128-byte row segment.
__device__ __forceinline__ void store_chunk_coalesced(const uint32_t (&r)[32], float* tile, float alpha, float* C, int ldc,
                                                      int row0, int col0, int M, int N, int lane) {
#pragma unroll
  for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = alpha * __uint_as_float(r[j]);
  __syncwarp();
  const int col = col0 + lane;
#pragma unroll 4
  for (int rr = 0; rr < 32; ++rr) {
    const int row = row0 + rr;
    if (row < M && col < N) C[(size_t)row * ldc + col] = tile[rr * 33 + lane];
  }
  __syncwarp();
}

// Split-K variant: the partial tile is ADDED to C (which the host zeroed), one `red.global.add.f32` per element, same
// coalesced shape.
__device__ __forceinline__ void add_chunk_coalesced(const uint32_t (&r)[32], float* tile, float alpha, float* C, int ldc,
                                                    int row0, int col0, int M, int N, int lane) {
#pragma unroll
  for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = alpha * __uint_as_float(r[j]);
  __syncwarp();
  const int col = col0 + lane;
#pragma unroll 4
  for (int rr = 0; rr < 32; ++rr) {
    const int row = row0 + rr;
    if (row < M && col < N) atomicAdd(C + (size_t)row * ldc + col, tile[rr * 33 + lane]);
  }
  __syncwarp();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset       bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: D = fp32, both operands K-major, M x N tile.
//   kind::f16     a/b format 1 = bf16        kind::f8f6f4  a/b format 0 = e4m3
__device__ __forceinline__ constexpr uint32_t umma_idesc(int kind, int m, int n) {
  return (1u << 4) | ((kind == KIND_BF16 ? 1u : 0u) << 7) | ((kind == KIND_BF16 ? 1u : 0u) << 10) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == KIND_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    ADAPM_CHECK(e == cudaSuccess && qres == cudaDriverEntryPointSuccess && p, "cuTensorMapEncodeTiled is not available");
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// row-major [rows, cols] matrix of 2-byte (bf16) or 1-byte (e4m3) elements,
// box = [box_rows, 128 bytes of K], 128-byte swizzle
inline CUtensorMap make_map(const void* base, int64_t rows, int64_t cols, int box_rows, int kind) {
  CUtensorMap m;
  const int esz = kind == KIND_BF16 ? 2 : 1;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * esz};
  cuuint32_t box[2] = {(cuuint32_t)(ROW_BYTES / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, kind == KIND_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                           const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ADAPM_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " << (int)r);
  return m;
}


}  // namespace tc
}  // namespace cudaops
}  // namespace adapm
