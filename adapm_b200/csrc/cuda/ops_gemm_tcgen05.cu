// Hand-written 5th-generation tensor-core GEMM for sm_100a:
//     C[M, N] (fp32) = A[M, K] (bf16, K-major) x B[N, K]^T (bf16, K-major)
// used for the 1-vs-all scoring of the KGE filtered-ranking evaluation (q(s,r) x E^T)
// and for the dense layers of the CTR DeepFM model. Two epilogues:
//   STORE       write the fp32 scores
//   RANK_COUNT  do not materialise the scores at all: per row count the candidates whose score
//               beats the row's true score (reference kge.cc:716-774 loops over all entities per
//               triple) and atomically accumulate the count -> the [B x ne] score matrix never
//               touches HBM.
//
// Structure (one 128 x 128 output tile per CTA, 256 threads):
//   warp 0   TMA producer      cp.async.bulk.tensor.2d (128B-swizzled 128x64 bf16 boxes) -> smem ring
//   warp 1   MMA issuer        one elected thread issues tcgen05.mma.cta_group::1.kind::f16
//                              (UMMA 128x128x16, A/B from smem descriptors, D in TMEM),
//                              tcgen05.commit releases smem stages / signals the epilogue
//   warp 2   TMEM allocator    tcgen05.alloc / dealloc (128 columns)
//   warps 4-7 epilogue         tcgen05.ld 32x32b.x32 (TMEM lanes = tile rows) -> registers -> C / counts
// SASS evidence: UTMALDG (TMA), UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld) - see profiles/.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ops.h"
#include "group.cuh"

namespace adapm {
namespace cudaops {

namespace {

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
};

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

enum Epilogue : int { EPI_STORE = 0, EPI_RANK_COUNT = 1 };

template <int EPI, int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nt_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M,
                       int N, int K, float* __restrict__ C, int ldc, float alpha, const float* __restrict__ true_score,
                       const int* __restrict__ true_col, int* __restrict__ rank_out) {
  constexpr int BK = KIND == KIND_BF16 ? 64 : 128;   // elements per 128-byte K-block row
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full_bar[s], 1); mbar_init(&sm.empty_bar[s], 1); }
    mbar_init(&sm.tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {  // TMEM allocation is warp-collective; the base address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&sm.empty_bar[s], ph ^ 1u);  // first pass over the ring passes immediately
        mbar_expect_tx(&sm.full_bar[s], A_BYTES + B_BYTES);
        tma_load_2d(sm.a[s], &tmap_a, &sm.full_bar[s], kb * BK, tile_m * BM);
        tma_load_2d(sm.b[s], &tmap_b, &sm.full_bar[s], kb * BK, tile_n * BN);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(KIND, BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&sm.full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(sm.a[s]);
        const uint32_t b_addr = smem_u32(sm.b[s]);
#pragma unroll
        for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k) {
          const uint64_t ad = umma_desc(a_addr + k * MMA_K_BYTES);
          const uint64_t bd = umma_desc(b_addr + k * MMA_K_BYTES);
          umma<KIND>(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&sm.empty_bar[s]);  // frees this smem stage once the MMAs above retire
      }
      umma_commit(&sm.tmem_full_bar);    // accumulator complete -> epilogue
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int q = warp & 3;  // this warp may touch TMEM lanes [32q, 32q+32)
    mbar_wait(&sm.tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tile_m * BM + q * 32 + lane;
    float ts = 0.f;
    int tc = -1;
    int cnt = 0;
    if (EPI == EPI_RANK_COUNT && row < M) { ts = true_score[row]; tc = true_col[row]; }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      const int col0 = tile_n * BN + c0;
      if (EPI == EPI_STORE) {
        if (row < M) {
          float* dst = C + (size_t)row * ldc + col0;
          if (col0 + 32 <= N && (((uintptr_t)dst) & 15u) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) =
                  make_float4(alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                              alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) dst[j] = alpha * __uint_as_float(r[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          cnt += (col < N && col != tc && __uint_as_float(r[j]) > ts) ? 1 : 0;
        }
      }
    }
    if (EPI == EPI_RANK_COUNT && row < M && cnt) atomicAdd(rank_out + row, cnt);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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
CUtensorMap make_map(const void* base, int64_t rows, int64_t cols, int box_rows, int kind) {
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

template <int EPI, int KIND>
void launch(cudaStream_t stream, const void* A, const void* B, int M, int N, int K, float* C, int ldc, float alpha,
            const float* true_score, const int* true_col, int* rank_out) {
  ADAPM_CHECK((K * (KIND == KIND_BF16 ? 2 : 1)) % 16 == 0, "gemm_nt: the K extent must be a multiple of 16 bytes (TMA row pitch)");
  ADAPM_CHECK((((uintptr_t)A) & 15u) == 0 && (((uintptr_t)B) & 15u) == 0, "gemm_nt: operands must be 16-byte aligned");
  CUtensorMap ma = make_map(A, M, K, BM, KIND);
  CUtensorMap mb = make_map(B, N, K, BN, KIND);
  const size_t smem = sizeof(SmemLayout) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    ADAPM_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_tcgen05_kernel<EPI, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_nt_tcgen05_kernel<EPI, KIND><<<grid, kGemmThreads, smem, stream>>>(ma, mb, M, N, K, C, ldc, alpha, true_score,
                                                                          true_col, rank_out);
  ADAPM_COUNT_LAUNCH();
  ADAPM_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void gemm_nt_bf16(cudaStream_t stream, const void* A, const void* B, int M, int N, int K, float* C, int ldc) {
  if (M == 0 || N == 0) return;
  launch<EPI_STORE, KIND_BF16>(stream, A, B, M, N, K, C, ldc, 1.f, nullptr, nullptr, nullptr);
}

// fp8 (e4m3) operands, fp32 accumulation, C = alpha * A B^T  (alpha = product of the per-tensor scales)
void gemm_nt_e4m3(cudaStream_t stream, const void* A, const void* B, int M, int N, int K, float* C, int ldc, float alpha) {
  if (M == 0 || N == 0) return;
  launch<EPI_STORE, KIND_E4M3>(stream, A, B, M, N, K, C, ldc, alpha, nullptr, nullptr, nullptr);
}

void gemm_nt_bf16_rank_count(cudaStream_t stream, const void* A, const void* B, int M, int N, int K,
                             const float* true_score, const int* true_col, int* rank_out) {
  if (M == 0 || N == 0) return;
  launch<EPI_RANK_COUNT, KIND_BF16>(stream, A, B, M, N, K, nullptr, 0, 1.f, true_score, true_col, rank_out);
}

}  // namespace cudaops
}  // namespace adapm
