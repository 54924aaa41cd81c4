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

#include <algorithm>
#include <cstdlib>

#include "ops.h"
#include "tcgen05_utils.cuh"

namespace adapm {
namespace cudaops {

namespace {

using namespace tc;

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
        store_chunk_coalesced(r, sm.epi[q], alpha, C, ldc, tile_m * BM + q * 32, col0, M, N, lane);
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

// ------------------------------------------------------------------------------------------------
// Persistent variant for large problems: one CTA per SM walks over 128 x 256 output tiles; the fp32
// accumulator is double-buffered in TMEM (2 x 256 columns = all 512) so that the epilogue of tile i
// (tcgen05.ld -> registers -> HBM) overlaps the MMAs of tile i+1; UMMA shape 128 x 256 x 16 halves the
// number of MMA instructions and of A-tile re-reads per output element.
constexpr int PBN = 256;
constexpr int PSTAGES = 4;
constexpr int PB_BYTES = PBN * ROW_BYTES;     // 32 KiB
struct PSmemLayout {
  alignas(1024) unsigned char a[PSTAGES][A_BYTES];
  alignas(1024) unsigned char b[PSTAGES][PB_BYTES];
  alignas(8) unsigned long long full_bar[PSTAGES];
  alignas(8) unsigned long long empty_bar[PSTAGES];
  alignas(8) unsigned long long tmem_full_bar[2];
  alignas(8) unsigned long long tmem_empty_bar[2];
  unsigned int tmem_base;
  float epi[4][32 * 33];
};

template <int EPI, int KIND>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_nt_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                                  int M, int N, int K, float* __restrict__ C, int ldc, float alpha,
                                  const float* __restrict__ true_score, const int* __restrict__ true_col,
                                  int* __restrict__ rank_out) {
  constexpr int BK = KIND == KIND_BF16 ? 64 : 128;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  PSmemLayout& sm = *reinterpret_cast<PSmemLayout*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (K + BK - 1) / BK;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + PBN - 1) / PBN;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(&sm.full_bar[s], 1); mbar_init(&sm.empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full_bar[a], 1); mbar_init(&sm.tmem_empty_bar[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512));
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
    if (lane == 0) {
      uint32_t it = 0;  // running k-block counter across tiles (stage = it % PSTAGES)
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tile_m = tile % tiles_m, tile_n = tile / tiles_m;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % PSTAGES;
          mbar_wait(&sm.empty_bar[s], ((it / PSTAGES) & 1u) ^ 1u);
          mbar_expect_tx(&sm.full_bar[s], A_BYTES + PB_BYTES);
          tma_load_2d(sm.a[s], &tmap_a, &sm.full_bar[s], kb * BK, tile_m * BM);
          tma_load_2d(sm.b[s], &tmap_b, &sm.full_bar[s], kb * BK, tile_n * PBN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(KIND, BM, PBN);
      uint32_t it = 0, t_local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_local) {
        const uint32_t acc = t_local & 1u;
        mbar_wait(&sm.tmem_empty_bar[acc], ((t_local >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * PBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % PSTAGES;
          mbar_wait(&sm.full_bar[s], (it / PSTAGES) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(sm.a[s]);
          const uint32_t b_addr = smem_u32(sm.b[s]);
#pragma unroll
          for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k)
            umma<KIND>(d_tmem, umma_desc(a_addr + k * MMA_K_BYTES), umma_desc(b_addr + k * MMA_K_BYTES), idesc,
                       (kb | k) != 0 ? 1u : 0u);
          umma_commit(&sm.empty_bar[s]);
        }
        umma_commit(&sm.tmem_full_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    uint32_t t_local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_local) {
      const int tile_m = tile % tiles_m, tile_n = tile / tiles_m;
      const uint32_t acc = t_local & 1u;
      mbar_wait(&sm.tmem_full_bar[acc], (t_local >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = tile_m * BM + q * 32 + lane;
      float ts = 0.f;
      int tc = -1, cnt = 0;
      if (EPI == EPI_RANK_COUNT && row < M) { ts = true_score[row]; tc = true_col[row]; }
#pragma unroll 1
      for (int c0 = 0; c0 < PBN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * PBN + (uint32_t)c0, r);
        const int col0 = tile_n * PBN + c0;
        if (EPI == EPI_STORE) {
          if (col0 < N) store_chunk_coalesced(r, sm.epi[q], alpha, C, ldc, tile_m * BM + q * 32, col0, M, N, lane);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            cnt += (col < N && col != tc && __uint_as_float(r[j]) > ts) ? 1 : 0;
          }
        }
      }
      if (EPI == EPI_RANK_COUNT && row < M && cnt) atomicAdd(rank_out + row, cnt);
      // hand the accumulator back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.tmem_empty_bar[acc])) : "memory");
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int EPI, int KIND>
void launch(cudaStream_t stream, const void* A, const void* B, int M, int N, int K, float* C, int ldc, float alpha,
            const float* true_score, const int* true_col, int* rank_out) {
  ADAPM_CHECK((K * (KIND == KIND_BF16 ? 2 : 1)) % 16 == 0, "gemm_nt: the K extent must be a multiple of 16 bytes (TMA row pitch)");
  ADAPM_CHECK((((uintptr_t)A) & 15u) == 0 && (((uintptr_t)B) & 15u) == 0, "gemm_nt: operands must be 16-byte aligned");
  CUtensorMap ma = make_map(A, M, K, BM, KIND);
  // large problems: persistent kernel with 128x256 tiles and a double-buffered TMEM accumulator
  static const int impl = [] { const char* e = getenv("ADAPM_GEMM_IMPL"); return e ? (e[0] == 'p' ? 2 : 1) : 0; }();
  const long tiles_p = (long)((M + BM - 1) / BM) * ((N + PBN - 1) / PBN);
  if (impl == 2 || (impl == 0 && tiles_p >= 128)) {
    CUtensorMap mbp = make_map(B, N, K, PBN, KIND);
    const size_t psmem = sizeof(PSmemLayout) + 1024;
    static bool pattr_set = false;
    static int num_sms = 0;
    if (!pattr_set) {
      ADAPM_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_tcgen05_persistent_kernel<EPI, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
      int dev = 0;
      ADAPM_CUDA_CHECK(cudaGetDevice(&dev));
      ADAPM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
      pattr_set = true;
    }
    const int grid = (int)std::min<long>(tiles_p, num_sms);
    gemm_nt_tcgen05_persistent_kernel<EPI, KIND><<<grid, kGemmThreads, psmem, stream>>>(ma, mbp, M, N, K, C, ldc, alpha,
                                                                                      true_score, true_col, rank_out);
    ADAPM_COUNT_LAUNCH();
    ADAPM_CUDA_CHECK(cudaGetLastError());
    return;
  }
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
