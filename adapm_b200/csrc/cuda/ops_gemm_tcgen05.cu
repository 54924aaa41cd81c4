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
  // split-K (gridDim.z > 1, EPI_STORE only): this CTA accumulates k-blocks [kb0, kb0 + num_kb) and ADDS its partial tile
  // to C, which the host zeroed - for problems with few output tiles and a long K extent
  const int total_kb = (K + BK - 1) / BK;
  const int kb_per = (total_kb + (int)gridDim.z - 1) / (int)gridDim.z;
  const int kb0 = (int)blockIdx.z * kb_per;
  const int num_kb = min(kb_per, total_kb - kb0);
  if (num_kb <= 0) return;          // (uniform for the CTA: nothing was allocated yet)

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
        tma_load_2d(sm.a[s], &tmap_a, &sm.full_bar[s], (kb0 + kb) * BK, tile_m * BM);
        tma_load_2d(sm.b[s], &tmap_b, &sm.full_bar[s], (kb0 + kb) * BK, tile_n * BN);
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
        if (gridDim.z > 1) add_chunk_coalesced(r, sm.epi[q], alpha, C, ldc, tile_m * BM + q * 32, col0, M, N, lane);
        else store_chunk_coalesced(r, sm.epi[q], alpha, C, ldc, tile_m * BM + q * 32, col0, M, N, lane);
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

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 output tile with
// UMMA 256 x 256 x 16. Each CTA stages its own 128 rows of A and ONE HALF of the B tile (128 of the 256 columns'
// rows), so a k-block costs 32 KiB of shared memory per CTA instead of 48 KiB and every B byte is fetched once per
// pair: half the smem traffic per MMA and six pipeline stages. Each CTA keeps its 128 x 256 half of the fp32
// accumulator in its own TMEM, double-buffered (2 x 256 columns), and runs its own epilogue.
//   both CTAs  warp 0  TMA producer: cp.async.bulk.tensor ... cta_group::2, completion on the LEADER's full barrier
//   leader     warp 1  one thread issues tcgen05.mma.cta_group::2; tcgen05.commit ... multicast::cluster releases the
//                      smem stage / publishes the accumulator in BOTH CTAs
//   both CTAs  warp 2  tcgen05.alloc / dealloc cta_group::2
//   both CTAs  warps 4-7 epilogue; "accumulator drained" arrives on the leader's barrier (remote mbarrier arrive)
constexpr int CSTAGES = 6;
constexpr int CBN = 256;                     // N extent of the pair tile
constexpr int CB_HALF_BYTES = (CBN / 2) * ROW_BYTES;   // 16 KiB: this CTA's half of the B tile
struct CSmemLayout {
  alignas(1024) unsigned char a[CSTAGES][A_BYTES];
  alignas(1024) unsigned char b[CSTAGES][CB_HALF_BYTES];
  alignas(8) unsigned long long full_bar[CSTAGES];
  alignas(8) unsigned long long empty_bar[CSTAGES];
  alignas(8) unsigned long long tmem_full_bar[2];
  alignas(8) unsigned long long tmem_empty_bar[2];
  unsigned int tmem_base;
  float epi[4][32 * 33];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
// TMA load of this CTA's operand slice into its own shared memory; the bytes are signalled on `bar_cluster_addr`, a
// shared::cluster address that may belong to the peer CTA (the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == KIND_BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// completion of all MMAs issued so far arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(unsigned long long* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

template <int EPI, int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_nt_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                            int M, int N, int K, float* __restrict__ C, int ldc, float alpha,
                            const float* __restrict__ true_score, const int* __restrict__ true_col,
                            int* __restrict__ rank_out) {
  constexpr int BK = KIND == KIND_BF16 ? 64 : 128;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  CSmemLayout& sm = *reinterpret_cast<CSmemLayout*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();        // 0 = leader (issues the MMAs)
  const int num_kb = (K + BK - 1) / BK;
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + CBN - 1) / CBN;
  const int num_tiles = tiles_m * tiles_n;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CSTAGES; ++s) { mbar_init(&sm.full_bar[s], 1); mbar_init(&sm.empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tmem_full_bar[a], 1); mbar_init(&sm.tmem_empty_bar[a], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();     // barriers of BOTH CTAs are initialised before anybody signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int tile_m = tile % tiles_m, tile_n = tile / tiles_m;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % CSTAGES;
          mbar_wait(&sm.empty_bar[s], ((it / CSTAGES) & 1u) ^ 1u);       // my smem stage is free (multicast commit)
          const uint32_t leader_full = mapa_u32(smem_u32(&sm.full_bar[s]), 0);
          if (cta == 0) mbar_expect_tx(&sm.full_bar[s], 2 * (A_BYTES + CB_HALF_BYTES));   // both CTAs' bytes
          tma_load_2d_pair(sm.a[s], &tmap_a, leader_full, kb * BK, tile_m * 2 * BM + (int)cta * BM);
          tma_load_2d_pair(sm.b[s], &tmap_b, leader_full, kb * BK, tile_n * CBN + (int)cta * (CBN / 2));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (cta == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc(KIND, 2 * BM, CBN);
      uint32_t it = 0, t_local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++t_local) {
        const uint32_t acc = t_local & 1u;
        mbar_wait(&sm.tmem_empty_bar[acc], ((t_local >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained it
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * CBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % CSTAGES;
          mbar_wait(&sm.full_bar[s], (it / CSTAGES) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(sm.a[s]);
          const uint32_t b_addr = smem_u32(sm.b[s]);
#pragma unroll
          for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k)
            umma_pair<KIND>(d_tmem, umma_desc(a_addr + k * MMA_K_BYTES), umma_desc(b_addr + k * MMA_K_BYTES), idesc,
                            (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(&sm.empty_bar[s]);
        }
        umma_commit_pair(&sm.tmem_full_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: own 128 rows x 256 columns) =====================
    const int q = warp & 3;
    uint32_t t_local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++t_local) {
      const int tile_m = tile % tiles_m, tile_n = tile / tiles_m;
      const uint32_t acc = t_local & 1u;
      mbar_wait(&sm.tmem_full_bar[acc], (t_local >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row0 = tile_m * 2 * BM + (int)cta * BM + q * 32;
      const int row = row0 + lane;
      float ts = 0.f;
      int tc = -1, cnt = 0;
      if (EPI == EPI_RANK_COUNT && row < M) { ts = true_score[row]; tc = true_col[row]; }
#pragma unroll 1
      for (int c0 = 0; c0 < CBN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * CBN + (uint32_t)c0, r);
        const int col0 = tile_n * CBN + c0;
        if (EPI == EPI_STORE) {
          if (col0 < N) store_chunk_coalesced(r, sm.epi[q], alpha, C, ldc, row0, col0, M, N, lane);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            cnt += (col < N && col != tc && __uint_as_float(r[j]) > ts) ? 1 : 0;
          }
        }
      }
      if (EPI == EPI_RANK_COUNT && row < M && cnt) atomicAdd(rank_out + row, cnt);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {   // hand the accumulator back: the barrier lives in the leader CTA
        const uint32_t bar = mapa_u32(smem_u32(&sm.tmem_empty_bar[acc]), 0);
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();     // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
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
  // CTA pairs (cta_group::2, 256 x 256 tiles): ADAPM_GEMM_IMPL=c
  // (default for big problems: >= one 256 x 256 tile per SM pair and a K extent that amortises the pair set-up;
  //  measured 8192^3: 1417 vs 1313 TFLOP/s, 4096^3: 1315 vs 1167 - profiles/gemm_bench_v4.jsonl)
  static const int pair_env = [] { const char* e = getenv("ADAPM_GEMM_IMPL"); return e ? (e[0] == 'c' ? 1 : -1) : 0; }();
  const long tiles_pair = (long)((M + 2 * BM - 1) / (2 * BM)) * ((N + CBN - 1) / CBN);
  const bool pair_impl = pair_env == 1 || (pair_env == 0 && tiles_pair >= 64 && K >= 1024);
  if (pair_impl) {
    CUtensorMap mbh = make_map(B, N, K, CBN / 2, KIND);
    const size_t csmem = sizeof(CSmemLayout) + 1024;
    static bool cattr_set = false;
    static int num_sms_c = 0;
    if (!cattr_set) {
      ADAPM_CUDA_CHECK(cudaFuncSetAttribute(gemm_nt_tcgen05_pair_kernel<EPI, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
      int dev = 0;
      ADAPM_CUDA_CHECK(cudaGetDevice(&dev));
      ADAPM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_c, cudaDevAttrMultiProcessorCount, dev));
      cattr_set = true;
    }
    const long tiles_c = (long)((M + 2 * BM - 1) / (2 * BM)) * ((N + CBN - 1) / CBN);
    const int pairs = (int)std::min<long>(tiles_c, num_sms_c / 2);
    gemm_nt_tcgen05_pair_kernel<EPI, KIND><<<2 * pairs, kGemmThreads, csmem, stream>>>(ma, mbh, M, N, K, C, ldc, alpha,
                                                                                     true_score, true_col, rank_out);
    ADAPM_COUNT_LAUNCH();
    ADAPM_CUDA_CHECK(cudaGetLastError());
    return;
  }
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
  // split-K: few output tiles and a long K extent (e.g. the G^T E0 gradient GEMM of the shared-negative SGNS step:
  // 1024 x 304 x 32768 = 24 tiles) - spread the k-blocks of a tile over gridDim.z CTAs that add their partial tiles
  // into the zeroed C. ADAPM_GEMM_SPLITK=<n> forces n splits (1 = off).
  if (EPI == EPI_STORE) {
    static const int split_env = [] { const char* e = getenv("ADAPM_GEMM_SPLITK"); return e ? atoi(e) : 0; }();
    static int num_sms_t = 0;
    if (!num_sms_t) {
      int dev = 0;
      ADAPM_CUDA_CHECK(cudaGetDevice(&dev));
      ADAPM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_t, cudaDevAttrMultiProcessorCount, dev));
    }
    const int BKh = KIND == KIND_BF16 ? 64 : 128;
    const int total_kb = (K + BKh - 1) / BKh;
    const long tiles = (long)grid.x * grid.y;
    int splits = 1;
    if (split_env > 0) splits = split_env;
    else if (tiles * 2 <= num_sms_t && total_kb >= 32) splits = (int)std::min<long>(num_sms_t / tiles, total_kb / 8);
    splits = std::max(1, std::min(splits, std::min(total_kb, 64)));
    if (splits > 1) {
      const int kb_per = (total_kb + splits - 1) / splits;
      splits = (total_kb + kb_per - 1) / kb_per;          // no empty split
      ADAPM_CUDA_CHECK(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, stream));
      grid.z = (unsigned)splits;
    }
  }
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
