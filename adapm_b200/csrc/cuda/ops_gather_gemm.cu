// Fused parameter gather + tensor-core GEMM (sm_100a): the "Intent-prefetch all-gather (+) consumer GEMM"
// hot path of the north star. Scores C[M, N] = Q[M, K] x E[keys[0..N), 0..K)^T where the rows of E are
// NOT a dense matrix: they live in the parameter store, spread over the HBM of all GPUs. Producer warps
// resolve each key through the replicated directory, read the fp32 embedding rows straight from the local
// slab / local replica / a peer GPU over NVLink (16-byte loads on peer-mapped pointers), convert to bf16
// and write them into the 128-byte-swizzled K-major shared-memory layout that tcgen05.mma consumes - in the
// same kernel as the UTCHMMA tiles, overlapped with them through an mbarrier ring. The query tile arrives
// through TMA. Epilogues: STORE (fp32 scores) or RANK_COUNT (filtered-ranking evaluation of the KGE models:
// the [M x N] score matrix and the gathered entity matrix never exist in HBM).
//
// Warp roles (384 threads): 0 TMA(Q) | 1 MMA issuer | 2 TMEM alloc | 4-7 epilogue | 8-11 row gatherers.
#include <cuda_runtime.h>

#include "ops.h"
#include "pm_kernels.cuh"
#include "tcgen05_utils.cuh"

namespace adapm {
namespace cudaops {

namespace {

using namespace tc;

constexpr int kGatherThreads = 384;
constexpr int BK_ELEMS = 64;  // bf16 elements per 128-byte K-block row

enum GEpilogue : int { G_STORE = 0, G_RANK_COUNT = 1 };

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int EPI>
__global__ void __launch_bounds__(kGatherThreads, 1)
gather_gemm_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ Ctx c,
                   const Key* __restrict__ keys, int M, int N, int K, float* __restrict__ C, int ldc,
                   const float* __restrict__ true_score, const int* __restrict__ true_col, int* __restrict__ rank_out,
                   unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int num_kb = (K + BK_ELEMS - 1) / BK_ELEMS;
  dev::cta_enter(c);   // the gather warps read store rows through the directory: count this CTA in the grace epoch

  if (warp == 1 && lane == 0) {
    // full: 1 arrival (TMA expect_tx for the Q tile) + 4 arrivals (one per gather warp)
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full_bar[s], 5); mbar_init(&sm.empty_bar[s], 1); }
    mbar_init(&sm.tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===================== TMA producer: query tile =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&sm.empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&sm.full_bar[s], A_BYTES);
        tma_load_2d(sm.a[s], &tmap_q, &sm.full_bar[s], kb * BK_ELEMS, tile_m * BM);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(KIND_BF16, BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&sm.full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(sm.a[s]);
        const uint32_t b_addr = smem_u32(sm.b[s]);
#pragma unroll
        for (int k = 0; k < ROW_BYTES / MMA_K_BYTES; ++k)
          umma<KIND_BF16>(tmem_base, umma_desc(a_addr + k * MMA_K_BYTES), umma_desc(b_addr + k * MMA_K_BYTES), idesc,
                          (kb | k) != 0 ? 1u : 0u);
        umma_commit(&sm.empty_bar[s]);
      }
      umma_commit(&sm.tmem_full_bar);
    }
  } else if (warp >= 8) {
    // ===================== gatherers: store rows (local HBM / NVLink peers) -> swizzled smem =====================
    const int t = threadIdx.x - 256;          // tile row handled by this thread (0..127)
    const int n = tile_n * BN + t;
    const float* row = nullptr;
    unsigned nl = 0, nr = 0;
    if (n < N) {
      const Key key = keys[n];
      const int cls = class_of_key(c, key);
      for (int attempt = 0; attempt < 4096 && !row; ++attempt) {
        dev::Target tg = dev::resolve_fast(c, key, cls, &nl, &nr);
        row = tg.row;
        if (!row) __nanosleep(256);            // transitional slot (relocation in flight): retry
      }
    }
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&sm.empty_bar[s], ph ^ 1u);
      unsigned char* dst_row = sm.b[s] + (size_t)t * ROW_BYTES;
      const int k0 = kb * BK_ELEMS;
      // 64 floats of this row -> 8 chunks of 8 bf16; chunk c goes to position c ^ (t & 7)  (128B swizzle)
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = k0 + 4 * j;
        v[j] = (row && k < K) ? dev::ld_row4(row + k) : make_float4(0.f, 0.f, 0.f, 0.f);  // K % 4 == 0
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 o;
        o.x = pack_bf16(v[2 * ch].x, v[2 * ch].y);
        o.y = pack_bf16(v[2 * ch].z, v[2 * ch].w);
        o.z = pack_bf16(v[2 * ch + 1].x, v[2 * ch + 1].y);
        o.w = pack_bf16(v[2 * ch + 1].z, v[2 * ch + 1].w);
        *reinterpret_cast<uint4*>(dst_row + ((ch ^ (t & 7)) << 4)) = o;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.full_bar[s])) : "memory");
      }
    }
    if (stats) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { nl += __shfl_xor_sync(0xffffffffu, nl, o); nr += __shfl_xor_sync(0xffffffffu, nr, o); }
      if (lane == 0) {
        if (nl) atomicAdd(stats + 0, (unsigned long long)nl);
        if (nr) atomicAdd(stats + 1, (unsigned long long)nr);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    mbar_wait(&sm.tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = tile_m * BM + q * 32 + lane;
    float ts = 0.f;
    int tcol = -1, cnt = 0;
    if (EPI == G_RANK_COUNT && row < M) { ts = true_score[row]; tcol = true_col[row]; }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      const int col0 = tile_n * BN + c0;
      if (EPI == G_STORE) {
        store_chunk_coalesced(r, sm.epi[q], 1.f, C, ldc, tile_m * BM + q * 32, col0, M, N, lane);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          cnt += (col < N && col != tcol && __uint_as_float(r[j]) > ts) ? 1 : 0;
        }
      }
    }
    if (EPI == G_RANK_COUNT && row < M && cnt) atomicAdd(rank_out + row, cnt);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
  dev::cta_exit(c);
}

template <int EPI>
void launch_gather(CudaBackend& be, cudaStream_t stream, const void* Q, const Key* keys, int M, int N, int K, int ldq, float* C,
                   int ldc, const float* true_score, const int* true_col, int* rank_out, unsigned long long* stats) {
  ADAPM_CHECK(K % 4 == 0, "gather_gemm: the embedding length must be a multiple of 4");
  ADAPM_CHECK(ldq % 8 == 0 && ldq >= K, "gather_gemm: Q must be bf16 with a row pitch that is a multiple of 8 elements");
  ADAPM_CHECK(be.ctx().L.val_bytes == 4, "the fused ops need float32 rows (Options::dtype)");
  be.track_stream(stream);
  CUtensorMap mq = make_map(Q, M, ldq, BM, KIND_BF16);
  const size_t smem = sizeof(SmemLayout) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    ADAPM_CUDA_CHECK(cudaFuncSetAttribute(gather_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gather_gemm_kernel<EPI><<<grid, kGatherThreads, smem, stream>>>(mq, be.ctx(), keys, M, N, K, C, ldc, true_score, true_col,
                                                                  rank_out, stats);
  ADAPM_COUNT_LAUNCH();
  ADAPM_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

void gather_gemm(CudaBackend& be, cudaStream_t stream, const void* Q, const Key* keys, int M, int N, int K, int ldq, float* C,
                 int ldc, unsigned long long* stats) {
  if (M == 0 || N == 0) return;
  launch_gather<G_STORE>(be, stream, Q, keys, M, N, K, ldq, C, ldc, nullptr, nullptr, nullptr, stats);
}

void gather_gemm_rank_count(CudaBackend& be, cudaStream_t stream, const void* Q, const Key* keys, int M, int N, int K, int ldq,
                            const float* true_score, const int* true_col, int* rank_out, unsigned long long* stats) {
  if (M == 0 || N == 0) return;
  launch_gather<G_RANK_COUNT>(be, stream, Q, keys, M, N, K, ldq, nullptr, 0, true_score, true_col, rank_out, stats);
}

}  // namespace cudaops
}  // namespace adapm
