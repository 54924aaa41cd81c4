// Device-side building blocks of the parameter manager that the generic kernels
// (cuda_backend.cu) and the fused application kernels (ops_*.cu) share.
#pragma once
#include <cuda_runtime.h>
#include "group.cuh"

namespace adapm {
namespace dev {

// 16-byte row accesses. Rows may live in a peer GPU's HBM (NVLink) - the same
// instructions work on both; `relaxed.sys` keeps them coherent with concurrent REDs.
__device__ __forceinline__ float4 ld_row4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// weak 16-byte load through L1/L2 (local rows that are not re-read after a remote write in
// the same kernel)
__device__ __forceinline__ float4 ld_row4_weak(const float* p) {
  return __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void red_row4(float* p, float4 v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Warp-cooperative: where does `key` live for a read? Wraps protocol.h's locate_pull and
// returns raw pointers. All lanes get the same answer.
struct RowRef {
  const float* row;    // DIRECT: the row; SUM3: target row
  const float* base;   // SUM3 only
  const float* row2;   // SUM3 only
  const uint32_t* meta_ptr;
  uint32_t meta_val;
  int kind;            // LocKind
  bool local;
};

__device__ __forceinline__ RowRef locate_read(const Ctx& c, Key key) {
  WarpGroup g;
  PullLoc<float> l = locate_pull<float>(c, g, key, false);
  RowRef r;
  r.row = l.row; r.base = l.base; r.row2 = l.row2; r.meta_ptr = l.meta_ptr; r.meta_val = l.meta_val;
  r.kind = l.kind; r.local = l.local;
  return r;
}

// Fast local probe used by hot kernels: returns the row pointer if `key` is usable from local
// HBM (owned or valid replica), nullptr otherwise. One lane's worth of work (no shuffles).
__device__ __forceinline__ float* local_row_or_null(const Ctx& c, Key key, uint32_t* state_out, int32_t* slot_out) {
  int32_t s = __ldcg(slot_of(c, c.rank) + key);
  if (s < 0) return nullptr;
  uint32_t st = meta_state(mem::ld_acquire(meta_of(c, c.rank) + s));
  if (state_out) *state_out = st;
  if (slot_out) *slot_out = s;
  if (st == S_OWNED || st == S_REPLICA || st == S_INCOMING_REPLICA) return row_ptr<float>(c, c.rank, class_of_key(c, key), (uint32_t)s);
  return nullptr;
}

}  // namespace dev
}  // namespace adapm
