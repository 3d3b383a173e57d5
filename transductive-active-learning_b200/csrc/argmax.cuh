// Arg-max record algebra shared by argmax.cu and fused.cu (see argmax.cu for the reference path).
#pragma once
#include "common.cuh"

namespace tal {

struct Rec {
  double val;
  i64 idx;
  i64 cnt;
  int flags;
};

__device__ __forceinline__ Rec rec_empty() {
  Rec r;
  r.val = -INFINITY;
  r.idx = 0x7fffffffffffffffLL;
  r.cnt = 0;
  r.flags = 0;
  return r;
}

// max value; among equal values the lower index; counts add up.
__device__ __forceinline__ Rec rec_combine(const Rec& a, const Rec& b) {
  Rec r;
  r.flags = a.flags | b.flags;
  if (b.cnt == 0 || (a.cnt != 0 && a.val > b.val)) {
    r.val = a.val; r.idx = a.idx; r.cnt = a.cnt;
  } else if (a.cnt == 0 || b.val > a.val) {
    r.val = b.val; r.idx = b.idx; r.cnt = b.cnt;
  } else {
    r.val = a.val; r.idx = a.idx < b.idx ? a.idx : b.idx; r.cnt = a.cnt + b.cnt;
  }
  return r;
}

__device__ __forceinline__ Rec rec_shfl_xor(const Rec& a, int o) {
  Rec r;
  r.val = __shfl_xor_sync(0xffffffffu, a.val, o);
  r.idx = __shfl_xor_sync(0xffffffffu, a.idx, o);
  r.cnt = __shfl_xor_sync(0xffffffffu, a.cnt, o);
  r.flags = __shfl_xor_sync(0xffffffffu, a.flags, o);
  return r;
}

__device__ __forceinline__ Rec block_reduce_rec(Rec r, Rec* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = rec_combine(r, rec_shfl_xor(r, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = r;
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    r = lane < nw ? sh[lane] : rec_empty();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = rec_combine(r, rec_shfl_xor(r, o));
  }
  return r;  // valid in warp 0
}

__device__ __forceinline__ int status_of(int flags) {
  if (!(flags & 1)) return TAL_ARGMAX_EMPTY_SAFE_SET;  // marginal.py:336-337
  if (!(flags & 2)) return TAL_ARGMAX_ALL_NEG_INF;     // :338-339
  if (!(flags & 4)) return TAL_ARGMAX_ALL_NAN;         // :341-342
  return TAL_ARGMAX_OK;
}

// idx is -1 whenever the arg-max failed (status != OK): the update kernels treat a negative index
// as "no observation", so a failed selection can never be applied by a device-resident loop.
__device__ __forceinline__ void write_result(const Rec& r, tal_argmax_result* out) {
  const int status = status_of(r.flags);
  out->idx = (r.cnt && status == TAL_ARGMAX_OK) ? r.idx : -1;
  out->val = r.val;
  out->n_ties = r.cnt;
  out->flags = r.flags;
  out->status = status;
}

// per-candidate record of the masked objective (lib/algorithms/__init__.py:80, marginal.py:336-345)
__device__ __forceinline__ Rec rec_of(double f, bool is_safe, i64 global_idx) {
  int flags = 0;
  if (is_safe) flags |= 1;
  if (!(f == -INFINITY)) flags |= 2;  // NaN != -inf counts, like jnp.all(F == -inf)
  const double sf = is_safe ? f : -INFINITY;  // F.at[~safe_set].set(-inf)
  if (is_safe && sf == sf) flags |= 4;
  Rec e;
  e.flags = flags;
  e.val = sf;
  e.idx = global_idx;
  e.cnt = (sf == sf) ? 1 : 0;  // nanmax ignores NaN
  return e;
}

}  // namespace tal
