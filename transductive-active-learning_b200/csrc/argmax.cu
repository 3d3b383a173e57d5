// Safe-set-masked arg-max over the candidates.
//
// Reference path: DiscreteGreedyAlgorithm._step masks the objective by the
// sample region (lib/algorithms/__init__.py:80), then
// MarginalModel.objective_argmax (lib/model/marginal.py:330-348) raises on an
// empty safe set / all -inf / all NaN, masks by the operating safe set, takes
// nanmax and the set of exact ties (argwhere) and draws one uniformly.  Here
// one kernel does masking, the three checks, the max, the tie count and the
// lowest tied index with warp shuffles + a block reduction + a last-block
// pass over the per-block records.
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

__device__ __forceinline__ void write_result(const Rec& r, tal_argmax_result* out) {
  out->idx = r.cnt ? r.idx : -1;
  out->val = r.val;
  out->n_ties = r.cnt;
  out->flags = r.flags;
  out->status = status_of(r.flags);
}

constexpr int kArgmaxMaxBlocks = 2 * kNumSM;

struct ArgmaxScratch {
  unsigned int ticket;
  unsigned int pad[7];
  Rec part[kArgmaxMaxBlocks];
};

__global__ void __launch_bounds__(256)
masked_argmax_kernel(double* __restrict__ F, const unsigned char* __restrict__ safe,
                     const double* __restrict__ l, int q, int fc, i64 l_stride,
                     const unsigned char* __restrict__ sample, i64 N, i64 idx_offset,
                     ArgmaxScratch* __restrict__ scratch, tal_argmax_result* __restrict__ out) {
  __shared__ Rec sh[32];
  __shared__ bool is_last;
  Rec r = rec_empty();
  for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < N; x += (i64)gridDim.x * blockDim.x) {
    double f = F[x];
    if (sample && !sample[x]) {  // F.at[~get_mask(domain, sample_region)].set(-inf)
      f = -INFINITY;
      F[x] = f;
    }
    bool is_safe;
    if (safe) {
      is_safe = safe[x] != 0;
    } else {  // pessimistic safe set: all_{i >= fc} l_i[x] >= 0 (marginal.py:176-179)
      is_safe = true;
      for (int i = fc; i < q; ++i) is_safe = is_safe && (l[(i64)i * l_stride + x] >= 0.0);
    }
    int flags = 0;
    if (is_safe) flags |= 1;
    if (!(f == -INFINITY)) flags |= 2;  // NaN != -inf counts, like jnp.all(F == -inf)
    const double sf = is_safe ? f : -INFINITY;  // F.at[~safe_set].set(-inf)
    if (is_safe && sf == sf) flags |= 4;
    Rec e;
    e.flags = flags;
    e.val = sf;
    e.idx = x + idx_offset;
    e.cnt = (sf == sf) ? 1 : 0;  // nanmax ignores NaN
    r = rec_combine(r, e);
  }
  r = block_reduce_rec(r, sh);
  if (threadIdx.x == 0) {
    scratch->part[blockIdx.x] = r;
    __threadfence();
    const unsigned int t = atomicAdd(&scratch->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  Rec a = rec_empty();
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    const volatile Rec* p = &scratch->part[i];
    Rec b;
    b.val = p->val; b.idx = p->idx; b.cnt = p->cnt; b.flags = p->flags;
    a = rec_combine(a, b);
  }
  __syncthreads();
  a = block_reduce_rec(a, sh);
  if (threadIdx.x == 0) {
    write_result(a, out);
    scratch->ticket = 0;  // ready for the next call on this stream
  }
}

__global__ void argmax_combine_kernel(const tal_argmax_result* __restrict__ parts, int G,
                                      tal_argmax_result* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Rec a = rec_empty();
  for (int g = 0; g < G; ++g) {
    Rec b;
    b.val = parts[g].val;
    b.idx = parts[g].idx;
    b.cnt = parts[g].n_ties;
    b.flags = parts[g].flags;
    a = rec_combine(a, b);
  }
  write_result(a, out);
}

}  // namespace tal

using namespace tal;

extern "C" {

long long tal_argmax_scratch_bytes(void) { return (long long)sizeof(ArgmaxScratch); }

int tal_masked_argmax(double* F_io, const unsigned char* safe, const double* l, int q,
                      int first_constraint, long long l_stride, const unsigned char* sample,
                      long long N, long long idx_offset, void* scratch, tal_argmax_result* out,
                      void* stream) {
  TAL_ARG(F_io && scratch && out && N >= 0);
  TAL_ARG(safe || l || first_constraint >= q);
  TAL_ARG(q >= 0 && q <= TAL_MAX_Q);
  i64 blocks = (N + 1023) / 1024;
  if (blocks < 1) blocks = 1;
  if (blocks > kArgmaxMaxBlocks) blocks = kArgmaxMaxBlocks;
  masked_argmax_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      F_io, safe, l, q, first_constraint, l_stride, sample, N, idx_offset,
      (ArgmaxScratch*)scratch, out);
  TAL_LAUNCH_CHECK("tal_masked_argmax");
  return TAL_OK;
}

int tal_argmax_combine(const tal_argmax_result* parts, int G, tal_argmax_result* out,
                       void* stream) {
  TAL_ARG(parts && out && G >= 1);
  argmax_combine_kernel<<<1, 32, 0, as_stream(stream)>>>(parts, G, out);
  TAL_LAUNCH_CHECK("tal_argmax_combine");
  return TAL_OK;
}

}  // extern "C"
