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
#include "argmax.cuh"

namespace tal {

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
    r = rec_combine(r, rec_of(f, is_safe, x + idx_offset));
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
