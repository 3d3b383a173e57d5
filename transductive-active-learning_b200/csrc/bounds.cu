// Confidence-bound masks, ordered mask compaction and the scalar-wise subset
// test.
//
// Reference path: MarginalModel.optimistic_safe_set / pessimistic_safe_set /
// potential_expanders / max_l / potential_maximizers
// (lib/model/marginal.py:171-224), domain[mask] / set_to_idx
// (lib/utils.py:74-75) and is_subset_of (lib/utils.py:65-71).
#include "common.cuh"

namespace tal {

constexpr int kBoundsMaxBlocks = 2 * kNumSM;
struct BoundsScratch {
  unsigned int ticket;
  int has_nan[kBoundsMaxBlocks];
  double part[kBoundsMaxBlocks];
};
__device__ BoundsScratch g_bounds;

__device__ __forceinline__ void safe_sets(const double* __restrict__ l, const double* __restrict__ u,
                                          int q, int fc, i64 N, i64 x, bool& pess, bool& opt) {
  pess = true;
  opt = true;
  for (int i = fc; i < q; ++i) {  // jnp.all(... >= 0, axis=0); empty slice -> all true
    pess = pess && (l[(i64)i * N + x] >= 0.0);
    opt = opt && (u[(i64)i * N + x] >= 0.0);
  }
}

// pass 1: pess/opt masks and max_l = max l[0, operating_safe_set] (initial -inf)
__global__ void __launch_bounds__(256)
bounds_pass1_kernel(const double* __restrict__ l, const double* __restrict__ u, int q, int fc,
                    i64 N, const unsigned char* __restrict__ op_safe_in,
                    unsigned char* __restrict__ pess_out, unsigned char* __restrict__ opt_out,
                    double* __restrict__ max_l_out) {
  __shared__ double shv[32];
  __shared__ int shn[32];
  __shared__ bool is_last;
  double mx = -INFINITY;
  int nan = 0;
  for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < N; x += (i64)gridDim.x * blockDim.x) {
    bool pess, opt;
    safe_sets(l, u, q, fc, N, x, pess, opt);
    if (pess_out) pess_out[x] = pess;
    if (opt_out) opt_out[x] = opt;
    const bool op = op_safe_in ? (op_safe_in[x] != 0) : pess;
    if (op) {
      const double v = l[x];  // l[0, x]
      if (v != v) nan = 1;    // jnp.max propagates NaN
      else mx = fmax(mx, v);
    }
  }
  mx = warp_max(mx);
  nan = __any_sync(0xffffffffu, nan);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { shv[wid] = mx; shn[wid] = nan; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    mx = lane < nw ? shv[lane] : -INFINITY;
    nan = lane < nw ? shn[lane] : 0;
    mx = warp_max(mx);
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) {
      g_bounds.part[blockIdx.x] = mx;
      g_bounds.has_nan[blockIdx.x] = nan;
      __threadfence();
      const unsigned int t = atomicAdd(&g_bounds.ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (wid == 0) {
    mx = -INFINITY;
    nan = 0;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      mx = fmax(mx, *((volatile double*)&g_bounds.part[i]));
      nan |= *((volatile int*)&g_bounds.has_nan[i]);
    }
    mx = warp_max(mx);
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) {
      if (max_l_out) *max_l_out = nan ? __longlong_as_double(0x7ff8000000000000LL) : mx;
      g_bounds.ticket = 0;
    }
  }
}

// pass 2: potential maximizers / expanders (marginal.py:197-200, 221-224)
__global__ void __launch_bounds__(256)
bounds_pass2_kernel(const double* __restrict__ l, const double* __restrict__ u, int q, int fc,
                    i64 N, const double* __restrict__ max_l, unsigned char* __restrict__ pm,
                    unsigned char* __restrict__ pe) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= N) return;
  bool pess, opt;
  safe_sets(l, u, q, fc, N, x, pess, opt);
  if (pm) pm[x] = opt && (u[x] >= *max_l);
  if (pe) pe[x] = opt && !pess;
}

// ---- ordered compaction ----------------------------------------------------
constexpr int kCompactPerThread = 16;
constexpr int kCompactPerBlock = 256 * kCompactPerThread;
constexpr int kCompactMaxBlocks = 65536;
__device__ i64 g_compact_counts[kCompactMaxBlocks];

__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
  __shared__ int wsum[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int s = wsum[w];
    if (w < wid) base += s;
    tot += s;
  }
  total = tot;
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(256)
compact_count_kernel(const unsigned char* __restrict__ mask, i64 N) {
  const i64 base = (i64)blockIdx.x * kCompactPerBlock + (i64)threadIdx.x * kCompactPerThread;
  int c = 0;
#pragma unroll
  for (int j = 0; j < kCompactPerThread; ++j)
    if (base + j < N && mask[base + j]) ++c;
  int total;
  block_exclusive_scan(c, total);
  if (threadIdx.x == 0) g_compact_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) compact_scan_kernel(int nblocks, i64* __restrict__ count_out) {
  // single block: exclusive scan of g_compact_counts in place (sequential over tiles of 256)
  __shared__ i64 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < nblocks; t0 += 256) {
    const int i = t0 + threadIdx.x;
    const int v = i < nblocks ? (int)g_compact_counts[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, total);
    if (i < nblocks) g_compact_counts[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count_out = carry;
}

__global__ void __launch_bounds__(256)
compact_write_kernel(const unsigned char* __restrict__ mask, i64 N, i64* __restrict__ idx_out) {
  const i64 base = (i64)blockIdx.x * kCompactPerBlock + (i64)threadIdx.x * kCompactPerThread;
  unsigned int bits = 0;
  int c = 0;
#pragma unroll
  for (int j = 0; j < kCompactPerThread; ++j)
    if (base + j < N && mask[base + j]) { bits |= 1u << j; ++c; }
  int total;
  i64 o = g_compact_counts[blockIdx.x] + block_exclusive_scan(c, total);
#pragma unroll
  for (int j = 0; j < kCompactPerThread; ++j)
    if (bits & (1u << j)) idx_out[o++] = base + j;
}

// ---- scalar-wise isin ---------------------------------------------------------
__device__ int g_subset[2];  // {violations, |X|}

__global__ void subset_mark_kernel(const int* __restrict__ sid, int d, const i64* __restrict__ roi_idx,
                                   const i64* __restrict__ roi_count_dev, i64 roi_count_host,
                                   unsigned char* __restrict__ present) {
  const i64 m = roi_count_dev ? *roi_count_dev : roi_count_host;
  const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) { g_subset[0] = 0; g_subset[1] = 0; }
  if (e >= m * d) return;
  const i64 j = e / d;
  const int dd = (int)(e % d);
  present[sid[roi_idx[j] * d + dd]] = 1;
}

__global__ void subset_check_kernel(const int* __restrict__ sid, int d, i64 N,
                                    const unsigned char* __restrict__ x_mask,
                                    const unsigned char* __restrict__ present) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  int bad = 0, cnt = 0;
  if (x < N && x_mask[x]) {
    cnt = 1;
    for (int dd = 0; dd < d; ++dd)
      if (!present[sid[x * d + dd]]) bad = 1;
  }
  bad = __any_sync(0xffffffffu, bad);
  cnt = __any_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) {
    if (bad) atomicOr(&g_subset[0], 1);
    if (cnt) atomicOr(&g_subset[1], 1);
  }
}

__global__ void subset_final_kernel(const i64* __restrict__ roi_count_dev, i64 roi_count_host,
                                    int* __restrict__ out) {
  const i64 m = roi_count_dev ? *roi_count_dev : roi_count_host;
  int r;
  if (g_subset[1] == 0) r = (m == 0);  // utils.py:66-67
  else if (m == 0) r = 0;              // :68-69
  else r = (g_subset[0] == 0);         // :70-71
  *out = r;
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_bounds_masks(const double* l, const double* u, int q, int first_constraint, long long N,
                     const unsigned char* op_safe_in, unsigned char* pess, unsigned char* opt,
                     unsigned char* pm, unsigned char* pe, double* max_l_out, void* stream) {
  TAL_ARG(l && u && q >= 1 && q <= TAL_MAX_Q && N >= 1 && first_constraint >= 0);
  TAL_ARG(!(pm && !max_l_out));
  cudaStream_t st = as_stream(stream);
  i64 blocks = (N + 1023) / 1024;
  if (blocks > kBoundsMaxBlocks) blocks = kBoundsMaxBlocks;
  bounds_pass1_kernel<<<(unsigned)blocks, 256, 0, st>>>(l, u, q, first_constraint, N, op_safe_in,
                                                       pess, opt, max_l_out);
  TAL_LAUNCH_CHECK("tal_bounds_masks/pass1");
  if (pm || pe) {
    bounds_pass2_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(l, u, q, first_constraint, N,
                                                                    max_l_out, pm, pe);
    TAL_LAUNCH_CHECK("tal_bounds_masks/pass2");
  }
  return TAL_OK;
}

int tal_mask_to_indices(const unsigned char* mask, long long N, long long* idx_out,
                        long long* count_out, void* stream) {
  TAL_ARG(mask && idx_out && count_out && N >= 0);
  const i64 nblocks = (N + kCompactPerBlock - 1) / kCompactPerBlock;
  TAL_ARG(nblocks <= kCompactMaxBlocks);
  cudaStream_t st = as_stream(stream);
  if (nblocks > 0) {
    compact_count_kernel<<<(unsigned)nblocks, 256, 0, st>>>(mask, N);
    TAL_LAUNCH_CHECK("tal_mask_to_indices/count");
  }
  compact_scan_kernel<<<1, 256, 0, st>>>((int)nblocks, count_out);
  TAL_LAUNCH_CHECK("tal_mask_to_indices/scan");
  if (nblocks > 0) {
    compact_write_kernel<<<(unsigned)nblocks, 256, 0, st>>>(mask, N, idx_out);
    TAL_LAUNCH_CHECK("tal_mask_to_indices/write");
  }
  return TAL_OK;
}

int tal_scalar_subset(const int* sid, int d, long long N, long long n_sid,
                      const unsigned char* x_mask, const long long* roi_idx,
                      const long long* roi_count_dev, long long roi_count_host,
                      unsigned char* present, int* out, void* stream) {
  TAL_ARG(sid && x_mask && roi_idx && present && out && d >= 1 && N >= 1 && n_sid >= 1);
  cudaStream_t st = as_stream(stream);
  TAL_CUDA(cudaMemsetAsync(present, 0, (size_t)n_sid, st));
  // upper bound on the ROI size is N
  const i64 max_e = (roi_count_dev ? N : roi_count_host) * d;
  const i64 mark_blocks = max_e > 0 ? (max_e + 255) / 256 : 1;
  subset_mark_kernel<<<(unsigned)mark_blocks, 256, 0, st>>>(sid, d, roi_idx, roi_count_dev,
                                                           roi_count_host, present);
  TAL_LAUNCH_CHECK("tal_scalar_subset/mark");
  subset_check_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(sid, d, N, x_mask, present);
  TAL_LAUNCH_CHECK("tal_scalar_subset/check");
  subset_final_kernel<<<1, 1, 0, st>>>(roi_count_dev, roi_count_host, out);
  TAL_LAUNCH_CHECK("tal_scalar_subset/final");
  return TAL_OK;
}

}  // extern "C"
