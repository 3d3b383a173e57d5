// Gram construction and nearest-index lookup.
//
// Reference path: Kernel.cross_covariance / covariance
// (lib/gp/kernels/base.py:17-25) over the scalar kernels of
// lib/gp/kernels/stationary.py:23-55 and non_stationary.py:11-13;
// get_indices (lib/utils.py:53-56).
//
// The reference materialises an n x m x d broadcast; here every thread keeps
// the coordinates of its own columns in registers, walks down a chunk of
// rows whose coordinates sit in shared memory, and writes 16-byte vectors
// with a fused distance -> exp epilogue.  Write-bound: n_rows*ld*sizeof(T).
#include "common.cuh"

namespace tal {

constexpr int kMaxD = 8;

template <int KIND>
__device__ __forceinline__ double kernel_value(const double* x, const double* y, int d,
                                               double variance, double lengthscale) {
  if constexpr (KIND == TAL_KERNEL_LINEAR) {
    double acc = 0.0;
    for (int t = 0; t < d; ++t) acc += (variance * x[t]) * y[t];  // (variance * x.T) @ y
    return acc;
  } else if constexpr (KIND == TAL_KERNEL_LAPLACE) {
    double n1 = 0.0;
    for (int t = 0; t < d; ++t) n1 += fabs(x[t] - y[t]);
    return variance * exp(-n1 / lengthscale);  // stationary.py:33-35
  } else {
    double ss = 0.0;
    for (int t = 0; t < d; ++t) {
      const double df = x[t] - y[t];
      ss += df * df;
    }
    const double nrm = sqrt(ss);  // jnp.linalg.norm(x - y, ord=2)
    if constexpr (KIND == TAL_KERNEL_GAUSSIAN) {
      const double z = nrm / lengthscale;
      return variance * exp(-0.5 * (z * z));  // stationary.py:25-28 (sqrt then square)
    } else if constexpr (KIND == TAL_KERNEL_MATERN32) {
      const double tau = nrm / lengthscale;
      const double s3 = sqrt(3.0);
      return variance * (1.0 + s3 * tau) * exp(-s3 * tau);  // :40-45
    } else {
      const double tau = nrm / lengthscale;
      const double s5 = sqrt(5.0);
      return variance * (1.0 + s5 * tau + 5.0 / 3.0 * (tau * tau)) * exp(-s5 * tau);  // :50-55
    }
  }
}

constexpr int kGramRows = 32;  // rows of Y per CTA

template <typename T, int KIND>
__global__ void __launch_bounds__(256)
gram_kernel(const double* __restrict__ X, i64 n, const double* __restrict__ Y, int d,
            double variance, double lengthscale, T* __restrict__ out, i64 ld, i64 j0, i64 n_rows, i64 sym_block) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  __shared__ double ys[kGramRows][kMaxD];
  const i64 jr0 = (i64)blockIdx.y * kGramRows;  // first local row of this CTA
  // lower-block storage: the rows of a CTA share one diagonal block (kGramRows == the block size, j0 aligned);
  // column tiles beyond its end are never read by anyone and are not written (half of the one-off build)
  if (sym_block > 0 && (i64)blockIdx.x * 256 * VN >= ((j0 + jr0) / sym_block + 1) * sym_block) return;
  const int nr = (int)min((i64)kGramRows, n_rows - jr0);
  for (int e = threadIdx.x; e < nr * d; e += blockDim.x)
    ys[e / d][e % d] = Y[(j0 + jr0 + e / d) * d + e % d];
  __syncthreads();
  const i64 col = ((i64)blockIdx.x * 256 + threadIdx.x) * VN;
  if (col >= ld) return;
  double xs[VN][kMaxD];
#pragma unroll
  for (int v = 0; v < VN; ++v)
    for (int t = 0; t < d; ++t) xs[v][t] = (col + v < n) ? X[(col + v) * d + t] : 0.0;
  for (int r = 0; r < nr; ++r) {
    T vals[VN];
#pragma unroll
    for (int v = 0; v < VN; ++v) {
      const double kv = kernel_value<KIND>(xs[v], ys[r], d, variance, lengthscale);
      vals[v] = (col + v < n) ? T(kv) : T(0);  // pad columns stay zero
    }
    V o;
    if constexpr (VN == 2) { o.x = vals[0]; o.y = vals[1]; }
    else { o.x = vals[0]; o.y = vals[1]; o.z = vals[2]; o.w = vals[3]; }
    st_stream(reinterpret_cast<V*>(out + (jr0 + r) * ld + col), o);
  }
}

template <typename T>
static void launch_gram(int kind, const double* X, i64 n, const double* Y, int d, double variance,
                        double lengthscale, T* out, i64 ld, i64 j0, i64 n_rows, cudaStream_t st, i64 sym_block = 0) {
  constexpr int VN = Vec16<T>::n;
  dim3 grid((unsigned)((ld + 256 * VN - 1) / (256 * VN)), (unsigned)((n_rows + kGramRows - 1) / kGramRows));
#define TAL_GRAM(K) \
  gram_kernel<T, K><<<grid, 256, 0, st>>>(X, n, Y, d, variance, lengthscale, out, ld, j0, n_rows, sym_block)
  switch (kind) {
    case TAL_KERNEL_GAUSSIAN: TAL_GRAM(TAL_KERNEL_GAUSSIAN); break;
    case TAL_KERNEL_LAPLACE: TAL_GRAM(TAL_KERNEL_LAPLACE); break;
    case TAL_KERNEL_MATERN32: TAL_GRAM(TAL_KERNEL_MATERN32); break;
    case TAL_KERNEL_MATERN52: TAL_GRAM(TAL_KERNEL_MATERN52); break;
    default: TAL_GRAM(TAL_KERNEL_LINEAR); break;
  }
#undef TAL_GRAM
}

// get_indices: one warp per query point A[k]; lanes stride over the domain,
// keep (distance, index) with the first minimum winning, shuffle-reduce.
__global__ void __launch_bounds__(256)
nearest_index_kernel(const double* __restrict__ X, i64 n, const double* __restrict__ A, i64 m, int d,
                     i64* __restrict__ out) {
  const i64 k = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= m) return;
  double a[kMaxD];
  for (int t = 0; t < d; ++t) a[t] = A[k * d + t];
  double best = INFINITY;
  i64 bi = 0x7fffffffffffffffLL;
  bool any_nan = false;
  for (i64 j = lane; j < n; j += 32) {
    double ss = 0.0;
    for (int t = 0; t < d; ++t) {
      const double df = X[j * d + t] - a[t];
      ss += df * df;
    }
    const double dist = sqrt(ss);  // jnp.linalg.norm(X[:, None] - A, axis=2)
    if (dist != dist) any_nan = true;
    if (dist < best) { best = dist; bi = j; }  // strict <: first minimum wins within the lane
  }
  (void)any_nan;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const i64 oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) out[k] = bi == 0x7fffffffffffffffLL ? 0 : bi;
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_gram_stationary(int kind, const double* X, long long n, const double* Y, long long m,
                        int d, double variance, double lengthscale, void* out, long long ld,
                        long long j0, long long n_rows, int dtype, void* stream) {
  TAL_ARG(X && Y && out && n >= 1 && m >= 1 && d >= 1 && d <= kMaxD);
  TAL_ARG(kind >= TAL_KERNEL_GAUSSIAN && kind <= TAL_KERNEL_LINEAR);
  TAL_ARG(ld >= n && ld % 4 == 0 && j0 >= 0 && n_rows >= 0 && j0 + n_rows <= m);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  if (n_rows == 0) return TAL_OK;
  TAL_ARG((n_rows + kGramRows - 1) / kGramRows <= 65535);
  if (dtype == TAL_F64)
    launch_gram<double>(kind, X, n, Y, d, variance, lengthscale, (double*)out, ld, j0, n_rows,
                        as_stream(stream));
  else
    launch_gram<float>(kind, X, n, Y, d, variance, lengthscale, (float*)out, ld, j0, n_rows,
                       as_stream(stream));
  TAL_LAUNCH_CHECK("tal_gram_stationary");
  return TAL_OK;
}

/* K(X, X) rows [j0, j0 + n_rows) in lower-block storage: only the column tiles that reach into the kept part
 * of a row block (columns below the end of its diagonal block) are computed and written. */
int tal_gram_stationary_lower(int kind, const double* X, long long n, int d, double variance, double lengthscale,
                              void* out, long long ld, long long j0, long long n_rows, int dtype, void* stream) {
  TAL_ARG(X && out && n >= 1 && d >= 1 && d <= kMaxD);
  TAL_ARG(kind >= TAL_KERNEL_GAUSSIAN && kind <= TAL_KERNEL_LINEAR);
  TAL_ARG(ld >= n && ld % 32 == 0 && j0 >= 0 && j0 % kGramRows == 0 && n_rows >= 0 && j0 + n_rows <= n);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  static_assert(kGramRows == 32, "a CTA's rows must share one diagonal block");
  if (n_rows == 0) return TAL_OK;
  TAL_ARG((n_rows + kGramRows - 1) / kGramRows <= 65535);
  const i64 blk = tal_sym_block(dtype);
  if (dtype == TAL_F64)
    launch_gram<double>(kind, X, n, X, d, variance, lengthscale, (double*)out, ld, j0, n_rows, as_stream(stream), blk);
  else
    launch_gram<float>(kind, X, n, X, d, variance, lengthscale, (float*)out, ld, j0, n_rows, as_stream(stream), blk);
  TAL_LAUNCH_CHECK("tal_gram_stationary_lower");
  return TAL_OK;
}

int tal_nearest_index(const double* X, long long n, const double* A, long long m, int d,
                      long long* out, void* stream) {
  TAL_ARG(X && A && out && n >= 1 && m >= 0 && d >= 1 && d <= kMaxD);
  if (m == 0) return TAL_OK;
  nearest_index_kernel<<<(unsigned)((m + 7) / 8), 256, 0, as_stream(stream)>>>(X, n, A, m, d, out);
  TAL_LAUNCH_CHECK("tal_nearest_index");
  return TAL_OK;
}

}  // extern "C"
