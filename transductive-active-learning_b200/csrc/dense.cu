// Dense helpers around the Cholesky of a sub-block of a covariance (Thompson-sampling
// ROI, entropy metrics, information ratio).
//
// Reference path: GaussianDistribution.sample (lib/gp/gaussian_distribution.py:90-98: one SVD of the
// N x N covariance per draw), .entropy (:141-149: slogdet), .information_gain (:151-180: two slogdets of
// |B| x |B| blocks, one of them of the posterior given the target set), .sub_distr (:182-187),
// MarginalModel.masked_entropy (lib/model/marginal.py:119-127).  Here every one of them is
//   gather the block (this file) -> blocked DMMA Cholesky (condvar.cu: tal_potrf_lower)
// followed by a reduction over the factor's diagonal (log-determinant), a triangular matrix-vector
// pass (sampling: x = mean + L z, all k draws in ONE read of L), or a DMMA SYRK (condvar.cu:
// tal_syrk_lower_sub) for the Schur complement of the information gain.
#include "common.cuh"

namespace tal {

// S[j][k] = cov[a_j][a_k] + (j == k) (add + add_vec[j]^2); identity pad for j, k >= m.  idx == null: a_j = j.
// Lower-block storage: entry (a, b) beyond the end of a's diagonal block is read from its mirror (b, a).
// grid (m_pad, ceil(m_pad / 256)).
template <typename T>
__global__ void __launch_bounds__(256)
gather_sym_block_kernel(const T* __restrict__ cov, i64 ld, const i64* __restrict__ idx, i64 m, i64 m_pad,
                        const double* __restrict__ add_std, double add, double* __restrict__ S, i64 lds,
                        i64 sym_block) {
  const i64 j = blockIdx.x;
  const i64 k = (i64)blockIdx.y * blockDim.x + threadIdx.x;
  if (k >= m_pad) return;
  double v;
  if (j < m && k < m) {
    const i64 a = idx ? idx[j] : j, b = idx ? idx[k] : k;
    const bool mirrored = sym_block > 0 && b >= (a / sym_block + 1) * sym_block;
    v = mirrored ? (double)cov[b * ld + a] : (double)cov[a * ld + b];
    if (j == k) {
      v += add;
      if (add_std) v += add_std[j] * add_std[j];  // noise_covariance_matrix (lib/utils.py:44-50)
    }
  } else {
    v = (j == k) ? 1.0 : 0.0;
  }
  S[j * lds + k] = v;
}

// out[0] = 2 sum_{i < m} log L[i][i]  (one CTA; NaN / -inf propagate like slogdet of a singular matrix)
__global__ void __launch_bounds__(256) chol_logdet_kernel(const double* __restrict__ L, i64 lds, i64 m,
                                                          double* __restrict__ out) {
  __shared__ double sh[33];
  double acc = 0.0;
  for (i64 i = threadIdx.x; i < m; i += blockDim.x) acc += log(L[i * lds + i]);
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = 2.0 * acc;
}

// out[s][r] = mean[r] + sum_{c <= r} L[r][c] Z[s][c], s < KS: one warp per row, the row of L read once for
// all KS draws (HBM-bound on L: n^2 / 2 x 8 bytes).  grid ceil(n / 8), block 256.
template <int KS>
__global__ void __launch_bounds__(256)
tri_sample_kernel(const double* __restrict__ L, i64 lds, i64 n, const double* __restrict__ Z, i64 ldz, int k,
                  const double* __restrict__ mean, double* __restrict__ out, i64 ldo) {
  const int lane = threadIdx.x & 31;
  const i64 r = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n) return;
  double acc[KS];
#pragma unroll
  for (int s = 0; s < KS; ++s) acc[s] = 0.0;
  const double* Lr = L + r * lds;
  for (i64 c = lane; c <= r; c += 32) {
    const double l = Lr[c];
#pragma unroll
    for (int s = 0; s < KS; ++s)
      if (s < k) acc[s] = fma(l, Z[(i64)s * ldz + c], acc[s]);
  }
#pragma unroll
  for (int s = 0; s < KS; ++s) {
    const double v = warp_sum(acc[s]);
    if (lane == 0 && s < k) out[(i64)s * ldo + r] = (mean ? mean[r] : 0.0) + v;
  }
}

// Vt[i][j] = V[j][idx[i]] for i < nb, j < m_pad; zero rows for nb <= i < nb_pad.  32 x 32 tiles through
// shared memory: reads run along the rows of V where idx is contiguous, writes along the rows of Vt.
// grid (nb_pad / 32, m_pad / 32).
__global__ void __launch_bounds__(256)
gather_cols_t_kernel(const double* __restrict__ V, i64 ldv, const i64* __restrict__ idx, i64 nb,
                     double* __restrict__ Vt, i64 ldvt) {
  __shared__ double tile[32][33];
  const i64 i0 = (i64)blockIdx.x * 32, j0 = (i64)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const i64 col = (i0 + tx < nb) ? (idx ? idx[i0 + tx] : i0 + tx) : -1;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = col >= 0 ? V[(j0 + r) * ldv + col] : 0.0;
  __syncthreads();
  for (int r = ty; r < 32; r += 8) Vt[(i0 + r) * ldvt + j0 + tx] = tile[tx][r];
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_gather_sym_block(const void* cov, long long ld, long long N, const long long* idx, long long m,
                         long long m_pad, const double* add_std, double add, double* S, long long lds,
                         int lower, int dtype, void* stream) {
  TAL_ARG(cov && S && m >= 1 && m_pad >= m && lds >= m_pad && ld >= N && (idx || m <= N));
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG((m_pad + 255) / 256 <= 65535);
  const i64 blk = lower ? tal_sym_block(dtype) : 0;
  dim3 grid((unsigned)m_pad, (unsigned)((m_pad + 255) / 256));
  if (dtype == TAL_F64)
    gather_sym_block_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)cov, ld, idx, m, m_pad,
                                                                         add_std, add, S, lds, blk);
  else
    gather_sym_block_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)cov, ld, idx, m, m_pad,
                                                                        add_std, add, S, lds, blk);
  TAL_LAUNCH_CHECK("tal_gather_sym_block");
  return TAL_OK;
}

int tal_chol_logdet(const double* L, long long lds, long long m, double* out, void* stream) {
  TAL_ARG(L && out && m >= 0 && lds >= m);
  chol_logdet_kernel<<<1, 256, 0, as_stream(stream)>>>(L, lds, m, out);
  TAL_LAUNCH_CHECK("tal_chol_logdet");
  return TAL_OK;
}

int tal_tri_sample(const double* L, long long lds, long long n, const double* Z, long long ldz, int k,
                   const double* mean, double* out, long long ldo, void* stream) {
  TAL_ARG(L && Z && out && n >= 1 && lds >= n && ldz >= n && ldo >= n && k >= 1);
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)((n + 7) / 8);
  for (int s0 = 0; s0 < k; s0 += 8) {  // 8 draws per read of L
    const int ks = k - s0 < 8 ? k - s0 : 8;
    tri_sample_kernel<8><<<grid, 256, 0, st>>>(L, lds, n, Z + (i64)s0 * ldz, ldz, ks, mean, out + (i64)s0 * ldo, ldo);
    TAL_LAUNCH_CHECK("tal_tri_sample");
  }
  return TAL_OK;
}

int tal_gather_cols_t(const double* V, long long ldv, long long m_pad, const long long* idx, long long nb,
                      long long nb_pad, double* Vt, long long ldvt, void* stream) {
  TAL_ARG(V && Vt && m_pad >= 32 && m_pad % 32 == 0 && nb >= 1 && nb_pad >= nb && nb_pad % 32 == 0 && ldvt >= m_pad);
  TAL_ARG(m_pad / 32 <= 65535);
  dim3 grid((unsigned)(nb_pad / 32), (unsigned)(m_pad / 32));
  gather_cols_t_kernel<<<grid, 256, 0, as_stream(stream)>>>(V, ldv, idx, nb, Vt, ldvt);
  TAL_LAUNCH_CHECK("tal_gather_cols_t");
  return TAL_OK;
}

}  // extern "C"
