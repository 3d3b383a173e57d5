// Incremental target-set conditioning for a FIXED region of interest, "append" form.
//
// Reference path: ITL._directed (lib/algorithms/itl.py:33-60) needs, at every step,
//   var_{x|A} = diag(Sigma_t - Sigma_t[:, A] (Sigma_t[A, A] + eps I)^-1 Sigma_t[A, :])
// i.e. the variances of the GP conditioned on the observations so far AND on the
// target set A.  Gaussian conditioning commutes, so that GP can be carried along
// next to the model's own: with W_0 = L^-1 Sigma_{t0}[A, :] from the set-up step
// (Cholesky + TRSM, csrc/condvar.cu),
//
//   Sigma^A_t = Sigma_{t0} - W_t^T W_t,      Sigma_{t0} = Sigma_t + H_t^T H_t
//
// where every observation (x, rho^2) appends ONE row to each factor:
//   h = k / sqrt(s),       k  = Sigma_t[x, :],                      s  = k[x]  + rho^2
//   g = kA / sqrt(sA),     kA = Sigma^A_t[x, :] = k + H^T H[:, x] - W^T W[:, x],  sA = kA[x] + rho^2
// and  var_{.|A} -= g^2,  i.e.  cs := var - var_{.|A}  follows  cs += g^2 - k w.
// The rows are stored interleaved in one buffer  Wb = [W_0 (m_pad rows); g_0; h_0; g_1; h_1; ...]
// and a step is ONE read-only pass over it (a GEMV with the column Wb[:, x] as
// weights, signs -,-,+,-,+,...) plus two appended rows: (m_pad + 2t) ncols 8 bytes of
// HBM traffic, no serial recurrence, nothing rewritten.  (The downdate form in
// condinc.cu rewrites all of V every step: 2 m_pad ncols 8 bytes.)
#include "common.cuh"

namespace tal {

__device__ __forceinline__ double row_sign(i64 j, i64 m_pad) {
  return (j < m_pad || ((j - m_pad) & 1) == 0) ? -1.0 : 1.0;  // W_0 and g rows: -, h rows: +
}

// scal2[0] = sA = k[x] + sum_j sign_j col_j^2 + rho^2,  scal2[1] = s = k[x] + rho^2   (one block)
template <typename T>
__global__ void __launch_bounds__(256)
cond_app_scalars_kernel(const double* __restrict__ col, i64 rows, i64 m_pad, const T* __restrict__ kbuf,
                        const double* __restrict__ rho, const i64* __restrict__ idx_dev, i64 idx_host,
                        double* __restrict__ scal2) {
  __shared__ double sh[33];
  double acc = 0.0;
  for (i64 j = threadIdx.x; j < rows; j += blockDim.x) acc += row_sign(j, m_pad) * col[j] * col[j];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    const i64 idx = idx_dev ? *idx_dev : idx_host;
    if (idx < 0) {  // failed arg-max (idx = -1): two zero rows are appended, cs stays (scal2[2] = 0)
      scal2[0] = 1.0; scal2[1] = 1.0; scal2[2] = 0.0;
      return;
    }
    const double r = rho[idx];
    const double kx = (double)kbuf[idx];
    scal2[0] = kx + acc + r * r;
    scal2[1] = (double)T(kx + r * r);  // as step_vectors_kernel
    scal2[2] = 1.0;
  }
}

// part[s][x] = sum_{j in split s} sign_j col_j Wb[j][x]; grid (ceil(ncols / 512), n_split), 2 columns
// per thread, 8 rows in flight.
__global__ void __launch_bounds__(256)
cond_app_partial_kernel(const double* __restrict__ Wb, i64 ldw, i64 rows, i64 m_pad,
                        const double* __restrict__ col, i64 ncols, double* __restrict__ part, int n_split) {
  const i64 x = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (x >= ncols) return;
  const i64 per = ((rows + n_split - 1) / n_split + 7) / 8 * 8;
  const i64 j0 = (i64)blockIdx.y * per;
  const i64 j1 = j0 + per < rows ? j0 + per : rows;
  const double2* Wp = reinterpret_cast<const double2*>(Wb + x);
  const i64 ld2 = ldw / 2;
  double a0 = 0.0, a1 = 0.0;
  i64 j = j0;
  for (; j + 8 <= j1; j += 8) {
    double2 v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = ld_stream(Wp + (j + t) * ld2);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const double c = row_sign(j + t, m_pad) * __ldg(col + j + t);
      a0 += c * v[t].x;
      a1 += c * v[t].y;
    }
  }
  for (; j < j1; ++j) {
    const double2 v = ld_stream(Wp + j * ld2);
    const double c = row_sign(j, m_pad) * __ldg(col + j);
    a0 += c * v.x;
    a1 += c * v.y;
  }
  *reinterpret_cast<double2*>(part + (i64)blockIdx.y * ncols + x) = make_double2(a0, a1);
}

// kA = k + sum_s part[s]; g = kA / sqrt(sA), h = k / sqrt(s); append both rows; cs += g^2 - k w.
template <typename T>
__global__ void __launch_bounds__(256)
cond_app_finish_kernel(double* __restrict__ Wb, i64 ldw, i64 rows, i64 ncols, i64 n_valid, i64 cand0,
                       const double* __restrict__ part, int n_split, const T* __restrict__ kbuf,
                       const T* __restrict__ wbuf, const double* __restrict__ scal2, double* __restrict__ cs) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= ncols) return;
  double g = 0.0, h = 0.0;
  if (x < n_valid && scal2[2] != 0.0) {
    const double k = (double)kbuf[cand0 + x], w = (double)wbuf[cand0 + x];
    double kA = k;
    for (int s = 0; s < n_split; ++s) kA += part[(i64)s * ncols + x];  // fixed order: deterministic
    g = kA / sqrt(scal2[0]);
    h = k / sqrt(scal2[1]);
    cs[x] = cs[x] + (g * g - k * w);
  }
  Wb[rows * ldw + x] = g;        // pad columns stay zero
  Wb[(rows + 1) * ldw + x] = h;
}

}  // namespace tal

using namespace tal;

extern "C" {

long long tal_cond_append_scratch_bytes(long long ncols, int n_split) {
  return ((long long)n_split * ncols + 8) * (long long)sizeof(double);
}

int tal_cond_append_update(double* Wb, long long ldw, long long cap_rows, long long m_pad, long long rows,
                           long long ncols, long long n_valid, long long cand0, const double* col,
                           const void* kbuf_i, const void* wbuf_i, const double* rho_i,
                           const long long* idx_dev, long long idx_host, double* cs, double* scratch,
                           int n_split, int dtype, void* stream) {
  TAL_ARG(Wb && col && kbuf_i && wbuf_i && rho_i && cs && scratch);
  TAL_ARG(m_pad >= 1 && rows >= m_pad && ((rows - m_pad) & 1) == 0 && rows + 2 <= cap_rows);
  TAL_ARG(ncols >= 2 && ncols % 2 == 0 && ldw >= ncols && ldw % 2 == 0 && n_valid <= ncols);
  TAL_ARG(n_split >= 1 && n_split <= 1024);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  cudaStream_t st = as_stream(stream);
  double* scal2 = scratch;
  double* part = scratch + 8;
  if (dtype == TAL_F64)
    cond_app_scalars_kernel<double><<<1, 256, 0, st>>>(col, rows, m_pad, (const double*)kbuf_i, rho_i,
                                                       idx_dev, idx_host, scal2);
  else
    cond_app_scalars_kernel<float><<<1, 256, 0, st>>>(col, rows, m_pad, (const float*)kbuf_i, rho_i,
                                                      idx_dev, idx_host, scal2);
  TAL_LAUNCH_CHECK("tal_cond_append_update/scalars");
  dim3 grid((unsigned)((ncols / 2 + 255) / 256), (unsigned)n_split);
  cond_app_partial_kernel<<<grid, 256, 0, st>>>(Wb, ldw, rows, m_pad, col, ncols, part, n_split);
  TAL_LAUNCH_CHECK("tal_cond_append_update/partial");
  const unsigned fg = (unsigned)((ncols + 255) / 256);
  if (dtype == TAL_F64)
    cond_app_finish_kernel<double><<<fg, 256, 0, st>>>(Wb, ldw, rows, ncols, n_valid, cand0, part, n_split,
                                                       (const double*)kbuf_i, (const double*)wbuf_i, scal2, cs);
  else
    cond_app_finish_kernel<float><<<fg, 256, 0, st>>>(Wb, ldw, rows, ncols, n_valid, cand0, part, n_split,
                                                      (const float*)kbuf_i, (const float*)wbuf_i, scal2, cs);
  TAL_LAUNCH_CHECK("tal_cond_append_update/finish");
  return TAL_OK;
}

}  // extern "C"
