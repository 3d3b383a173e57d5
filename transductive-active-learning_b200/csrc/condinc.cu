// Incremental target-set conditioning for a FIXED region of interest.
//
// Reference path: ITL._directed (lib/algorithms/itl.py:33-60) re-conditions on
// the ROI from scratch at every step: Cholesky of Sigma_AA + jitter^2 I and two
// solves against Sigma_A. (lib/gp/gaussian_distribution.py:16-37,
// lib/utils.py:10-24) — O(m^3 + m^2 N) flops per step.  When the ROI does not
// change between steps the same quantities follow from the previous step's by
// an exact rank-1 downdate.  With V = L^-1 Sigma_A. (m x N) and one observation
// at x* (k = Sigma[x*, :], s = k[x*] + rho^2, w = k / s as in rank1.cu):
//
//   Sigma_AA' + eps I = L L^T - u u^T,   u = k_A / sqrt(s)
//   Sigma_A.'         = Sigma_A. - k_A w^T
//   p = L^-1 u = V[:, x*] / sqrt(s),     L' = L T,  T = chol(I - p p^T)
//   V' = L'^-1 Sigma_A.' = T^-1 (V - V[:, x*] w^T)
//
// T has closed form (gamma_j = 1 - sum_{i<j} p_i^2):
//   T_jj = sqrt(gamma_{j+1}/gamma_j),  T_ij = -p_i p_j / sqrt(gamma_j gamma_{j+1})  (i > j)
// so T^-1 applied to a column is the linear recurrence
//   z_j = d_j (y_j + p_j sigma_j),  sigma_{j+1} = a_j sigma_j + b_j y_j
// with d_j = 1/T_jj, c_j = p_j / sqrt(gamma_j gamma_{j+1}), b_j = c_j d_j,
// a_j = 1 + b_j p_j.  One pass over V (read + write m*N*8 bytes, HBM-bound)
// also yields the new column sums of squares cs'[x] = sum_j z_j^2, i.e. the
// explained variance that directed ITL needs.  ||p||^2 <= var/(var+rho^2) < 1,
// so T is well conditioned even though Sigma_AA + 1e-10 I is not.
#include "common.cuh"

namespace tal {

struct alignas(32) CondCoef {
  double p, a, b, d;
};
// the same downdate written as a prefix sum (used when the rows are split over
// several CTAs): sigma_j = P_j / gamma_j with P_j = sum_{i<j} p_i y_i, because
// row j of T is -p_j p_{<j}^T T_{<j}^-T / T_jj and
// p_{<j}^T (I - p_{<j} p_{<j}^T)^-1 y_{<j} = (p_{<j}^T y_{<j}) / gamma_j.
struct alignas(32) CondCoef2 {
  double p, d, ginv, vx;
};

// pcol[j] = V[j][col] if this shard owns the column, else 0 (sum over shards = broadcast)
__global__ void __launch_bounds__(256)
cond_extract_col_kernel(const double* __restrict__ V, i64 ldv, i64 m_pad, i64 col0, i64 ncols,
                        const i64* __restrict__ idx_dev, i64 idx_host, double* __restrict__ pcol) {
  const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m_pad) return;
  const i64 c = (idx_dev ? *idx_dev : idx_host) - col0;
  pcol[j] = (c >= 0 && c < ncols) ? V[j * ldv + c] : 0.0;
}

// One block: p = pcol / sqrt(s); inclusive prefix sums of p^2; coefficients.
template <typename T>
__global__ void __launch_bounds__(1024)
cond_coeffs_kernel(const double* __restrict__ pcol, i64 m_pad, const T* __restrict__ kbuf,
                   const double* __restrict__ rho, const i64* __restrict__ idx_dev, i64 idx_host,
                   CondCoef* __restrict__ coef, CondCoef2* __restrict__ coef2) {
  __shared__ double wsum[32];
  __shared__ double carry_sh;
  const i64 idx = idx_dev ? *idx_dev : idx_host;
  const double r = idx >= 0 ? rho[idx] : 1.0;  // failed arg-max (idx = -1): pcol = 0, identity downdate
  const double s = idx >= 0 ? (double)T((double)kbuf[idx] + r * r) : 1.0;  // as in step_vectors_kernel
  const double rs = sqrt(s);
  if (threadIdx.x == 0) carry_sh = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (i64 j0 = 0; j0 < m_pad; j0 += blockDim.x) {
    const i64 j = j0 + threadIdx.x;
    const double p = j < m_pad ? pcol[j] / rs : 0.0;
    double inc = p * p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    double base = carry_sh;
    for (int w = 0; w < wid; ++w) base += wsum[w];
    const double incl = base + inc;       // sum_{i<=j} p_i^2
    const double g0 = 1.0 - (incl - p * p);  // gamma_j
    const double g1 = 1.0 - incl;            // gamma_{j+1}
    if (j < m_pad) {
      CondCoef c;
      c.p = p;
      c.d = sqrt(g0 / g1);
      const double cj = p / sqrt(g0 * g1);
      c.b = cj * c.d;
      c.a = 1.0 + c.b * p;
      coef[j] = c;
      CondCoef2 c2;
      c2.p = p;
      c2.d = c.d;
      c2.ginv = 1.0 / g0;
      c2.vx = pcol[j];
      coef2[j] = c2;
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_sh = incl;
    __syncthreads();
    (void)nw;
  }
}

// V <- T^-1 (V - V[:, x*] w^T), cs[x] = sum_j V'[j][x]^2.  One thread per pair of
// columns walks down the m rows (the recurrence is sequential in j but its
// loads are not): rows are read UNROLL at a time and the NEXT batch of loads is
// issued before the current batch is processed and stored, so that
// 2 * UNROLL * 16 bytes per thread are in flight (there are only N/2 threads).
template <typename T, int UNROLL>
__global__ void __launch_bounds__(64)
cond_update_kernel(double* __restrict__ V, i64 ldv, i64 m_pad, i64 ncols, i64 n_valid, i64 col0,
                   const CondCoef* __restrict__ coef, const double* __restrict__ pcol,
                   const T* __restrict__ wbuf, double* __restrict__ cs) {
  const i64 col = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (col >= ncols) return;
  // w for the (global) columns col0 + col, +1; columns past n_valid are padding
  const double w0 = col < n_valid ? (double)wbuf[col0 + col] : 0.0;
  const double w1 = col + 1 < n_valid ? (double)wbuf[col0 + col + 1] : 0.0;
  double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
  double2* Vp = reinterpret_cast<double2*>(V + col);
  const i64 ldv2 = ldv / 2;
  double2 y[UNROLL], yn[UNROLL];
#pragma unroll
  for (int t = 0; t < UNROLL; ++t) y[t] = ld_stream(Vp + (i64)t * ldv2);
  for (i64 j = 0; j < m_pad; j += UNROLL) {
    if (j + UNROLL < m_pad) {
#pragma unroll
      for (int t = 0; t < UNROLL; ++t) yn[t] = ld_stream(Vp + (j + UNROLL + t) * ldv2);
    }
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) {
      const CondCoef c = coef[j + t];
      const double vx = pcol[j + t];  // V[j][x*]
      const double y0 = y[t].x - vx * w0, y1 = y[t].y - vx * w1;
      const double z0 = c.d * (y0 + c.p * s0), z1 = c.d * (y1 + c.p * s1);
      s0 = c.a * s0 + c.b * y0;
      s1 = c.a * s1 + c.b * y1;
      q0 += z0 * z0;
      q1 += z1 * z1;
      st_stream(Vp + (j + t) * ldv2, make_double2(z0, z1));
    }
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) y[t] = yn[t];
  }
  cs[col] = q0;
  cs[col + 1] = q1;
}

// ---- row-parallel form: the m rows are split into chunks of CH rows --------
// pass 1: tot[c][x] = sum_{i in chunk c} p_i y_i[x],  y_i[x] = V[i][x] - vx_i w[x]
template <typename T>
__global__ void __launch_bounds__(128)
cond_chunk_totals_kernel(const double* __restrict__ V, i64 ldv, int CH, i64 ncols, i64 n_valid,
                         i64 col0, const CondCoef2* __restrict__ coef, const T* __restrict__ wbuf,
                         double* __restrict__ tot) {
  const i64 col = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (col >= ncols) return;
  const double w0 = col < n_valid ? (double)wbuf[col0 + col] : 0.0;
  const double w1 = col + 1 < n_valid ? (double)wbuf[col0 + col + 1] : 0.0;
  const i64 j0 = (i64)blockIdx.y * CH;
  const double2* Vp = reinterpret_cast<const double2*>(V + col);
  const i64 ldv2 = ldv / 2;
  double t0 = 0.0, t1 = 0.0;
  for (int j = 0; j < CH; j += 8) {
    double2 y[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) y[t] = ld_stream(Vp + (j0 + j + t) * ldv2);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const CondCoef2 c = coef[j0 + j + t];
      t0 += c.p * (y[t].x - c.vx * w0);
      t1 += c.p * (y[t].y - c.vx * w1);
    }
  }
  tot[(i64)blockIdx.y * ncols + col] = t0;
  tot[(i64)blockIdx.y * ncols + col + 1] = t1;
}

// pass 2: within chunk c start from P = sum_{c' < c} tot[c'][x] and apply
// z_j = d_j (y_j + p_j P_j / gamma_j), P_{j+1} = P_j + p_j y_j; csp[c][x] = sum z^2.
template <typename T>
__global__ void __launch_bounds__(128)
cond_chunk_apply_kernel(double* __restrict__ V, i64 ldv, int CH, i64 ncols, i64 n_valid, i64 col0,
                        const CondCoef2* __restrict__ coef, const T* __restrict__ wbuf,
                        const double* __restrict__ tot, double* __restrict__ csp) {
  const i64 col = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (col >= ncols) return;
  const double w0 = col < n_valid ? (double)wbuf[col0 + col] : 0.0;
  const double w1 = col + 1 < n_valid ? (double)wbuf[col0 + col + 1] : 0.0;
  const i64 j0 = (i64)blockIdx.y * CH;
  double P0 = 0.0, P1 = 0.0;
  for (int c = 0; c < (int)blockIdx.y; ++c) {  // fixed order: deterministic
    const double2 t = *reinterpret_cast<const double2*>(tot + (i64)c * ncols + col);
    P0 += t.x;
    P1 += t.y;
  }
  double2* Vp = reinterpret_cast<double2*>(V + col);
  const i64 ldv2 = ldv / 2;
  double q0 = 0.0, q1 = 0.0;
  double2 y[8], yn[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) y[t] = ld_stream(Vp + (j0 + t) * ldv2);
  for (int j = 0; j < CH; j += 8) {
    if (j + 8 < CH) {
#pragma unroll
      for (int t = 0; t < 8; ++t) yn[t] = ld_stream(Vp + (j0 + j + 8 + t) * ldv2);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const CondCoef2 c = coef[j0 + j + t];
      const double y0 = y[t].x - c.vx * w0, y1 = y[t].y - c.vx * w1;
      const double pg = c.p * c.ginv;
      const double z0 = c.d * (y0 + pg * P0), z1 = c.d * (y1 + pg * P1);
      P0 += c.p * y0;
      P1 += c.p * y1;
      q0 += z0 * z0;
      q1 += z1 * z1;
      st_stream(Vp + (j0 + j + t) * ldv2, make_double2(z0, z1));
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) y[t] = yn[t];
  }
  csp[(i64)blockIdx.y * ncols + col] = q0;
  csp[(i64)blockIdx.y * ncols + col + 1] = q1;
}

__global__ void __launch_bounds__(256)
cond_sum_chunks_kernel(const double* __restrict__ csp, int R, i64 ncols, double* __restrict__ cs) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= ncols) return;
  double acc = 0.0;
  for (int c = 0; c < R; ++c) acc += csp[(i64)c * ncols + x];
  cs[x] = acc;
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_cond_extract_col(const double* V, long long ldv, long long m_pad, long long col0,
                         long long ncols, const long long* idx_dev, long long idx_host,
                         double* pcol, void* stream) {
  TAL_ARG(V && pcol && m_pad >= 1 && ldv >= ncols);
  cond_extract_col_kernel<<<(unsigned)((m_pad + 255) / 256), 256, 0, as_stream(stream)>>>(
      V, ldv, m_pad, col0, ncols, idx_dev, idx_host, pcol);
  TAL_LAUNCH_CHECK("tal_cond_extract_col");
  return TAL_OK;
}

int tal_cond_update(double* V, long long ldv, long long m_pad, long long ncols, long long n_valid,
                    long long col0, const double* pcol, const void* kbuf_i, const void* wbuf_i,
                    const double* rho_i, const long long* idx_dev, long long idx_host,
                    void* coef_scratch, double* cs, int chunks, double* chunk_scratch, int dtype,
                    void* stream) {
  TAL_ARG(V && pcol && kbuf_i && wbuf_i && rho_i && coef_scratch && cs && n_valid <= ncols);
  TAL_ARG(m_pad >= 128 && m_pad % 128 == 0 && ncols >= 2 && ncols % 2 == 0 && ldv >= ncols && ldv % 2 == 0);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG(chunks >= 0 && (chunks <= 1 || chunk_scratch));
  cudaStream_t st = as_stream(stream);
  CondCoef* coef = (CondCoef*)coef_scratch;
  CondCoef2* coef2 = (CondCoef2*)(coef + m_pad);
  if (dtype == TAL_F64)
    cond_coeffs_kernel<double><<<1, 1024, 0, st>>>(pcol, m_pad, (const double*)kbuf_i, rho_i,
                                                   idx_dev, idx_host, coef, coef2);
  else
    cond_coeffs_kernel<float><<<1, 1024, 0, st>>>(pcol, m_pad, (const float*)kbuf_i, rho_i, idx_dev,
                                                  idx_host, coef, coef2);
  TAL_LAUNCH_CHECK("tal_cond_update/coeffs");
  if (chunks == 0) {  // auto (measured, profiles/): >= 64k columns fill the machine in one pass
    chunks = 1;
    if (ncols < 60000 && chunk_scratch)
      while (chunks < 32 && m_pad / (chunks * 2) >= 64 && (m_pad / (chunks * 2)) % 8 == 0) chunks *= 2;
  }
  if (chunks <= 1) {
    const unsigned grid = (unsigned)((ncols / 2 + 63) / 64);
    if (dtype == TAL_F64)
      cond_update_kernel<double, 8><<<grid, 64, 0, st>>>(V, ldv, m_pad, ncols, n_valid, col0, coef,
                                                         pcol, (const double*)wbuf_i, cs);
    else
      cond_update_kernel<float, 8><<<grid, 64, 0, st>>>(V, ldv, m_pad, ncols, n_valid, col0, coef,
                                                        pcol, (const float*)wbuf_i, cs);
    TAL_LAUNCH_CHECK("tal_cond_update");
    return TAL_OK;
  }
  TAL_ARG(m_pad % chunks == 0 && (m_pad / chunks) % 8 == 0);
  const int CH = (int)(m_pad / chunks);
  double* tot = chunk_scratch;
  double* csp = chunk_scratch + (i64)chunks * ncols;
  dim3 grid((unsigned)((ncols / 2 + 127) / 128), (unsigned)chunks);
  if (dtype == TAL_F64) {
    cond_chunk_totals_kernel<double><<<grid, 128, 0, st>>>(V, ldv, CH, ncols, n_valid, col0, coef2,
                                                          (const double*)wbuf_i, tot);
    TAL_LAUNCH_CHECK("tal_cond_update/totals");
    cond_chunk_apply_kernel<double><<<grid, 128, 0, st>>>(V, ldv, CH, ncols, n_valid, col0, coef2,
                                                         (const double*)wbuf_i, tot, csp);
  } else {
    cond_chunk_totals_kernel<float><<<grid, 128, 0, st>>>(V, ldv, CH, ncols, n_valid, col0, coef2,
                                                         (const float*)wbuf_i, tot);
    TAL_LAUNCH_CHECK("tal_cond_update/totals");
    cond_chunk_apply_kernel<float><<<grid, 128, 0, st>>>(V, ldv, CH, ncols, n_valid, col0, coef2,
                                                        (const float*)wbuf_i, tot, csp);
  }
  TAL_LAUNCH_CHECK("tal_cond_update/apply");
  cond_sum_chunks_kernel<<<(unsigned)((ncols + 255) / 256), 256, 0, st>>>(csp, chunks, ncols, cs);
  TAL_LAUNCH_CHECK("tal_cond_update/sum");
  return TAL_OK;
}

long long tal_cond_coef_bytes(long long m_pad) {
  return m_pad * (long long)(sizeof(CondCoef) + sizeof(CondCoef2));
}

long long tal_cond_chunk_scratch_bytes(long long ncols) { return 2LL * 32 * ncols * (long long)sizeof(double); }

}  // extern "C"
