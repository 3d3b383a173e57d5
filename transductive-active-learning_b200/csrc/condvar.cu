// Target-set conditioning for directed ITL, in fp64 on the DMMA tensor path.
//
// Reference path: ITL._directed (lib/algorithms/itl.py:33-60) calls
// compute_posterior_covariance(cov_i, roi_idx, None)
// (lib/gp/gaussian_distribution.py:16-55): Sigma_AA + jitter^2 I
// (lib/utils.py:44-50), solve_linear_system = cholesky + two solves
// (lib/utils.py:10-24), then an N x N x m GEMM of which only the diagonal is
// used (itl.py:48).  Here: gather, blocked Cholesky, forward TRSM only, and a
// column sum of squares:  post_var[x] = var[x] - || L^-1 cov[A, x] ||^2.
//
// All dense contractions (SYRK trailing update, panel solve, TRSM) run on
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — the only fp64 tensor shape on
// sm_100a; tcgen05 has no f64 kind) with cp.async multi-stage staging.
// m and the column count are padded by the caller to multiples of the tiles.
#include "common.cuh"
#include <cstdlib>

namespace tal {

constexpr int NB = 128;  // Cholesky / TRSM block size (rows of a block row)

// --------------------------------------------------------------------------
// gather
// --------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_roi_B_kernel(const T* __restrict__ cov, i64 ld, i64 N, const i64* __restrict__ roi_idx,
                    i64 m, double* __restrict__ B, i64 ldb) {
  // grid.x over padded rows (up to N of them: the ROI can be the whole domain); pad rows / columns are zero
  const i64 j = blockIdx.x;
  const i64 col = (i64)blockIdx.y * blockDim.x + threadIdx.x;
  if (col >= ldb) return;
  double v = 0.0;
  if (j < m && col < N) v = (double)cov[roi_idx[j] * ld + col];
  B[j * ldb + col] = v;
}

// lower-block storage (one GPU holding all rows): entry (a, col) beyond the end of a's diagonal
// block is read from its mirror image (col, a).  One CTA per 32 x 32 tile of B: the kept part is read
// along the rows of the target points (coalesced); the mirrored part is a column gather -- one sector
// per entry whatever the order -- which is read with the lanes running over the target points (the 32
// sectors of a warp lie in ONE covariance row, instead of 32 rows 8 N bytes apart) and transposed
// through shared memory so that the writes stay coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
gather_roi_B_lower_kernel(const T* __restrict__ cov, i64 ld, i64 N, const i64* __restrict__ roi_idx,
                          i64 m, double* __restrict__ B, i64 ldb, i64 sym_block) {
  __shared__ double tile[32][33];
  __shared__ i64 a_sm[32];
  const i64 j0 = (i64)blockIdx.x * 32, col0 = (i64)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (threadIdx.x < 32) a_sm[threadIdx.x] = (j0 + threadIdx.x < m) ? roi_idx[j0 + threadIdx.x] : -1;
  __syncthreads();
  const i64 col = col0 + tx;
  int any_mirror = 0;
  for (int r = ty; r < 32; r += 8) {  // kept entries + padding, written directly
    const i64 a = a_sm[r];
    double v = 0.0;
    bool direct = true;
    if (a >= 0 && col < N) {
      if (col < (a / sym_block + 1) * sym_block) v = (double)cov[a * ld + col];
      else direct = false;
    }
    if (direct) { if (col < ldb) B[(j0 + r) * ldb + col] = v; }
    else any_mirror = 1;
  }
  if (!__syncthreads_or(any_mirror)) return;
  const i64 a = a_sm[tx];
  for (int r = ty; r < 32; r += 8) {  // lanes over the target points, one covariance row per warp
    const i64 c = col0 + r;
    if (a >= 0 && c < N && c >= (a / sym_block + 1) * sym_block) tile[r][tx] = (double)cov[c * ld + a];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const i64 aj = a_sm[r];
    if (aj >= 0 && col < N && col >= (aj / sym_block + 1) * sym_block) B[(j0 + r) * ldb + col] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(256)
gather_roi_S_kernel(const double* __restrict__ B, i64 ldb, const i64* __restrict__ roi_idx, i64 m,
                    i64 m_pad, double jitter, double* __restrict__ S, i64 lds) {
  const i64 j = blockIdx.x;
  const i64 k = (i64)blockIdx.y * blockDim.x + threadIdx.x;
  if (k >= m_pad) return;
  double v;
  if (j < m && k < m) {
    v = B[j * ldb + roi_idx[k]];
    if (j == k) v = v + jitter * jitter;  // identity(k) * square(default) (utils.py:48-50)
  } else {
    v = (j == k) ? 1.0 : 0.0;  // pad: blockdiag(S, I)
  }
  S[j * lds + k] = v;
}

// Row-sharded variants: this GPU holds covariance rows [row0, row0 + n_loc).
// B[j][x] = cov[A_j][row0 + x] is read through the symmetric entry
// cov_loc[x][A_j] (a transposing gather staged in shared memory so that both
// the scattered reads stay inside one row and the writes are coalesced);
// S gets the rows j whose A_j this GPU owns (others zero: a sum over the
// shards assembles S; the identity pad is contributed by the shard with row0 = 0).
template <typename T>
__global__ void __launch_bounds__(256)
gather_roi_cols_kernel(const T* __restrict__ cov, i64 ld, i64 n_loc, const i64* __restrict__ roi_idx,
                       i64 m, double* __restrict__ B, i64 ldb, i64 ncols) {
  __shared__ double tile[32][33];
  const i64 x0 = (i64)blockIdx.x * 32, j0 = (i64)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const i64 j = j0 + tx;
  const i64 a = j < m ? roi_idx[j] : -1;
  for (int r = ty; r < 32; r += 8) {
    const i64 x = x0 + r;
    double v = 0.0;
    if (a >= 0 && x < n_loc) v = (double)cov[x * ld + a];
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const i64 jj = j0 + r, x = x0 + tx;
    if (x < ncols) B[jj * ldb + x] = tile[tx][r];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_roi_S_rows_kernel(const T* __restrict__ cov, i64 ld, i64 row0, i64 n_loc,
                         const i64* __restrict__ roi_idx, i64 m, i64 m_pad, double jitter,
                         double* __restrict__ S, i64 lds) {
  const i64 j = blockIdx.x;
  const i64 k = (i64)blockIdx.y * blockDim.x + threadIdx.x;
  if (k >= m_pad) return;
  double v = 0.0;
  if (j < m && k < m) {
    const i64 r = roi_idx[j] - row0;
    if (r >= 0 && r < n_loc) {
      v = (double)cov[r * ld + roi_idx[k]];
      if (j == k) v = v + jitter * jitter;
    }
  } else if (j == k && row0 == 0) {
    v = 1.0;
  }
  S[j * lds + k] = v;
}

// --------------------------------------------------------------------------
// DMMA GEMM core:  acc[BM x BN] += A[BM x K] * B[K x BN]
//   A row-major (k contiguous).  B either [k][n] (NN) or [n][k] (NT).
// --------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int BM_, int BN_, int WM_, int WN_, int STAGES_>
struct GemmCfg {
  static constexpr int BM = BM_, BN = BN_, BK = 16, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int THREADS = WM * WN * 32;
  static constexpr int WTM = BM / WM, WTN = BN / WN;  // warp tile
  static constexpr int MF = WTM / 8, NF = WTN / 8;    // 8x8 fragments per warp tile
  static constexpr int SA = BK + 4;                   // A smem row stride (doubles), == 4 mod 16
  static constexpr int SB_NN = BN + 4;                // B [k][n] row stride, == 4 mod 16
  static constexpr int SB_NT = BK + 4;                // B [n][k] row stride
  static constexpr int A_ELEMS = BM * SA;
  static constexpr int B_ELEMS = (BK * SB_NN > BN * SB_NT) ? BK * SB_NN : BN * SB_NT;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double);
  static_assert(BN % 16 == 0 && BM % (WM * 8) == 0 && BN % (WN * 8) == 0, "tile shape");
};

template <class C, bool NT>
__device__ __forceinline__ void gemm_load_stage(double* st, const double* __restrict__ A, i64 lda,
                                                const double* __restrict__ Bm, i64 ldb, int k0) {
  // A tile: BM rows x 16 doubles = BM * 8 chunks of 16 B
  double* As = st;
  double* Bs = st + C::A_ELEMS;
  for (int c = threadIdx.x; c < C::BM * 8; c += C::THREADS) {
    const int r = c >> 3, ch = c & 7;
    cp_async16(As + r * C::SA + ch * 2, A + (i64)r * lda + k0 + ch * 2);
  }
  if constexpr (NT) {
    for (int c = threadIdx.x; c < C::BN * 8; c += C::THREADS) {
      const int r = c >> 3, ch = c & 7;
      cp_async16(Bs + r * C::SB_NT + ch * 2, Bm + (i64)r * ldb + k0 + ch * 2);
    }
  } else {
    constexpr int CPR = C::BN / 2;  // chunks per row
    for (int c = threadIdx.x; c < C::BK * CPR; c += C::THREADS) {
      const int r = c / CPR, ch = c % CPR;
      cp_async16(Bs + r * C::SB_NN + ch * 2, Bm + (i64)(k0 + r) * ldb + ch * 2);
    }
  }
}

// acc layout: acc[mf][nf][2] : row = wm*WTM + mf*8 + (lane>>2), col = wn*WTN + nf*8 + (lane&3)*2 + e
template <class C, bool NT>
__device__ __forceinline__ void gemm_accumulate(double* smem, const double* __restrict__ A, i64 lda,
                                                const double* __restrict__ Bm, i64 ldb, int K,
                                                double (&acc)[C::MF][C::NF][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / C::WN, wn = warp % C::WN;
  const int g = lane >> 2, t = lane & 3;
  const int nk = K / C::BK;
  if (nk == 0) return;
  // prologue
#pragma unroll
  for (int s = 0; s < C::STAGES - 1; ++s) {
    if (s < nk) gemm_load_stage<C, NT>(smem + s * C::STAGE_ELEMS, A, lda, Bm, ldb, s * C::BK);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<C::STAGES - 2>();
    __syncthreads();
    // prefetch tile kt + STAGES - 1 into the slot freed at iteration kt - 1
    const int nxt = kt + C::STAGES - 1;
    if (nxt < nk)
      gemm_load_stage<C, NT>(smem + (nxt % C::STAGES) * C::STAGE_ELEMS, A, lda, Bm, ldb, nxt * C::BK);
    cp_async_commit();
    const double* As = smem + (kt % C::STAGES) * C::STAGE_ELEMS;
    const double* Bs = As + C::A_ELEMS;
#pragma unroll
    for (int ks = 0; ks < C::BK / 4; ++ks) {
      double a[C::MF], b[C::NF];
#pragma unroll
      for (int mf = 0; mf < C::MF; ++mf)
        a[mf] = As[(wm * C::WTM + mf * 8 + g) * C::SA + ks * 4 + t];
#pragma unroll
      for (int nf = 0; nf < C::NF; ++nf) {
        if constexpr (NT) b[nf] = Bs[(wn * C::WTN + nf * 8 + g) * C::SB_NT + ks * 4 + t];
        else b[nf] = Bs[(ks * 4 + t) * C::SB_NN + wn * C::WTN + nf * 8 + g];
      }
#pragma unroll
      for (int mf = 0; mf < C::MF; ++mf)
#pragma unroll
        for (int nf = 0; nf < C::NF; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], a[mf], b[nf]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

template <class C>
__device__ __forceinline__ void acc_zero(double (&acc)[C::MF][C::NF][2]) {
#pragma unroll
  for (int mf = 0; mf < C::MF; ++mf)
#pragma unroll
    for (int nf = 0; nf < C::NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
}

// MODE 0: D = acc ; MODE 1: D = D - acc
template <class C, int MODE>
__device__ __forceinline__ void acc_store(double* __restrict__ D, i64 ldd,
                                          const double (&acc)[C::MF][C::NF][2]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / C::WN, wn = warp % C::WN;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mf = 0; mf < C::MF; ++mf)
#pragma unroll
    for (int nf = 0; nf < C::NF; ++nf) {
      double2* p = reinterpret_cast<double2*>(D + (i64)(wm * C::WTM + mf * 8 + g) * ldd +
                                              wn * C::WTN + nf * 8 + t * 2);
      double2 v;
      if constexpr (MODE == 1) {
        v = *p;
        v.x -= acc[mf][nf][0];
        v.y -= acc[mf][nf][1];
      } else {
        v.x = acc[mf][nf][0];
        v.y = acc[mf][nf][1];
      }
      *p = v;
    }
}

typedef GemmCfg<128, 64, 4, 2, 3> TrsmCfg;  // 256 threads, warp tile 32 x 32, 2 CTAs / SM

// --------------------------------------------------------------------------
// TRSM:  B <- L^-1 B.  One CTA per strip of BN columns; block row i:
//   T_i = B_i - sum_{j<i} L_ij X_j          (DMMA, K = i * NB)
//   X_i = inv(L_ii) T_i                     (DMMA, K = NB)
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(TrsmCfg::THREADS)
trsm_lower_kernel(const double* __restrict__ L, i64 lds, int m, const double* __restrict__ Linv,
                  double* __restrict__ B, i64 ldb) {
  typedef TrsmCfg C;
  extern __shared__ __align__(16) double smem[];
  double* Bs = B + (i64)blockIdx.x * C::BN;
  double acc[C::MF][C::NF][2];
  const int nblk = m / NB;
  for (int i = 0; i < nblk; ++i) {
    double* Bi = Bs + (i64)i * NB * ldb;
    if (i > 0) {
      acc_zero<C>(acc);
      gemm_accumulate<C, false>(smem, L + (i64)i * NB * lds, lds, Bs, ldb, i * NB, acc);
      acc_store<C, 1>(Bi, ldb, acc);
      __threadfence_block();
      __syncthreads();
    }
    acc_zero<C>(acc);
    gemm_accumulate<C, false>(smem, Linv + (i64)i * NB * NB, NB, Bi, ldb, NB, acc);
    acc_store<C, 0>(Bi, ldb, acc);
    __threadfence_block();
    __syncthreads();
  }
}

// --------------------------------------------------------------------------
// Cholesky, right-looking, block size NB.
// --------------------------------------------------------------------------
// (1) factor the diagonal block in shared memory, write L_jj and inv(L_jj).
//     512 threads = 4 lanes per row.  Left-looking (Crout): column c of the
//     UNscaled factor U[r][c] = L[r][c] * L[c][c] is a dot product over the
//     finished columns, U[r][c] = A[r][c] - sum_{k<c} U[r][k] U[c][k] g[k] with
//     g[k] = 1 / U[k][k]; dot products carry no stores, so the loads pipeline
//     (a right-looking update is latency-bound on shared memory here), the 4
//     lanes of a row are reduced by shuffles, and one barrier per column
//     suffices.  The inverse is then formed in place row by row:
//     X[r][c] = -(sum_{c<=k<r} L[r][k] X[k][c]) / L[r][r] only needs row r of L.
//     Row pitch 132 doubles keeps the 64-bit accesses of both phases conflict-free.
constexpr int DIAG_P = NB + 4;
constexpr int DIAG_THREADS = 4 * NB;
__global__ void __launch_bounds__(DIAG_THREADS)
potrf_diag_kernel(double* __restrict__ S, i64 lds, int j, double* __restrict__ Linv_all) {
  extern __shared__ __align__(16) double sm[];  // [NB][DIAG_P] + g[NB]
  constexpr int P = DIAG_P;
  double* g = sm + NB * P;
  double* D = S + (i64)j * NB * lds + (i64)j * NB;
  double* Linv = Linv_all + (i64)j * NB * NB;
  const int r = threadIdx.x >> 2, h = threadIdx.x & 3;
  for (int e = threadIdx.x; e < NB * NB; e += DIAG_THREADS) {
    const int rr = e >> 7, cc = e & (NB - 1);
    sm[rr * P + cc] = (cc <= rr) ? D[(i64)rr * lds + cc] : 0.0;
  }
  __syncthreads();
  for (int c = 0; c < NB; ++c) {
    double acc = 0.0;
    if (r >= c) {
      const double* ur = sm + r * P;
      const double* uc = sm + c * P;
      // four independent partial sums: the fp64 FMA chain is the critical path
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int k = h;
      for (; k + 12 < c; k += 16) {
        a0 += ur[k] * (uc[k] * g[k]);
        a1 += ur[k + 4] * (uc[k + 4] * g[k + 4]);
        a2 += ur[k + 8] * (uc[k + 8] * g[k + 8]);
        a3 += ur[k + 12] * (uc[k + 12] * g[k + 12]);
      }
      for (; k < c; k += 4) a0 += ur[k] * (uc[k] * g[k]);
      acc = (a0 + a1) + (a2 + a3);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (r >= c && h == 0) {
      const double u = sm[r * P + c] - acc;
      sm[r * P + c] = u;
      if (r == c) g[c] = 1.0 / u;
    }
    __syncthreads();
  }
  // scale to L: L[r][c] = U[r][c] / sqrt(U[c][c]); NaN when not positive definite
  for (int e = threadIdx.x; e < NB * NB; e += DIAG_THREADS) {
    const int rr = e >> 7, cc = e & (NB - 1);
    if (cc <= rr) {
      const double d = sqrt(sm[cc * P + cc]);
      const double v = (cc == rr) ? d : sm[rr * P + cc] / d;
      D[(i64)rr * lds + cc] = v;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NB * NB; e += DIAG_THREADS) {
    const int rr = e >> 7, cc = e & (NB - 1);
    if (cc < rr) sm[rr * P + cc] *= sqrt(g[cc]);  // 1/sqrt(U[cc][cc])
  }
  __syncthreads();
  if (threadIdx.x < NB) {
    const int t = threadIdx.x;
    const double d = sqrt(sm[t * P + t]);
    sm[t * P + t] = d;
    g[t] = 1.0 / d;  // reciprocal diagonal, off the serial path of the inverse
  }
  __syncthreads();
  // in-place inverse, row by row; here thread (c, h): column c, lane h
  const int c = r;
  for (int rr = 0; rr < NB; ++rr) {
    const double* lr = sm + rr * P;
    double acc = 0.0;
    if (c < rr) {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int k = c + h;
      for (; k + 12 < rr; k += 16) {
        a0 += lr[k] * sm[k * P + c];
        a1 += lr[k + 4] * sm[(k + 4) * P + c];
        a2 += lr[k + 8] * sm[(k + 8) * P + c];
        a3 += lr[k + 12] * sm[(k + 12) * P + c];
      }
      for (; k < rr; k += 4) a0 += lr[k] * sm[k * P + c];
      acc = (a0 + a1) + (a2 + a3);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    const double il = g[rr];
    __syncthreads();  // row rr of L has been read by everyone
    if (h == 0) {
      if (c < rr) sm[rr * P + c] = -acc * il;
      else if (c == rr) sm[rr * P + rr] = il;
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < NB * NB; e += DIAG_THREADS) {
    const int rr = e >> 7, cc = e & (NB - 1);
    Linv[e] = (cc <= rr) ? sm[rr * P + cc] : 0.0;
  }
}

// (1') The same diagonal-block step on the tensor path (the default; TAL_POTRF_DIAG=0 selects the scalar
//     kernel above, kept as a cross-check).  The scalar kernel spends 0.21 ms per block in 3 x 128 barrier-
//     separated dot-product steps on one SM -- 6.7 of the 8.5 ms of a 4,096-point Cholesky.  Here the block is
//     factored right-looking in 16-column panels inside one CTA of 8 warps:
//       (a) warp 0 factors the 16 x 16 diagonal block and inverts it, one row (then one column) per lane in
//           registers, operands exchanged by shuffles;
//       (b) panel  L_ip = S_ip inv(L_pp)^T  and (c) trailing update  S_ik -= L_ip L_kp^T  as 8 x 8 DMMA
//           fragments straight out of shared memory (pitch 132 doubles: conflict-free fragment loads);
//     and the 128 x 128 inverse is assembled by blocked forward substitution (a DMMA product and a
//     16-step solve per column for every 16-row block row).  The strictly lower part of the inverse lives
//     TRANSPOSED in the unused upper triangle of the block's shared-memory copy, its diagonal in a side array.
constexpr int D2_PA = NB + 4;   // pitch of the block copy
constexpr int D2_PT = 64 + 4;   // (the product scratch holds 64 x 68 doubles; used as 16 rows of pitch 116)
constexpr int D2_PD = 16 + 4;   // pitch of the dense 16 x 16 inverse of the current diagonal block
constexpr int D2_THREADS = 256;
constexpr int D2_SMEM_BYTES = (NB * D2_PA + 64 * D2_PT + NB + 16 * D2_PD) * (int)sizeof(double);

__device__ __forceinline__ double d2_inv_at(const double* A, const double* invd, int r, int c) {
  return r > c ? A[c * D2_PA + r] : (r == c ? invd[r] : 0.0);
}

// warp 0: Cholesky of the 16 x 16 block at (c0, c0) and its inverse
__device__ __forceinline__ void d2_factor16(double* A, double* invd, double* Dinv, int c0) {
  const int lane = threadIdx.x & 31;
  const int r = lane & 15;  // (lanes 16..31 mirror 0..15 and write nothing)
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = A[(c0 + r) * D2_PA + c0 + k];
  double myinv = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const double akk = __shfl_sync(0xffffffffu, a[k], k);
    const double d = sqrt(akk);  // NaN when the block is not positive definite (propagates)
    const double id = 1.0 / d;
    const double lrk = (r == k) ? d : a[k] * id;
    a[k] = lrk;
    if (r == k) myinv = id;
#pragma unroll
    for (int c = k + 1; c < 16; ++c) {
      const double lck = __shfl_sync(0xffffffffu, lrk, c);
      a[c] = fma(-lrk, lck, a[c]);  // (entries above the diagonal of a row collect unused values)
    }
  }
  // X = inv(L): lane c holds column c, x[i] = X[i][c] (zero above the diagonal)
  const int c = r;
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (i == c) ? myinv : 0.0;
#pragma unroll
  for (int rr = 1; rr < 16; ++rr) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < rr; ++k) {
      const double lrk = __shfl_sync(0xffffffffu, a[k], rr);  // L[rr][k]
      if (k & 1) s1 = fma(lrk, x[k], s1);
      else s0 = fma(lrk, x[k], s0);
    }
    const double irr = __shfl_sync(0xffffffffu, myinv, rr);
    if (rr > c) x[rr] = -(s0 + s1) * irr;
  }
  if (lane < 16) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k <= r) A[(c0 + r) * D2_PA + c0 + k] = a[k];
      else A[(c0 + c) * D2_PA + c0 + k] = x[k];  // X[k][c], k > c, transposed into the upper triangle
      Dinv[k * D2_PD + c] = x[k];
    }
    invd[c0 + r] = myinv;
  }
}

__global__ void __launch_bounds__(D2_THREADS)
potrf_diag_mma_kernel(double* __restrict__ S, i64 lds, int j, double* __restrict__ Linv_all) {
  extern __shared__ __align__(16) double sm[];
  constexpr int PA = D2_PA, PD = D2_PD;
  double* A = sm;
  double* T = A + NB * PA;
  double* invd = T + 64 * D2_PT;
  double* Dinv = invd + NB;
  double* D = S + (i64)j * NB * lds + (i64)j * NB;
  double* Linv = Linv_all + (i64)j * NB * NB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int e = threadIdx.x; e < NB * NB; e += D2_THREADS) {
    const int rr = e >> 7, cc = e & (NB - 1);
    A[rr * PA + cc] = (cc <= rr) ? D[(i64)rr * lds + cc] : 0.0;
  }
  __syncthreads();
  for (int c0 = 0; c0 < NB; c0 += 16) {
    if (warp == 0) d2_factor16(A, invd, Dinv, c0);
    __syncthreads();
    const int nfr = (NB - c0 - 16) / 8;  // 8-row fragments below the diagonal block
    // (b) panel: rows below times inv(L_pp)^T, in place (a warp owns the rows it rewrites)
    for (int f = warp; f < nfr; f += 8) {
      const int rr = c0 + 16 + 8 * f;
      double av[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) av[ks] = A[(rr + g) * PA + c0 + ks * 4 + t];
      double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nf = 0; nf < 2; ++nf)
          dmma884(acc[nf][0], acc[nf][1], av[ks], Dinv[(nf * 8 + g) * PD + ks * 4 + t]);
      __syncwarp();
#pragma unroll
      for (int nf = 0; nf < 2; ++nf) {
        A[(rr + g) * PA + c0 + nf * 8 + t * 2] = acc[nf][0];
        A[(rr + g) * PA + c0 + nf * 8 + t * 2 + 1] = acc[nf][1];
      }
    }
    __syncthreads();
    // (c) trailing update of the fragments on or below the diagonal
    for (int fi = warp; fi < nfr; fi += 8) {
      const int ri = c0 + 16 + 8 * fi;
      double av[4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) av[ks] = A[(ri + g) * PA + c0 + ks * 4 + t];
      for (int fk = 0; fk <= fi; ++fk) {
        const int rk = c0 + 16 + 8 * fk;
        double u0 = 0.0, u1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) dmma884(u0, u1, av[ks], A[(rk + g) * PA + c0 + ks * 4 + t]);
        double* q = A + (ri + g) * PA + rk + t * 2;
        q[0] -= u0;
        q[1] -= u1;
      }
    }
    __syncthreads();
  }
  // inverse by blocked forward substitution, one 16-row block row after the other:
  //   X[bi, <bi] = -inv(L_bb) (L[bi, <bi] X[<bi, <bi])
  // the product on DMMA fragments (X is lower triangular: k-steps below the column fragment are skipped),
  // the 16 x 16 solve per column in one thread.  Row-wise this is what scalar forward substitution does, so
  // L X - I stays at rounding level; that residual is what the panel (S_ij X^T) and the TRSM (X T_i) inherit.
  // (A block-doubling inverse, inv([[A,0],[B,C]]) = [[inv A,0],[-inv C B inv A, inv C]], was measured first: its
  // residual is amplified by |B inv A| -- 7.5e-10 on the RBF Gram blocks of BASELINE configs[0], enough to push a
  // pivot of the NEXT block below zero against the 1e-10 jitter.)
  constexpr int PR = 116;  // pitch of the 16-row product scratch (== 4 mod 16)
  for (int bi = 1; bi < NB / 16; ++bi) {
    const int r0 = bi * 16, ncf = r0 / 8;
    for (int idx = warp; idx < 2 * ncf; idx += 8) {
      const int mf = idx & 1, nf = idx >> 1;
      double u0 = 0.0, u1 = 0.0;
      for (int ks = 2 * nf; ks < r0 / 4; ++ks)
        dmma884(u0, u1, A[(r0 + mf * 8 + g) * PA + ks * 4 + t], d2_inv_at(A, invd, ks * 4 + t, nf * 8 + g));
      T[(mf * 8 + g) * PR + nf * 8 + t * 2] = u0;
      T[(mf * 8 + g) * PR + nf * 8 + t * 2 + 1] = u1;
    }
    __syncthreads();
    if ((int)threadIdx.x < r0) {
      const int n = threadIdx.x;
      double x[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        double sacc = -T[r * PR + n];
#pragma unroll
        for (int k = 0; k < r; ++k) sacc = fma(-A[(r0 + r) * PA + r0 + k], x[k], sacc);
        x[r] = sacc * invd[r0 + r];
      }
#pragma unroll
      for (int r = 0; r < 16; ++r) A[n * PA + r0 + r] = x[r];  // X[r0 + r][n], transposed
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < NB * NB; e += D2_THREADS) {
    const int rr = e >> 7, cc = e & (NB - 1);
    if (cc <= rr) D[(i64)rr * lds + cc] = A[rr * PA + cc];
    Linv[e] = d2_inv_at(A, invd, rr, cc);
  }
}

typedef GemmCfg<128, 128, 4, 2, 3> PanelCfg;  // one CTA owns a whole NB x NB block
typedef GemmCfg<128, 64, 4, 2, 3> SyrkCfg;

// (2) panel:  L_ij = S_ij inv(L_jj)^T  for block rows i > j, in place.
//     One CTA per block row reads all of S_ij before it writes it back.
__global__ void __launch_bounds__(PanelCfg::THREADS)
potrf_panel_kernel(double* __restrict__ S, i64 lds, int j, const double* __restrict__ Linv_all) {
  typedef PanelCfg C;
  extern __shared__ __align__(16) double smem[];
  const int i = j + 1 + blockIdx.x;
  const double* Linv = Linv_all + (i64)j * NB * NB;
  double* Sij = S + (i64)i * NB * lds + (i64)j * NB;
  double acc[C::MF][C::NF][2];
  acc_zero<C>(acc);
  // C[r][n] = sum_k S_ij[r][k] * Linv[n][k]
  gemm_accumulate<C, true>(smem, Sij, lds, Linv, NB, NB, acc);
  acc_store<C, 0>(Sij, lds, acc);
}

// (3) trailing update:  S_ik -= L_ij L_kj^T  for j < k <= i.
//     grid.x enumerates the (i, k) block pairs of the trailing lower triangle,
//     grid.y the BN-wide halves of a block.
__global__ void __launch_bounds__(SyrkCfg::THREADS)
potrf_syrk_kernel(double* __restrict__ S, i64 lds, int j) {
  typedef SyrkCfg C;
  extern __shared__ __align__(16) double smem[];
  // decode pair index p -> (a, b) with 0 <= b <= a: a = floor((sqrt(8p+1)-1)/2)
  const int p = blockIdx.x;
  int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= p) ++a;
  while (a * (a + 1) / 2 > p) --a;
  const int b = p - a * (a + 1) / 2;
  const int i = j + 1 + a, k = j + 1 + b;
  const int n0 = blockIdx.y * C::BN;
  const double* Lij = S + (i64)i * NB * lds + (i64)j * NB;
  const double* Lkj = S + ((i64)k * NB + n0) * lds + (i64)j * NB;
  double* Sik = S + (i64)i * NB * lds + (i64)k * NB + n0;
  double acc[C::MF][C::NF][2];
  acc_zero<C>(acc);
  gemm_accumulate<C, true>(smem, Lij, lds, Lkj, lds, NB, acc);
  acc_store<C, 1>(Sik, lds, acc);
}

// (4) C <- C - A A^T on the tiles on or below the diagonal (the Schur complement
//     Sigma_BB - V_B^T V_B of GaussianDistribution.information_gain,
//     lib/gp/gaussian_distribution.py:151-180, with A = V_B^T stored [n][K]).
__global__ void __launch_bounds__(SyrkCfg::THREADS)
syrk_sub_kernel(double* __restrict__ Cm, i64 ldc, const double* __restrict__ A, i64 lda, int K) {
  typedef SyrkCfg C;
  extern __shared__ __align__(16) double smem[];
  const int p = blockIdx.x;
  int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= p) ++a;
  while (a * (a + 1) / 2 > p) --a;
  const int b = p - a * (a + 1) / 2;  // block row a >= block column b
  const int n0 = blockIdx.y * C::BN;
  double acc[C::MF][C::NF][2];
  acc_zero<C>(acc);
  gemm_accumulate<C, true>(smem, A + (i64)a * NB * lda, lda, A + ((i64)b * NB + n0) * lda, lda, K, acc);
  acc_store<C, 1>(Cm + (i64)a * NB * ldc + (i64)b * NB + n0, ldc, acc);
}

// --------------------------------------------------------------------------
// out[x] = sum_j V[j][x]^2   (dense column sums of squares; HBM-bound read)
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsumsq_dense_kernel(const double* __restrict__ V, i64 ldv, i64 m, i64 n_cols,
                      double* __restrict__ part) {
  const i64 col = ((i64)blockIdx.x * 256 + threadIdx.x) * 2;
  if (col >= n_cols) return;
  const i64 per = (m + gridDim.y - 1) / gridDim.y;
  const i64 j0 = (i64)blockIdx.y * per, j1 = min(m, j0 + per);
  double a0 = 0.0, a1 = 0.0;
  for (i64 j = j0; j < j1; j += 4) {
    double2 c[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (j + t < j1) c[t] = ld_stream(reinterpret_cast<const double2*>(V + (j + t) * ldv + col));
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (j + t < j1) { a0 += c[t].x * c[t].x; a1 += c[t].y * c[t].y; }
  }
  part[(i64)blockIdx.y * n_cols + col] = a0;
  if (col + 1 < n_cols) part[(i64)blockIdx.y * n_cols + col + 1] = a1;
}

__global__ void __launch_bounds__(256)
postvar_kernel(const double* __restrict__ part, int n_split, i64 n_cols, double* __restrict__ out) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_cols) return;
  double acc = 0.0;
  for (int s = 0; s < n_split; ++s) acc += part[(i64)s * n_cols + x];
  out[x] = acc;
}

template <class C, class K>
static int set_smem(K kern, bool& done) {
  if (!done) {
    TAL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)C::SMEM_BYTES));
    done = true;
  }
  return TAL_OK;
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_gather_roi(const void* cov, long long ld, long long N, const long long* roi_idx,
                   long long m, long long m_pad, double jitter, double* S, long long lds,
                   double* B, long long ldb, int dtype, void* stream) {
  TAL_ARG(cov && roi_idx && S && B && m >= 1 && m_pad >= m && m_pad % NB == 0);
  TAL_ARG(lds >= m_pad && ldb >= N && ldb <= 65535 * 256);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  cudaStream_t st = as_stream(stream);
  dim3 gb((unsigned)m_pad, (unsigned)((ldb + 255) / 256));
  if (dtype == TAL_F64)
    gather_roi_B_kernel<double><<<gb, 256, 0, st>>>((const double*)cov, ld, N, roi_idx, m, B, ldb);
  else
    gather_roi_B_kernel<float><<<gb, 256, 0, st>>>((const float*)cov, ld, N, roi_idx, m, B, ldb);
  TAL_LAUNCH_CHECK("tal_gather_roi/B");
  dim3 gs((unsigned)m_pad, (unsigned)((m_pad + 255) / 256));
  gather_roi_S_kernel<<<gs, 256, 0, st>>>(B, ldb, roi_idx, m, m_pad, jitter, S, lds);
  TAL_LAUNCH_CHECK("tal_gather_roi/S");
  return TAL_OK;
}

int tal_gather_roi_lower(const void* cov, long long ld, long long N, const long long* roi_idx,
                         long long m, long long m_pad, double jitter, double* S, long long lds,
                         double* B, long long ldb, int dtype, void* stream) {
  TAL_ARG(cov && roi_idx && S && B && m >= 1 && m_pad >= m && m_pad % NB == 0);
  TAL_ARG(lds >= m_pad && ldb >= N && ldb <= 65535 * 32);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  cudaStream_t st = as_stream(stream);
  const i64 blk = tal_sym_block(dtype);
  dim3 gb((unsigned)(m_pad / 32), (unsigned)((ldb + 31) / 32));
  if (dtype == TAL_F64)
    gather_roi_B_lower_kernel<double><<<gb, 256, 0, st>>>((const double*)cov, ld, N, roi_idx, m, B, ldb, blk);
  else
    gather_roi_B_lower_kernel<float><<<gb, 256, 0, st>>>((const float*)cov, ld, N, roi_idx, m, B, ldb, blk);
  TAL_LAUNCH_CHECK("tal_gather_roi_lower/B");
  dim3 gs((unsigned)m_pad, (unsigned)((m_pad + 255) / 256));
  gather_roi_S_kernel<<<gs, 256, 0, st>>>(B, ldb, roi_idx, m, m_pad, jitter, S, lds);
  TAL_LAUNCH_CHECK("tal_gather_roi_lower/S");
  return TAL_OK;
}

int tal_gather_roi_sharded(const void* cov, long long ld, long long row0, long long n_loc,
                           const long long* roi_idx, long long m, long long m_pad, double jitter,
                           double* S, long long lds, double* B, long long ldb, long long ncols,
                           int dtype, void* stream) {
  TAL_ARG(cov && roi_idx && S && B && m >= 1 && m_pad >= m && m_pad % NB == 0);
  TAL_ARG(lds >= m_pad && ldb >= ncols && ncols >= n_loc && m_pad <= 65535 * 32 && n_loc >= 0);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  cudaStream_t st = as_stream(stream);
  dim3 gb((unsigned)((ncols + 31) / 32), (unsigned)(m_pad / 32));
  dim3 gs((unsigned)m_pad, (unsigned)((m_pad + 255) / 256));
  if (dtype == TAL_F64) {
    gather_roi_cols_kernel<double><<<gb, 256, 0, st>>>((const double*)cov, ld, n_loc, roi_idx, m, B, ldb, ncols);
    TAL_LAUNCH_CHECK("tal_gather_roi_sharded/B");
    gather_roi_S_rows_kernel<double><<<gs, 256, 0, st>>>((const double*)cov, ld, row0, n_loc, roi_idx, m,
                                                        m_pad, jitter, S, lds);
  } else {
    gather_roi_cols_kernel<float><<<gb, 256, 0, st>>>((const float*)cov, ld, n_loc, roi_idx, m, B, ldb, ncols);
    TAL_LAUNCH_CHECK("tal_gather_roi_sharded/B");
    gather_roi_S_rows_kernel<float><<<gs, 256, 0, st>>>((const float*)cov, ld, row0, n_loc, roi_idx, m,
                                                       m_pad, jitter, S, lds);
  }
  TAL_LAUNCH_CHECK("tal_gather_roi_sharded/S");
  return TAL_OK;
}

long long tal_potrf_work_bytes(long long m_pad) {
  return (m_pad / NB) * (long long)NB * NB * (long long)sizeof(double);
}

int tal_potrf_lower(double* S, long long lds, long long m_pad, void* work, void* stream) {
  TAL_ARG(S && work && m_pad >= NB && m_pad % NB == 0 && lds >= m_pad && lds % 2 == 0);
  cudaStream_t st = as_stream(stream);
  static bool a0 = false, a1 = false, a2 = false;
  static int diag_mma = -1;
  if (diag_mma < 0) {
    const char* e = getenv("TAL_POTRF_DIAG");
    diag_mma = (e && atoi(e) == 0) ? 0 : 1;
  }
  const int diag_smem = (NB * DIAG_P + NB) * (int)sizeof(double);
  if (!a0) {
    TAL_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  diag_smem));
    TAL_CUDA(cudaFuncSetAttribute(potrf_diag_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  D2_SMEM_BYTES));
    a0 = true;
  }
  if (set_smem<PanelCfg>(potrf_panel_kernel, a1) != TAL_OK) return TAL_ERR_CUDA;
  if (set_smem<SyrkCfg>(potrf_syrk_kernel, a2) != TAL_OK) return TAL_ERR_CUDA;
  const int nblk = (int)(m_pad / NB);
  double* Linv = (double*)work;
  for (int j = 0; j < nblk; ++j) {
    if (diag_mma) potrf_diag_mma_kernel<<<1, D2_THREADS, D2_SMEM_BYTES, st>>>(S, lds, j, Linv);
    else potrf_diag_kernel<<<1, DIAG_THREADS, diag_smem, st>>>(S, lds, j, Linv);
    TAL_LAUNCH_CHECK("tal_potrf_lower/diag");
    const int rem = nblk - j - 1;
    if (rem > 0) {
      potrf_panel_kernel<<<rem, PanelCfg::THREADS, PanelCfg::SMEM_BYTES, st>>>(S, lds, j, Linv);
      TAL_LAUNCH_CHECK("tal_potrf_lower/panel");
      dim3 g((unsigned)(rem * (rem + 1) / 2), NB / SyrkCfg::BN);
      potrf_syrk_kernel<<<g, SyrkCfg::THREADS, SyrkCfg::SMEM_BYTES, st>>>(S, lds, j);
      TAL_LAUNCH_CHECK("tal_potrf_lower/syrk");
    }
  }
  return TAL_OK;
}

int tal_trsm_lower(const double* L, long long lds, long long m_pad, const void* work, double* B,
                   long long ldb, long long n_cols, void* stream) {
  TAL_ARG(L && work && B && m_pad >= NB && m_pad % NB == 0 && lds >= m_pad && lds % 2 == 0);
  TAL_ARG(n_cols >= 1 && n_cols % TrsmCfg::BN == 0 && ldb >= n_cols && ldb % 2 == 0);
  static bool a = false;
  if (set_smem<TrsmCfg>(trsm_lower_kernel, a) != TAL_OK) return TAL_ERR_CUDA;
  trsm_lower_kernel<<<(unsigned)(n_cols / TrsmCfg::BN), TrsmCfg::THREADS, TrsmCfg::SMEM_BYTES,
                      as_stream(stream)>>>(L, lds, (int)m_pad, (const double*)work, B, ldb);
  TAL_LAUNCH_CHECK("tal_trsm_lower");
  return TAL_OK;
}

int tal_syrk_lower_sub(double* Cm, long long ldc, long long n, const double* A, long long lda, long long K,
                       void* stream) {
  TAL_ARG(Cm && A && n >= NB && n % NB == 0 && K >= 16 && K % 16 == 0 && ldc >= n && lda >= K);
  TAL_ARG(ldc % 2 == 0 && lda % 2 == 0);
  static bool a = false;
  if (set_smem<SyrkCfg>(syrk_sub_kernel, a) != TAL_OK) return TAL_ERR_CUDA;
  const long long nb = n / NB;
  dim3 g((unsigned)(nb * (nb + 1) / 2), NB / SyrkCfg::BN);
  syrk_sub_kernel<<<g, SyrkCfg::THREADS, SyrkCfg::SMEM_BYTES, as_stream(stream)>>>(Cm, ldc, A, lda, (int)K);
  TAL_LAUNCH_CHECK("tal_syrk_lower_sub");
  return TAL_OK;
}

int tal_colsumsq_dense(const double* V, long long ldv, long long m, long long n_cols,
                       double* part, int n_split, double* out, void* stream) {
  TAL_ARG(V && part && out && m >= 1 && n_cols >= 1 && ldv >= n_cols && ldv % 2 == 0);
  TAL_ARG(n_split >= 1 && n_split <= 65535);
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)((n_cols + 511) / 512), (unsigned)n_split);
  colsumsq_dense_kernel<<<grid, 256, 0, st>>>(V, ldv, m, n_cols, n_split == 1 ? out : part);
  TAL_LAUNCH_CHECK("tal_colsumsq_dense");
  if (n_split > 1) {
    postvar_kernel<<<(unsigned)((n_cols + 255) / 256), 256, 0, st>>>(part, n_split, n_cols, out);
    TAL_LAUNCH_CHECK("tal_colsumsq_dense/sum");
  }
  return TAL_OK;
}

}  // extern "C"

// --------------------------------------------------------------------------
// Peak probes (roofline denominators for the fp64 tensor path)
// --------------------------------------------------------------------------
namespace tal {
__global__ void __launch_bounds__(256) probe_dmma_kernel(int iters, double* __restrict__ out) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) probe_dfma_kernel(int iters, double* __restrict__ out) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}
}  // namespace tal

extern "C" {
/* kind 0: DMMA.8x8x4 (2*256 flop per warp instruction), kind 1: DFMA.  Launches
 * `blocks` CTAs of 256 threads running `iters` x 16 independent ops per thread;
 * flops = blocks * iters * 16 * (kind == 0 ? 8 * 512 : 256 * 2).              */
int tal_probe_fp64(int kind, int blocks, int iters, double* out, void* stream) {
  TAL_ARG(out && blocks >= 1 && iters >= 1);
  if (kind == 0) tal::probe_dmma_kernel<<<blocks, 256, 0, tal::as_stream(stream)>>>(iters, out);
  else tal::probe_dfma_kernel<<<blocks, 256, 0, tal::as_stream(stream)>>>(iters, out);
  TAL_LAUNCH_CHECK("tal_probe_fp64");
  return TAL_OK;
}
}
