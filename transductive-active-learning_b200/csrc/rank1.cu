// Posterior update after one observation, batched over the q GPs.
//
// Reference path: MarginalModel.step (lib/model/marginal.py:350-368) ->
// GaussianDistribution.posterior (lib/gp/gaussian_distribution.py:100-121) ->
// _compute_posterior_covariance (:16-37) with m = 1.  There Sigma_oo is 1x1,
// so solve_linear_system (lib/utils.py:22-23) is  L = sqrt(s), w = (k/L)/L and
// the N x N update is  Sigma' = Sigma - k^T w  (:36), written out of place by
// XLA.  Here it is applied in place by a pure streaming kernel.
#include "common.cuh"

namespace tal {

// Lower-block ("symmetric") storage: of global row g only the columns
// [0, sym_cend(g)) are maintained; entry (g, c) with c >= sym_cend(g) is read as
// (c, g).  sym_block (32 elements) is the row count of an update CTA and a multiple
// of the per-thread vector width.
__device__ __forceinline__ i64 sym_cend(i64 grow, i64 block, i64 ld) {
  const i64 e = (grow / block + 1) * block;
  return e < ld ? e : ld;
}

// --------------------------------------------------------------------------
// Phase 1: copy the observed row (or zeros on a shard that does not own it).
// --------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
extract_row_kernel(const T* __restrict__ cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc,
                   const i64* __restrict__ idx_dev, i64 idx_host, const double* __restrict__ mean,
                   i64 N, T* __restrict__ kbuf, double* __restrict__ scal) {
  const int gp = blockIdx.y;
  const i64 idx = idx_dev ? *idx_dev : idx_host;
  const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  const i64 r = idx - row0;
  const bool own = r >= 0 && r < n_loc;
  if (col < ld) {
    T v = T(0);
    if (own) v = cov[gp * stride_q + r * ld + col];
    kbuf[(i64)gp * ld + col] = v;
  }
  if (col == 0 && idx >= 0 && idx < N) scal[gp] = mean[(i64)gp * N + idx];
}

__global__ void stash_mean_kernel(const double* __restrict__ mean, i64 N, int q,
                                  const i64* __restrict__ idx_dev, i64 idx_host, double* __restrict__ scal) {
  const i64 idx = idx_dev ? *idx_dev : idx_host;
  if ((int)threadIdx.x < q && idx >= 0 && idx < N) scal[threadIdx.x] = mean[(i64)threadIdx.x * N + idx];
}

// --------------------------------------------------------------------------
// Phase 2: the O(qN) state: w, mean, var, l, u.
// --------------------------------------------------------------------------
template <typename T, bool FMA>
__global__ void __launch_bounds__(256)
step_vectors_kernel(const T* __restrict__ kbuf, T* __restrict__ wbuf, i64 N, i64 ld,
                    const i64* __restrict__ idx_dev, i64 idx_host,
                    const double* __restrict__ y_dev, i64 y_stride, int y_gather,
                    const double* __restrict__ rho, BetaQ beta, const double* __restrict__ scal,
                    double* __restrict__ mean, double* __restrict__ var, double* __restrict__ l,
                    double* __restrict__ u) {
  const int gp = blockIdx.y;
  const i64 idx = idx_dev ? *idx_dev : idx_host;
  const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= ld) return;
  if (idx < 0 || idx >= N) {  // failed arg-max (idx = -1): the observation is a no-op (k = 0 from the
    wbuf[(i64)gp * ld + b] = T(0);  // row extract, w = 0: a queued update changes nothing)
    return;
  }
  const T* k = kbuf + (i64)gp * ld;
  // Sigma = cov[idx, idx] + noise_std^2 (gaussian_distribution.py:29-32); the
  // arithmetic is done in T so that var stays bit-identical to diag(cov).
  const double r = rho[(i64)gp * N + idx];
  const T s = T((double)k[idx] + r * r);
  const T L = sqrt(s);  // IEEE sqrt / div (no fast-math)
  const T kb = k[b];
  const T wb = (kb / L) / L;  // solve(L.T, solve(L, k)) for 1x1 L (utils.py:23)
  wbuf[(i64)gp * ld + b] = wb;
  if (b >= N) return;
  const double yv = y_gather ? y_dev[(i64)gp * y_stride + idx] : y_dev[(i64)gp * y_stride];
  const double delta = yv - scal[gp];  // obs - prior_mean[obs_idx] (:120)
  const i64 o = (i64)gp * N + b;
  const double m_new = add_prod(mean[o], (double)wb, delta);
  const T v_new = sub_prod_m<FMA>(T(var[o]), kb, wb);  // == diag of the matrix update
  mean[o] = m_new;
  var[o] = (double)v_new;
  const double sd = sqrt((double)v_new);  // NaN for negative variance, as the reference
  const double bt = beta.v[gp];
  l[o] = max_nan(l[o], __dsub_rn(m_new, __dmul_rn(bt, sd)));
  u[o] = min_nan(u[o], __dadd_rn(m_new, __dmul_rn(bt, sd)));
}

// --------------------------------------------------------------------------
// Phase 3: cov[a][b] -= k[a] * w[b], in place.  HBM streaming.
// --------------------------------------------------------------------------

// Variant 1: 16-byte LDG/STG, UNROLL independent rows in flight per thread.
template <typename T, int UNROLL, bool FMA>
__global__ void __launch_bounds__(256)
rank1_ldg16_kernel(T* __restrict__ cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc,
                   const T* __restrict__ kbuf, const T* __restrict__ wbuf, int rows_per_cta) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  const int gp = blockIdx.z;
  const i64 col = ((i64)blockIdx.y * 256 + threadIdx.x) * VN;
  if (col >= ld) return;
  T* C = cov + gp * stride_q + col;
  const T* k = kbuf + (i64)gp * ld + row0;
  const V wv = *reinterpret_cast<const V*>(wbuf + (i64)gp * ld + col);
  const i64 r0 = (i64)blockIdx.x * rows_per_cta;
  const i64 r1 = min(n_loc, r0 + (i64)rows_per_cta);
  for (i64 r = r0; r < r1; r += UNROLL) {
    V c[UNROLL];
    T kk[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      if (r + j < r1) {
        c[j] = ld_stream(reinterpret_cast<const V*>(C + (r + j) * ld));
        kk[j] = __ldg(k + r + j);
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      if (r + j < r1) {
        if constexpr (VN == 2) {
          c[j].x = sub_prod_m<FMA>(c[j].x, kk[j], wv.x);
          c[j].y = sub_prod_m<FMA>(c[j].y, kk[j], wv.y);
        } else {
          c[j].x = sub_prod_m<FMA>(c[j].x, kk[j], wv.x);
          c[j].y = sub_prod_m<FMA>(c[j].y, kk[j], wv.y);
          c[j].z = sub_prod_m<FMA>(c[j].z, kk[j], wv.z);
          c[j].w = sub_prod_m<FMA>(c[j].w, kk[j], wv.w);
        }
        st_stream(reinterpret_cast<V*>(C + (r + j) * ld), c[j]);
      }
    }
  }
}

// Variant 3: 32-byte LDG/STG (LDG.E.256, new on sm_100).
struct alignas(32) D4 {
  double v[4];
};
struct alignas(32) F8 {
  float v[8];
};
__device__ __forceinline__ D4 ld256(const double* p) {
  D4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(double* p, const D4& r) {
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]),
               "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3])
               : "memory");
}
__device__ __forceinline__ F8 ld256(const float* p) {
  F8 r;
  asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]),
                 "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st256(float* p, const F8& r) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
               "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]),
               "f"(r.v[6]), "f"(r.v[7])
               : "memory");
}
template <typename T>
struct Vec32;
template <>
struct Vec32<double> {
  typedef D4 type;
  static constexpr int n = 4;
};
template <>
struct Vec32<float> {
  typedef F8 type;
  static constexpr int n = 8;
};

template <typename T, int UNROLL, bool FMA>
__global__ void __launch_bounds__(256)
rank1_ldg32_kernel(T* __restrict__ cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc,
                   const T* __restrict__ kbuf, const T* __restrict__ wbuf, int rows_per_cta,
                   i64 sym_block) {
  typedef typename Vec32<T>::type V;
  constexpr int VN = Vec32<T>::n;
  const int gp = blockIdx.z;
  // lower-block storage: column tile in x (fastest), row block in y
  const i64 rb = sym_block > 0 ? blockIdx.y : blockIdx.x;
  const i64 ct = sym_block > 0 ? blockIdx.x : blockIdx.y;
  const i64 col = (ct * 256 + threadIdx.x) * VN;
  if (col >= ld) return;
  const i64 r0 = rb * rows_per_cta;
  // only the blocks on or below the diagonal are kept (the CTA's 32 rows share one sym block):
  // whole tiles above it exit here, the tile that straddles it is cut per thread
  if (sym_block > 0 && col >= sym_cend(row0 + r0, sym_block, ld)) return;
  T* C = cov + gp * stride_q + col;
  const T* k = kbuf + (i64)gp * ld + row0;
  V wv = *reinterpret_cast<const V*>(wbuf + (i64)gp * ld + col);
  const i64 r1 = min(n_loc, r0 + (i64)rows_per_cta);
  for (i64 r = r0; r < r1; r += UNROLL) {
    V c[UNROLL];
    T kk[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      if (r + j < r1) {
        c[j] = ld256(C + (r + j) * ld);
        kk[j] = __ldg(k + r + j);
      }
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      if (r + j < r1) {
#pragma unroll
        for (int e = 0; e < VN; ++e) c[j].v[e] = sub_prod_m<FMA>(c[j].v[e], kk[j], wv.v[e]);
        st256(C + (r + j) * ld, c[j]);
      }
    }
  }
}

// Variant 2: bulk-async (TMA 1-D) pipeline.  One CTA owns a column segment
// of SEG bytes and walks down a chunk of rows; a ring of STAGES shared-memory
// buffers is filled by cp.async.bulk (global -> shared, mbarrier completion),
// updated in place by all threads, and drained by cp.async.bulk (shared ->
// global, bulk-group completion).  Bytes in flight per CTA: STAGES * SEG.
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <typename T, int SEG_BYTES, int STAGES, int THREADS, bool FMA>
__global__ void __launch_bounds__(THREADS)
rank1_bulk_kernel(T* __restrict__ cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc,
                  const T* __restrict__ kbuf, const T* __restrict__ wbuf, int rows_per_cta) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  constexpr int SEG_ELEMS = SEG_BYTES / (int)sizeof(T);
  constexpr int VEC_PER_THREAD = SEG_BYTES / 16 / THREADS;
  static_assert(VEC_PER_THREAD >= 1 && SEG_BYTES % (16 * THREADS) == 0, "segment/thread mismatch");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* buf = reinterpret_cast<T*>(smem_raw);                                  // [STAGES][SEG_ELEMS]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * SEG_BYTES);

  const int gp = blockIdx.z;
  const i64 col0 = (i64)blockIdx.y * SEG_ELEMS;
  const i64 rem = ld - col0;  // elements left in the row from col0 (multiple of 16 B)
  const uint32_t seg_bytes = (uint32_t)((rem < SEG_ELEMS ? rem : (i64)SEG_ELEMS) * sizeof(T));
  const int seg_vecs = seg_bytes / 16;
  T* C = cov + gp * stride_q + col0;
  const T* k = kbuf + (i64)gp * ld + row0;
  const i64 r0 = (i64)blockIdx.x * rows_per_cta;
  const int nrows = (int)(min(n_loc, r0 + (i64)rows_per_cta) - r0);
  if (nrows <= 0) return;

  // this thread's slice of w stays in registers for the whole column walk
  V wv[VEC_PER_THREAD];
#pragma unroll
  for (int j = 0; j < VEC_PER_THREAD; ++j) {
    const int v = threadIdx.x + j * THREADS;
    if (v < seg_vecs) wv[j] = *reinterpret_cast<const V*>(wbuf + (i64)gp * ld + col0 + (i64)v * VN);
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int pre = nrows < STAGES ? nrows : STAGES;
    for (int s = 0; s < pre; ++s) {
      mbar_expect_tx(&full[s], seg_bytes);
      bulk_load(buf + (size_t)s * SEG_ELEMS, C + (r0 + s) * ld, seg_bytes, &full[s]);
    }
  }
  for (int it = 0; it < nrows; ++it) {
    const int s = it % STAGES;
    const uint32_t parity = (it / STAGES) & 1;
    mbar_wait(&full[s], parity);
    const T kk = __ldg(k + r0 + it);
    V* sv = reinterpret_cast<V*>(buf + (size_t)s * SEG_ELEMS);
#pragma unroll
    for (int j = 0; j < VEC_PER_THREAD; ++j) {
      const int v = threadIdx.x + j * THREADS;
      if (v < seg_vecs) {
        V c = sv[v];
        if constexpr (VN == 2) {
          c.x = sub_prod_m<FMA>(c.x, kk, wv[j].x);
          c.y = sub_prod_m<FMA>(c.y, kk, wv[j].y);
        } else {
          c.x = sub_prod_m<FMA>(c.x, kk, wv[j].x);
          c.y = sub_prod_m<FMA>(c.y, kk, wv[j].y);
          c.z = sub_prod_m<FMA>(c.z, kk, wv[j].z);
          c.w = sub_prod_m<FMA>(c.w, kk, wv[j].w);
        }
        sv[v] = c;
      }
    }
    fence_async_smem();  // generic-proxy writes -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(C + (r0 + it) * ld, sv, seg_bytes);
      bulk_commit();
      // refill the stage drained one iteration ago: its store has finished
      // reading shared memory once at most one bulk group is still pending.
      const int prev = it - 1;
      if (prev >= 0 && prev + STAGES < nrows) {
        bulk_wait_read<1>();
        const int ps = prev % STAGES;
        mbar_expect_tx(&full[ps], seg_bytes);
        bulk_load(buf + (size_t)ps * SEG_ELEMS, C + (r0 + prev + STAGES) * ld, seg_bytes,
                  &full[ps]);
      }
    }
  }
  if (threadIdx.x == 0) bulk_wait_all();
}

// --------------------------------------------------------------------------
// m-observation update (one GP per call)
// --------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const T* __restrict__ cov, i64 ld, i64 row0, i64 n_loc, i64 N,
                   const i64* __restrict__ obs_idx, double* __restrict__ B, i64 ldb) {
  const i64 j = blockIdx.y;
  const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  const i64 r = obs_idx[j] - row0;
  double v = 0.0;
  if (r >= 0 && r < n_loc) v = (double)cov[r * ld + col];
  B[j * ldb + col] = v;
}

// Lower-block storage: row idx is assembled from the stored part of the row
// (columns below sym_cend(idx), on the shard that owns the row) and from column idx
// of the rows at or beyond sym_cend(idx) (on the shards that own those rows); what
// this shard does not hold is written as zero, so a sum over shards is the row.
// grid (ceil(ld / 256), m rows, q).
template <typename T, typename O>
__global__ void __launch_bounds__(256)
gather_rows_sym_kernel(const T* __restrict__ cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc, i64 N,
                       i64 sym_block, const i64* __restrict__ idx_list, i64 idx_host,
                       O* __restrict__ out, i64 ldo, i64 out_stride_q, i64 ncols_out) {
  const int gp = blockIdx.z;
  const i64 idx = idx_list ? idx_list[blockIdx.y] : idx_host;
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncols_out) return;
  const T* C = cov + gp * stride_q;
  T v = T(0);
  if (idx >= 0 && idx < N && c < N) {
    if (c < sym_cend(idx, sym_block, ld)) {
      const i64 r = idx - row0;
      if (r >= 0 && r < n_loc) v = C[r * ld + c];
    } else {
      const i64 r = c - row0;
      if (r >= 0 && r < n_loc) v = C[r * ld + idx];
    }
  }
  out[gp * out_stride_q + (i64)blockIdx.y * ldo + c] = O(v);
}

// Fill the blocks above the diagonal from their mirror images (one GPU, all rows):
// cov[r][c] = cov[c][r] for c >= sym_cend(r).  32 x 32 tiles through shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
sym_mirror_kernel(T* __restrict__ cov, i64 stride_q, i64 ld, i64 N, i64 sym_block) {
  __shared__ T tile[32][33];
  const int gp = blockIdx.z;
  const i64 tr = blockIdx.y, tc = blockIdx.x;  // destination tile (rows tr*32.., cols tc*32..)
  if (tc * 32 < sym_cend(tr * 32, sym_block, ld)) return;  // kept block: nothing to do
  T* C = cov + gp * stride_q;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const i64 r = tc * 32 + j, c = tr * 32 + tx;  // source (r, c) = mirror of the destination tile
    tile[j][tx] = (r < N && c < N) ? C[r * ld + c] : T(0);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const i64 r = tr * 32 + j, c = tc * 32 + tx;
    if (r < N && c < N) C[r * ld + c] = tile[tx][j];
  }
}

// One CTA: S = B[:, obs] + D, Cholesky in shared memory; then every thread
// solves its own columns.  m <= 64.
constexpr int kSmallM = 64;
__global__ void __launch_bounds__(256)
small_chol_kernel(const double* __restrict__ B, i64 ldb, const i64* __restrict__ obs_idx, int m,
                  const double* __restrict__ noise_std, double jitter, double* __restrict__ Lout) {
  __shared__ double S[kSmallM][kSmallM + 1];
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e % m;
    double v = B[(i64)i * ldb + obs_idx[j]];
    if (i == j) {
      const double sd = noise_std ? noise_std[i] : jitter;  // utils.py:47-50 (default squared too)
      v = v + sd * sd;
    }
    S[i][j] = v;
  }
  __syncthreads();
  // right-looking Cholesky (lower), NaN on a non-positive pivot
  for (int c = 0; c < m; ++c) {
    if (threadIdx.x == 0) S[c][c] = sqrt(S[c][c]);
    __syncthreads();
    const double d = S[c][c];
    for (int i = c + 1 + threadIdx.x; i < m; i += blockDim.x) S[i][c] = S[i][c] / d;
    __syncthreads();
    for (int e = threadIdx.x; e < (m - c - 1) * (m - c - 1); e += blockDim.x) {
      const int i = c + 1 + e / (m - c - 1), j = c + 1 + e % (m - c - 1);
      if (j <= i) S[i][j] -= S[i][c] * S[j][c];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e % m;
    Lout[e] = j <= i ? S[i][j] : 0.0;
  }
}

// W[:, b] = (L L^T)^-1 B[:, b]; mean/var updates.  delta_j = y_j - mean[obs_j]
// must be computed before any mean is overwritten -> done in a first kernel.
__global__ void small_delta_kernel(const double* __restrict__ mean, const i64* __restrict__ obs_idx,
                                   const double* __restrict__ y, int m, double* __restrict__ delta) {
  const int j = threadIdx.x;
  if (j < m) delta[j] = y[j] - mean[obs_idx[j]];
}

__global__ void __launch_bounds__(128)
small_solve_kernel(const double* __restrict__ B, i64 ldb, i64 N, int m,
                   const double* __restrict__ Lg, const double* __restrict__ delta,
                   double* __restrict__ W, double* __restrict__ mean, double* __restrict__ var) {
  __shared__ double L[kSmallM * kSmallM];
  __shared__ double dl[kSmallM];
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) L[e] = Lg[e];
  if (threadIdx.x < m) dl[threadIdx.x] = delta[threadIdx.x];
  __syncthreads();
  const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  double x[kSmallM];
  // forward: L z = B[:, b]
  for (int i = 0; i < m; ++i) {
    double acc = B[(i64)i * ldb + b];
    for (int j = 0; j < i; ++j) acc -= L[i * m + j] * x[j];
    x[i] = acc / L[i * m + i];
  }
  // backward: L^T w = z
  for (int i = m - 1; i >= 0; --i) {
    double acc = x[i];
    for (int j = i + 1; j < m; ++j) acc -= L[j * m + i] * x[j];
    x[i] = acc / L[i * m + i];
  }
  double dm = 0.0, dv = 0.0;
  for (int j = 0; j < m; ++j) {
    W[(i64)j * ldb + b] = x[j];
    dm += x[j] * dl[j];
    dv = add_prod(dv, B[(i64)j * ldb + b], x[j]);  // same accumulation as the matrix kernel
  }
  mean[b] = mean[b] + dm;
  var[b] = __dsub_rn(var[b], dv);
}

template <typename T, int UNROLL>
__global__ void __launch_bounds__(256)
rankm_update_kernel(T* __restrict__ cov, i64 ld, i64 row0, i64 n_loc, i64 N,
                    const double* __restrict__ B, const double* __restrict__ W, i64 ldb, int m,
                    int rows_per_cta, i64 sym_block) {
  const i64 col = (i64)blockIdx.y * 256 + threadIdx.x;
  if (col >= N) return;
  const i64 r0 = (i64)blockIdx.x * rows_per_cta;
  if (sym_block > 0 && col >= sym_cend(row0 + r0, sym_block, ld)) return;
  double w[kSmallM];
  for (int j = 0; j < m; ++j) w[j] = W[(i64)j * ldb + col];
  const i64 r1 = min(n_loc, r0 + (i64)rows_per_cta);
  for (i64 r = r0; r < r1; ++r) {
    double acc = 0.0;
    for (int j = 0; j < m; ++j) acc = add_prod(acc, __ldg(B + (i64)j * ldb + row0 + r), w[j]);
    const double c = (double)cov[r * ld + col];
    cov[r * ld + col] = T(__dsub_rn(c, acc));
  }
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_extract_row(const void* cov, long long cov_stride_q, long long ld, long long row0,
                    long long n_loc, long long N, int q, const long long* idx_dev,
                    long long idx_host, const double* mean, void* kbuf, double* scal, int dtype,
                    void* stream) {
  TAL_ARG(cov && kbuf && scal && mean);
  TAL_ARG(q >= 1 && q <= TAL_MAX_Q && N >= 1 && ld >= N && n_loc >= 0);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  dim3 grid((unsigned)((ld + 255) / 256), q);
  if (dtype == TAL_F64)
    extract_row_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
        (const double*)cov, cov_stride_q, ld, row0, n_loc, idx_dev, idx_host, mean, N,
        (double*)kbuf, scal);
  else
    extract_row_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
        (const float*)cov, cov_stride_q, ld, row0, n_loc, idx_dev, idx_host, mean, N,
        (float*)kbuf, scal);
  TAL_LAUNCH_CHECK("tal_extract_row");
  return TAL_OK;
}

int tal_step_vectors(const void* kbuf, void* wbuf, long long N, int q, const long long* idx_dev,
                     long long idx_host, const double* y_dev, long long y_stride, int y_gather,
                     const double* rho, const double* beta_host, const double* scal,
                     double* mean, double* var, double* l, double* u, int dtype, void* stream) {
  TAL_ARG(kbuf && wbuf && y_dev && rho && beta_host && scal && mean && var && l && u);
  TAL_ARG(q >= 1 && q <= TAL_MAX_Q && N >= 1);
  const bool fma_mode = arith_fma(dtype);
  dtype = base_dtype(dtype);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  BetaQ beta;
  for (int i = 0; i < TAL_MAX_Q; ++i) beta.v[i] = i < q ? beta_host[i] : 0.0;
  // kbuf / wbuf rows are `ld` elements apart, ld = N rounded up to 32
  const i64 ld = (N + 31) / 32 * 32;
  dim3 grid((unsigned)((ld + 255) / 256), q);
#define TAL_SV(T, F)                                                                              \
  step_vectors_kernel<T, F><<<grid, 256, 0, as_stream(stream)>>>(                                 \
      (const T*)kbuf, (T*)wbuf, N, ld, idx_dev, idx_host, y_dev, y_stride, y_gather, rho, beta,   \
      scal, mean, var, l, u)
  if (dtype == TAL_F64) { if (fma_mode) TAL_SV(double, true); else TAL_SV(double, false); }
  else { if (fma_mode) TAL_SV(float, true); else TAL_SV(float, false); }
#undef TAL_SV
  TAL_LAUNCH_CHECK("tal_step_vectors");
  return TAL_OK;
}

}  // extern "C"

template <typename T, bool FMA>
static int launch_rank1(T* cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc, int q, const T* kbuf,
                        const T* wbuf, int variant, cudaStream_t st, i64 sym_block = 0) {
  if (n_loc == 0) return TAL_OK;
  if (variant == 0 || sym_block > 0) variant = 3;
  if (variant == 1) {
    constexpr int VN = Vec16<T>::n;
    const int rows_per_cta = 64;
    dim3 grid((unsigned)((n_loc + rows_per_cta - 1) / rows_per_cta),
              (unsigned)((ld + 256 * VN - 1) / (256 * VN)), q);
    rank1_ldg16_kernel<T, 8, FMA><<<grid, 256, 0, st>>>(cov, stride_q, ld, row0, n_loc, kbuf, wbuf,
                                                   rows_per_cta);
  } else if (variant == 3) {
    constexpr int VN = Vec32<T>::n;
    const int rows_per_cta = 32;
    const unsigned row_blocks = (unsigned)((n_loc + rows_per_cta - 1) / rows_per_cta);
    const unsigned col_tiles = (unsigned)((ld + 256 * VN - 1) / (256 * VN));
    if (sym_block > 0) {
      TAL_ARG(sym_block % VN == 0 && sym_block % rows_per_cta == 0 && row0 % rows_per_cta == 0);
      TAL_ARG(row_blocks <= 65535);
    }
    dim3 grid = sym_block > 0 ? dim3(col_tiles, row_blocks, q) : dim3(row_blocks, col_tiles, q);
    rank1_ldg32_kernel<T, 4, FMA><<<grid, 256, 0, st>>>(cov, stride_q, ld, row0, n_loc, kbuf, wbuf,
                                                   rows_per_cta, sym_block);
  } else if (variant == 2) {
    constexpr int SEG = 8192, STAGES = 6, THREADS = 256;
    constexpr int SEG_ELEMS = SEG / (int)sizeof(T);
    const int rows_per_cta = 128;
    const size_t smem = (size_t)STAGES * SEG + STAGES * sizeof(uint64_t);
    static bool attr_done = false;  // (one flag per instantiation of this function template)
    auto kern = rank1_bulk_kernel<T, SEG, STAGES, THREADS, FMA>;
    if (!attr_done) {
      TAL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done = true;
    }
    dim3 grid((unsigned)((n_loc + rows_per_cta - 1) / rows_per_cta),
              (unsigned)((ld + SEG_ELEMS - 1) / SEG_ELEMS), q);
    kern<<<grid, THREADS, smem, st>>>(cov, stride_q, ld, row0, n_loc, kbuf, wbuf, rows_per_cta);
  } else {
    tal::set_error("tal_rank1_update: unknown variant %d", variant);
    return TAL_ERR_ARG;
  }
  TAL_LAUNCH_CHECK("tal_rank1_update");
  return TAL_OK;
}

extern "C" {

int tal_rank1_update(void* cov, long long cov_stride_q, long long ld, long long row0,
                     long long n_loc, long long N, int q, const void* kbuf, const void* wbuf,
                     int dtype, int variant, void* stream) {
  TAL_ARG(cov && kbuf && wbuf);
  TAL_ARG(q >= 1 && q <= TAL_MAX_Q && N >= 1 && ld >= N && ld % 32 == 0 && n_loc >= 0);
  TAL_ARG(ld == (N + 31) / 32 * 32);
  const bool fma_mode = arith_fma(dtype);
  dtype = base_dtype(dtype);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
#define TAL_R1(T, F)                                                                               \
  return launch_rank1<T, F>((T*)cov, cov_stride_q, ld, row0, n_loc, q, (const T*)kbuf, (const T*)wbuf, \
                            variant, as_stream(stream))
  if (dtype == TAL_F64) { if (fma_mode) TAL_R1(double, true); TAL_R1(double, false); }
  if (fma_mode) TAL_R1(float, true);
  TAL_R1(float, false);
#undef TAL_R1
}

int tal_gather_rows(const void* cov, long long ld, long long row0, long long n_loc, long long N,
                    const long long* obs_idx, long long m, double* B, long long ldb, int dtype,
                    void* stream) {
  TAL_ARG(cov && obs_idx && B && m >= 0 && ldb >= N);
  if (m == 0) return TAL_OK;
  TAL_ARG(m <= 65535);
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)m);
  if (dtype == TAL_F64)
    gather_rows_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)cov, ld, row0,
                                                                    n_loc, N, obs_idx, B, ldb);
  else
    gather_rows_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)cov, ld, row0,
                                                                   n_loc, N, obs_idx, B, ldb);
  TAL_LAUNCH_CHECK("tal_gather_rows");
  return TAL_OK;
}

int tal_small_posterior_weights(const double* B, long long ldb, long long N,
                                const long long* obs_idx, int m, const double* noise_std,
                                double jitter, const double* y, double* W, double* mean,
                                double* var, double* scratch, void* stream) {
  TAL_ARG(B && obs_idx && y && W && mean && var && scratch);
  TAL_ARG(m >= 1 && m <= kSmallM);
  double* Lg = scratch;  // tal_small_scratch_bytes(): the m x m factor and the m residuals
  double* delta = Lg + kSmallM * kSmallM;
  cudaStream_t st = as_stream(stream);
  small_chol_kernel<<<1, 256, 0, st>>>(B, ldb, obs_idx, m, noise_std, jitter, Lg);
  TAL_LAUNCH_CHECK("tal_small_posterior_weights/chol");
  small_delta_kernel<<<1, kSmallM, 0, st>>>(mean, obs_idx, y, m, delta);
  TAL_LAUNCH_CHECK("tal_small_posterior_weights/delta");
  small_solve_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(B, ldb, N, m, Lg, delta, W, mean,
                                                                 var);
  TAL_LAUNCH_CHECK("tal_small_posterior_weights/solve");
  return TAL_OK;
}

long long tal_small_scratch_bytes(void) { return (long long)sizeof(double) * (kSmallM * kSmallM + kSmallM); }

int tal_rankm_update(void* cov, long long ld, long long row0, long long n_loc, long long N,
                     const double* B, const double* W, long long ldb, int m, int dtype,
                     void* stream) {
  TAL_ARG(cov && B && W && m >= 1 && m <= kSmallM);
  if (n_loc == 0) return TAL_OK;
  const int rows_per_cta = 32;
  dim3 grid((unsigned)((n_loc + rows_per_cta - 1) / rows_per_cta), (unsigned)((N + 255) / 256));
  if (dtype == TAL_F64)
    rankm_update_kernel<double, 1><<<grid, 256, 0, as_stream(stream)>>>(
        (double*)cov, ld, row0, n_loc, N, B, W, ldb, m, rows_per_cta, 0);
  else
    rankm_update_kernel<float, 1><<<grid, 256, 0, as_stream(stream)>>>(
        (float*)cov, ld, row0, n_loc, N, B, W, ldb, m, rows_per_cta, 0);
  TAL_LAUNCH_CHECK("tal_rankm_update");
  return TAL_OK;
}

/* ---- lower-block ("symmetric") storage ------------------------------------ */
long long tal_sym_block(int dtype) { (void)dtype; return 32; }

int tal_rank1_update_lower(void* cov, long long cov_stride_q, long long ld, long long row0,
                           long long n_loc, long long N, int q, const void* kbuf, const void* wbuf,
                           int dtype, void* stream) {
  TAL_ARG(cov && kbuf && wbuf);
  TAL_ARG(q >= 1 && q <= TAL_MAX_Q && N >= 1 && ld == (N + 31) / 32 * 32 && n_loc >= 0);
  const bool fma_mode = arith_fma(dtype);
  dtype = base_dtype(dtype);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  const i64 blk = tal_sym_block(dtype);
#define TAL_R1(T, F)                                                                               \
  return launch_rank1<T, F>((T*)cov, cov_stride_q, ld, row0, n_loc, q, (const T*)kbuf, (const T*)wbuf, \
                            3, as_stream(stream), blk)
  if (dtype == TAL_F64) { if (fma_mode) TAL_R1(double, true); TAL_R1(double, false); }
  if (fma_mode) TAL_R1(float, true);
  TAL_R1(float, false);
#undef TAL_R1
}

int tal_extract_row_lower(const void* cov, long long cov_stride_q, long long ld, long long row0,
                          long long n_loc, long long N, int q, const long long* idx_dev,
                          long long idx_host, const double* mean, void* kbuf, double* scal, int dtype,
                          void* stream) {
  TAL_ARG(cov && kbuf && scal && mean);
  TAL_ARG(q >= 1 && q <= TAL_MAX_Q && N >= 1 && ld >= N && n_loc >= 0);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  const i64 blk = tal_sym_block(dtype);
  dim3 grid((unsigned)((ld + 255) / 256), 1, q);
  if (dtype == TAL_F64)
    gather_rows_sym_kernel<double, double><<<grid, 256, 0, as_stream(stream)>>>(
        (const double*)cov, cov_stride_q, ld, row0, n_loc, N, blk, idx_dev, idx_host, (double*)kbuf, ld,
        ld, ld);
  else
    gather_rows_sym_kernel<float, float><<<grid, 256, 0, as_stream(stream)>>>(
        (const float*)cov, cov_stride_q, ld, row0, n_loc, N, blk, idx_dev, idx_host, (float*)kbuf, ld, ld,
        ld);
  TAL_LAUNCH_CHECK("tal_extract_row_lower");
  stash_mean_kernel<<<1, 32, 0, as_stream(stream)>>>(mean, N, q, idx_dev, idx_host, scal);
  TAL_LAUNCH_CHECK("tal_extract_row_lower/scal");
  return TAL_OK;
}

int tal_gather_rows_lower(const void* cov, long long ld, long long row0, long long n_loc, long long N,
                          const long long* obs_idx, long long m, double* B, long long ldb, int dtype,
                          void* stream) {
  TAL_ARG(cov && obs_idx && B && m >= 0 && ldb >= N);
  if (m == 0) return TAL_OK;
  TAL_ARG(m <= 65535);
  const i64 blk = tal_sym_block(dtype);
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)m, 1);
  if (dtype == TAL_F64)
    gather_rows_sym_kernel<double, double><<<grid, 256, 0, as_stream(stream)>>>(
        (const double*)cov, 0, ld, row0, n_loc, N, blk, obs_idx, 0, B, ldb, 0, N);
  else
    gather_rows_sym_kernel<float, double><<<grid, 256, 0, as_stream(stream)>>>(
        (const float*)cov, 0, ld, row0, n_loc, N, blk, obs_idx, 0, B, ldb, 0, N);
  TAL_LAUNCH_CHECK("tal_gather_rows_lower");
  return TAL_OK;
}

int tal_rankm_update_lower(void* cov, long long ld, long long row0, long long n_loc, long long N,
                           const double* B, const double* W, long long ldb, int m, int dtype,
                           void* stream) {
  TAL_ARG(cov && B && W && m >= 1 && m <= kSmallM);
  if (n_loc == 0) return TAL_OK;
  const int rows_per_cta = 32;
  const i64 blk = tal_sym_block(dtype);
  TAL_ARG(row0 % rows_per_cta == 0);
  dim3 grid((unsigned)((n_loc + rows_per_cta - 1) / rows_per_cta), (unsigned)((N + 255) / 256));
  if (dtype == TAL_F64)
    rankm_update_kernel<double, 1><<<grid, 256, 0, as_stream(stream)>>>(
        (double*)cov, ld, row0, n_loc, N, B, W, ldb, m, rows_per_cta, blk);
  else
    rankm_update_kernel<float, 1><<<grid, 256, 0, as_stream(stream)>>>(
        (float*)cov, ld, row0, n_loc, N, B, W, ldb, m, rows_per_cta, blk);
  TAL_LAUNCH_CHECK("tal_rankm_update_lower");
  return TAL_OK;
}

int tal_sym_mirror(void* cov, long long cov_stride_q, long long ld, long long N, int q, int dtype,
                   void* stream) {
  TAL_ARG(cov && q >= 1 && q <= TAL_MAX_Q && N >= 1 && ld >= N);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  const i64 blk = tal_sym_block(dtype);
  const unsigned t = (unsigned)((N + 31) / 32);
  TAL_ARG(t <= 65535);
  dim3 grid(t, t, q);
  if (dtype == TAL_F64)
    sym_mirror_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((double*)cov, cov_stride_q, ld, N, blk);
  else
    sym_mirror_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((float*)cov, cov_stride_q, ld, N, blk);
  TAL_LAUNCH_CHECK("tal_sym_mirror");
  return TAL_OK;
}

}  // extern "C"
