// Per-candidate acquisition scores.
//
// Reference path: ITL._undirected (lib/algorithms/itl.py:28-30), VTL.F
// (lib/algorithms/vtl.py:16-41) and the CTL / MM-ITL reductions over
// MarginalModel.sqd_correlation (lib/model/marginal.py:129-158,
// lib/algorithms/ctl.py:14-17, mm_itl.py:14-17).  The reference gathers
// cov[a, x] for a in the ROI through nested vmaps; here the ROI rows are
// streamed once (m*N*sizeof(T) bytes) and reduced per column.
#include "common.cuh"
#include <cstdlib>

namespace tal {

__global__ void __launch_bounds__(256)
itl_undirected_kernel(const double* __restrict__ var, const double* __restrict__ rho, int q, i64 N,
                      double* __restrict__ F) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= N) return;
  double acc = 0.0;
  for (int i = 0; i < q; ++i) {
    const double r = rho[(i64)i * N + x];
    acc += log(1.0 + var[(i64)i * N + x] / (r * r));  // itl.py:29-30 (no 1/2)
  }
  F[x] = acc;
}

// Element functor of the row reduce: what each cov[a][x] contributes.
//   VTL / CTL:  w_a * v^2
//   MM-ITL:    -0.5 * log(1 - w_a * v^2 * dinv[x])      (mm_itl.py:15-16)
//   TruVar:     max(beta * (var_a - v^2 * dinv[x]), eta^2) (tru_var/__init__.py:31-45)
template <int RULE>
__device__ __forceinline__ double elem(double v, double wa, double dinv, double p0, double p1) {
  if constexpr (RULE == TAL_RULE_VTL || RULE == TAL_RULE_CTL) {
    return wa * (v * v);
  } else if constexpr (RULE == TAL_RULE_MMITL) {
    return -0.5 * log(1.0 - (v * v) * wa * dinv);
  } else {
    return fmax(p0 * (wa - (v * v) * dinv), p1);
  }
}

// Column sums over ROI rows.  grid = (column tiles, row splits).  Each thread
// owns one 16-byte vector of columns; ROI rows [j0, j1) of its split are read
// UNROLL at a time so that UNROLL independent 16-byte loads are in flight.
template <typename T, int RULE, int UNROLL>
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const T* __restrict__ cov, i64 ld, i64 row0, i64 n_loc, i64 N,
                   const i64* __restrict__ roi_idx, i64 m, const double* __restrict__ row_weight,
                   const double* __restrict__ dinv, double p0, double p1,
                   double* __restrict__ part, i64 sym_block) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  const i64 col = ((i64)blockIdx.x * 256 + threadIdx.x) * VN;
  const i64 per = (m + gridDim.y - 1) / gridDim.y;
  const i64 j0 = (i64)blockIdx.y * per;
  const i64 j1 = min(m, j0 + per);
  if (col >= ld) return;
  double acc[VN];
  double di[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    acc[e] = 0.0;
    di[e] = (dinv && col + e < N) ? dinv[col + e] : 0.0;
  }
  for (i64 j = j0; j < j1; j += UNROLL) {
    V c[UNROLL];
    double wa[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) {
      ok[t] = false;
      if (j + t < j1) {
        const i64 a = roi_idx[j + t];
        const i64 r = a - row0;
        // lower-block storage: only the kept part of row a (columns below its diagonal block's
        // end); the pairs (a, x) beyond it are read from row x by colsum_cols_kernel
        if (r >= 0 && r < n_loc && (sym_block == 0 || col < (a / sym_block + 1) * sym_block)) {
          ok[t] = true;
          c[t] = ld_stream(reinterpret_cast<const V*>(cov + r * ld + col));
          wa[t] = row_weight ? row_weight[j + t] : 1.0;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) {
      if (ok[t]) {
        if constexpr (VN == 2) {
          acc[0] += elem<RULE>((double)c[t].x, wa[t], di[0], p0, p1);
          acc[1] += elem<RULE>((double)c[t].y, wa[t], di[1], p0, p1);
        } else {
          acc[0] += elem<RULE>((double)c[t].x, wa[t], di[0], p0, p1);
          acc[1] += elem<RULE>((double)c[t].y, wa[t], di[1], p0, p1);
          acc[2] += elem<RULE>((double)c[t].z, wa[t], di[2], p0, p1);
          acc[3] += elem<RULE>((double)c[t].w, wa[t], di[3], p0, p1);
        }
      }
    }
  }
  double* o = part + (i64)blockIdx.y * N;
#pragma unroll
  for (int e = 0; e < VN; ++e)
    if (col + e < N) o[col + e] = acc[e];
}

__global__ void __launch_bounds__(256)
sum_parts_kernel(const double* __restrict__ part, int n_split, i64 N, double* __restrict__ out) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= N) return;
  double acc = 0.0;
  for (int s = 0; s < n_split; ++s) acc += part[(i64)s * N + x];  // fixed order: deterministic
  out[x] = acc;
}

// Same sums for the shard's own candidates through the symmetric entries
// cov[r][a]: one warp per local row, lanes stride over the ROI.
template <typename T, int RULE>
__global__ void __launch_bounds__(256)
colsum_cols_kernel(const T* __restrict__ cov, i64 ld, i64 n_loc, const i64* __restrict__ roi_idx,
                   i64 m, const double* __restrict__ row_weight, const double* __restrict__ dinv,
                   double p0, double p1, double* __restrict__ out, i64 sym_block, i64 row0,
                   int accumulate) {
  const i64 r = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_loc) return;
  const T* row = cov + r * ld;
  const double di = dinv ? dinv[r] : 0.0;
  // lower-block storage: row x = row0 + r holds the pairs (a, x) of the ROI points a whose own
  // kept part ends at or before x, i.e. a below the start of x's diagonal block
  const i64 a_end = sym_block > 0 ? ((row0 + r) / sym_block) * sym_block : (i64)0x7fffffffffffffffLL;
  double acc = 0.0;
  for (i64 j = lane; j < m; j += 32) {
    const i64 a = roi_idx[j];
    if (a < a_end) {
      const double v = (double)row[a];
      acc += elem<RULE>(v, row_weight ? row_weight[j] : 1.0, di, p0, p1);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) out[r] = accumulate ? out[r] + acc : acc;
}

// Kept-part row reduce for an ROI that covers a large part of the domain (lower-block storage): rows are walked
// directly (a = j, masked by the ROI's byte mask) instead of through the index list, the end of a row's kept part
// is a shift (the block size is 32), and a thread starts at the first row that keeps its column.  The gather form
// above spends ~100 integer instructions per 16-byte load on the index, a 64-bit division and the bounds and runs
// at 0.35 of the copy peak when m ~ N; this one streams.  grid (column tiles, row splits); row split s covers rows
// [s * per, (s + 1) * per); UNROLL rows in flight per thread.
template <typename T, int RULE, bool HAS_W, int VPT>
__global__ void __launch_bounds__(256)
colsum_rows_dense_kernel(const T* __restrict__ cov, i64 ld, i64 row0, i64 n_loc, i64 N,
                         const unsigned char* __restrict__ roi_mask, const double* __restrict__ w_pt,
                         const double* __restrict__ dinv, double p0, double p1, double* __restrict__ part, i64 per) {
  // a CTA covers VPT * 256 consecutive 16-byte vectors of a row (VPT * 4 KB contiguous): thread t owns vectors
  // t, t + 256, ...; UNROLL rows in flight
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  constexpr int UNROLL = VPT == 1 ? 8 : 2;
  const i64 tile0 = (i64)blockIdx.x * 256 * VPT * VN;
  double acc[VPT][VN], di[VPT][VN];
  i64 colv[VPT];
  i64 lo = 0x7fffffffffffffffLL;
#pragma unroll
  for (int p = 0; p < VPT; ++p) {
    colv[p] = tile0 + ((i64)p * 256 + threadIdx.x) * VN;
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      acc[p][e] = 0.0;
      di[p][e] = (dinv && colv[p] + e < N) ? dinv[colv[p] + e] : 0.0;
    }
    if (colv[p] < ld && colv[p] < lo) lo = colv[p];
  }
  if (lo >= ld) return;
  // local rows [j0, j1) of this split; row a = row0 + j keeps column c iff a >= (c / 32) * 32
  i64 j0 = (i64)blockIdx.y * per, j1 = j0 + per;
  if (j1 > n_loc) j1 = n_loc;
  const i64 first = ((lo >> 5) << 5) - row0;  // (the thread's lowest column: later ones start later, checked per row)
  if (first > j0) j0 = first;
  for (i64 j = j0; j < j1; j += UNROLL) {
    V c[UNROLL][VPT];
    bool ok[UNROLL][VPT];
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) {
      const bool row_ok = j + t < j1 && roi_mask[row0 + j + t] != 0;
      const i64 a32 = ((row0 + j + t) >> 5) << 5;
#pragma unroll
      for (int p = 0; p < VPT; ++p) {
        ok[t][p] = row_ok && colv[p] < ld && colv[p] < a32 + 32;
        if (ok[t][p]) c[t][p] = ld_stream(reinterpret_cast<const V*>(cov + (j + t) * ld + colv[p]));
      }
    }
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) {
      const double wa = (HAS_W && j + t < j1) ? w_pt[row0 + j + t] : 1.0;
#pragma unroll
      for (int p = 0; p < VPT; ++p) {
        if (ok[t][p]) {
          T e[VN];
          memcpy(e, &c[t][p], sizeof(V));
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[p][k] += elem<RULE>((double)e[k], wa, di[p][k], p0, p1);
        }
      }
    }
  }
  double* o = part + (i64)blockIdx.y * N;
#pragma unroll
  for (int p = 0; p < VPT; ++p)
#pragma unroll
    for (int e = 0; e < VN; ++e)
      if (colv[p] + e < N) o[colv[p] + e] = acc[p][e];
}

// The same mirrored sums when the ROI is a large part of the domain (potential-maximiser ROIs of the
// Safe-BO configurations: m ~ N): instead of gathering m indexed entries per row -- one dependent
// index + value load per lane and iteration, latency-bound -- the kept part of row x is STREAMED with
// 16-byte loads and masked by the ROI's byte mask (64 KB, L1-resident); by-point weights w_pt only for
// the rules that have them.  One warp per local row, UNROLL vectors in flight per lane.
template <typename T, int RULE, bool HAS_W>
__global__ void __launch_bounds__(256)
colsum_cols_dense_kernel(const T* __restrict__ cov, i64 ld, i64 n_loc, i64 N,
                         const unsigned char* __restrict__ roi_mask, const double* __restrict__ w_pt,
                         const double* __restrict__ dinv, double p0, double p1, double* __restrict__ out,
                         i64 sym_block, i64 row0, int accumulate) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  constexpr int UNROLL = 4;
  const i64 r = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_loc) return;
  const T* row = cov + r * ld;
  const double di = dinv ? dinv[r] : 0.0;
  i64 a_end = sym_block > 0 ? ((row0 + r) / sym_block) * sym_block : N;  // (a multiple of 32 / the row length)
  if (a_end > N) a_end = N;
  double acc = 0.0;
  for (i64 c0 = (i64)lane * VN; c0 < a_end; c0 += (i64)32 * VN * UNROLL) {
    V v[UNROLL];
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) {
      const i64 c = c0 + (i64)t * 32 * VN;
      if (c < a_end) v[t] = ld_stream(reinterpret_cast<const V*>(row + c));
    }
#pragma unroll
    for (int t = 0; t < UNROLL; ++t) {
      const i64 c = c0 + (i64)t * 32 * VN;
      if (c < a_end) {
        T e[VN];
        memcpy(e, &v[t], sizeof(V));
#pragma unroll
        for (int k = 0; k < VN; ++k)
          if (c + k < a_end && roi_mask[c + k])
            acc += elem<RULE>((double)e[k], HAS_W ? w_pt[c + k] : 1.0, di, p0, p1);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) out[r] = accumulate ? out[r] + acc : acc;
}

// ONE pass for a VTL-type step in lower-block storage with an ROI that covers a large part of the domain: the
// single pending rank-1 update (k, w of the last observation) is applied to every kept entry of one GP and BOTH
// halves of the next score are accumulated from the updated values on the way --
//   kept part of the ROI rows:   part[s][x]    = sum_{a in split s, a in ROI} elem(v'[a][x], w_a, dinv[x])
//   mirrored part of row r:      rowpart[t][r] = sum_{c in tile t, c in ROI, c < blockstart(r)} elem(v'[r][c], w_c, dinv[r])
// -- instead of an update pass (read + write) followed by two read passes (tal_rankb_update +
// tal_colsum_lower_dense): 2 instead of 4 transfers of the matrix per step.  The update is the same rounded
// operation as in rank1.cu / rankb.cu (sub_prod_m<FMA> in T): the matrix stays bit-identical.
// grid (column tiles of 256 vectors, row splits of `per` rows); dynamic shared memory: per * (2 + 8) doubles + per T.
template <typename T, int RULE, bool HAS_W, bool FMA>
__global__ void __launch_bounds__(256, 2)
update_colsum_dense_kernel(T* __restrict__ cov, i64 ld, i64 row0, i64 n_loc, i64 N, const T* __restrict__ kvec,
                           const T* __restrict__ wvec, const unsigned char* __restrict__ roi_mask,
                           const double* __restrict__ w_pt, const double* __restrict__ dinv, double p0, double p1,
                           double* __restrict__ part, double* __restrict__ rowpart, i64 per) {
  typedef typename Vec16<T>::type V;
  constexpr int VN = Vec16<T>::n;
  constexpr int UNROLL = 8;  // rows in flight per thread; their 8 row sums share ONE butterfly reduction
  extern __shared__ __align__(16) unsigned char ucd_smem[];
  double* wrow = reinterpret_cast<double*>(ucd_smem);  // [per] weight of row a as a TARGET
  double* drow = wrow + per;                            // [per] dinv of row r as a CANDIDATE
  double* racc = drow + per;                            // [per][8] mirrored sums per warp
  T* krow = reinterpret_cast<T*>(racc + per * 8);       // [per]
  unsigned char* mrow = reinterpret_cast<unsigned char*>(krow + per);  // [per]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  i64 j0 = (i64)blockIdx.y * per, j1 = j0 + per;
  if (j1 > n_loc) j1 = n_loc;
  const i64 tile0 = (i64)blockIdx.x * 256 * VN;
  // the whole tile lies above the diagonal blocks of every row of this split: nothing kept, nothing to add
  const bool tile_dead = tile0 >= (((row0 + j1 - 1) >> 5) + 1) << 5;
  for (i64 j = j0 + threadIdx.x; j < j1; j += 256) {
    const i64 a = row0 + j;
    mrow[j - j0] = roi_mask[a];
    wrow[j - j0] = HAS_W ? w_pt[a] : 1.0;
    drow[j - j0] = dinv ? dinv[a] : 0.0;
    krow[j - j0] = kvec[a];
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) racc[(j - j0) * 8 + w8] = 0.0;
  }
  __syncthreads();
  const i64 col = tile0 + (i64)threadIdx.x * VN;
  double acc[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) acc[e] = 0.0;
  if (!tile_dead) {
    T wc[VN];
    double dic[VN], wpc[VN];
    bool mc[VN];
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const bool in = col + e < N;
      wc[e] = col + e < ld ? wvec[col + e] : T(0);
      dic[e] = (dinv && in) ? dinv[col + e] : 0.0;
      mc[e] = in && roi_mask[col + e] != 0;
      wpc[e] = (HAS_W && in) ? w_pt[col + e] : 1.0;
    }
    // warp-uniform row range: the first row that keeps the warp's first column
    const i64 wcol0 = tile0 + (i64)warp * 32 * VN;
    i64 js = ((wcol0 >> 5) << 5) - row0;
    if (js < j0) js = j0;
    for (i64 j = js; j < j1; j += UNROLL) {
      V c[UNROLL];
      bool ok[UNROLL];
#pragma unroll
      for (int t = 0; t < UNROLL; ++t) {
        const i64 a = row0 + j + t;
        ok[t] = j + t < j1 && col < ld && col < ((a >> 5) + 1) << 5;
        if (ok[t]) c[t] = ld_stream(reinterpret_cast<const V*>(cov + (j + t) * ld + col));
      }
      double rs[UNROLL];
#pragma unroll
      for (int t = 0; t < UNROLL; ++t) {
        rs[t] = 0.0;
        if (ok[t]) {
          const i64 jl = j + t - j0;
          const i64 a = row0 + j + t;
          T e[VN];
          memcpy(e, &c[t], sizeof(V));
          const T kr = krow[jl];
          const bool ma = mrow[jl] != 0;
          const double wa = wrow[jl], da = drow[jl];
          const i64 bstart = (a >> 5) << 5;  // columns before it are the mirrored pairs of row a
#pragma unroll
          for (int k = 0; k < VN; ++k) {
            e[k] = sub_prod_m<FMA>(e[k], kr, wc[k]);
            const double v = (double)e[k];
            if (ma) acc[k] += elem<RULE>(v, wa, dic[k], p0, p1);
            if (mc[k] && col + k < bstart) rs[t] += elem<RULE>(v, wpc[k], da, p0, p1);
          }
          V o;
          memcpy(&o, e, sizeof(V));
          st_stream(reinterpret_cast<V*>(cov + (j + t) * ld + col), o);
        }
      }
      // 8 row sums over 32 lanes in 9 exchanges: halve the set of rows a lane carries at every step (lane bits
      // 4, 3, 2 pick the row), then finish over the lanes that differ in bits 1, 0.  Fixed order: deterministic.
      double k4[4], k2[2], k1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double send = (lane & 16) ? rs[i] : rs[i + 4];
        const double recv = __shfl_xor_sync(0xffffffffu, send, 16);
        k4[i] = ((lane & 16) ? rs[i + 4] : rs[i]) + recv;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double send = (lane & 8) ? k4[i] : k4[i + 2];
        const double recv = __shfl_xor_sync(0xffffffffu, send, 8);
        k2[i] = ((lane & 8) ? k4[i + 2] : k4[i]) + recv;
      }
      {
        const double send = (lane & 4) ? k2[0] : k2[1];
        const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
        k1 = ((lane & 4) ? k2[1] : k2[0]) + recv;
      }
      k1 += __shfl_xor_sync(0xffffffffu, k1, 2);
      k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
      const int trow = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      if ((lane & 3) == 0 && j + trow < j1) racc[(j + trow - j0) * 8 + warp] = k1;
    }
  }
  double* o = part + (i64)blockIdx.y * N;
#pragma unroll
  for (int e = 0; e < VN; ++e)
    if (col + e < N) o[col + e] = acc[e];
  __syncthreads();
  for (i64 j = j0 + threadIdx.x; j < j1; j += 256) {
    const double* r8 = racc + (j - j0) * 8;
    double sum = 0.0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) sum += r8[w8];  // fixed order
    rowpart[(i64)blockIdx.x * n_loc + j] = sum;
  }
}

// c[x] = sum_s part[s][x] + (x in [row0, row0 + n_loc): sum_t rowpart[t][x - row0]), fixed order
__global__ void __launch_bounds__(256)
sum_parts2_kernel(const double* __restrict__ part, int n_split, const double* __restrict__ rowpart, int n_tiles,
                  i64 row0, i64 n_loc, i64 N, double* __restrict__ out) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= N) return;
  double acc = 0.0;
  for (int s = 0; s < n_split; ++s) acc += part[(i64)s * N + x];
  const i64 r = x - row0;
  if (r >= 0 && r < n_loc)
    for (int t = 0; t < n_tiles; ++t) acc += rowpart[(i64)t * n_loc + r];
  out[x] = acc;
}

// tr_A = sum_{a in roi} var[a]  (one block, deterministic)
__global__ void __launch_bounds__(256)
roi_trace_kernel(const double* __restrict__ var, const i64* __restrict__ roi_idx, i64 m,
                 double* __restrict__ out) {
  __shared__ double sh[33];
  double acc = 0.0;
  for (i64 j = threadIdx.x; j < m; j += blockDim.x) acc += var[roi_idx[j]];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) *out = acc;
}

// w_a for the CTL / MM-ITL (1/var[a]) and TruVar (var[a]) rules
__global__ void __launch_bounds__(256)
roi_weights_kernel(const double* __restrict__ var, const i64* __restrict__ roi_idx, i64 m,
                   int reciprocal, double* __restrict__ out) {
  const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const double v = var[roi_idx[j]];
  out[j] = reciprocal ? 1.0 / v : v;
}

// dinv[x] = 1 / (var[x] + rho[x]^2)
__global__ void __launch_bounds__(256)
pred_var_inv_kernel(const double* __restrict__ var, const double* __restrict__ rho, i64 N,
                    double* __restrict__ out) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= N) return;
  const double r = rho[x];
  out[x] = 1.0 / (var[x] + r * r);
}

__global__ void __launch_bounds__(256)
rows_finish_kernel(int rule, const double* __restrict__ c, const double* __restrict__ var,
                   const double* __restrict__ rho, const double* __restrict__ tr, i64 n,
                   int accumulate, double* __restrict__ F) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  double f;
  if (rule == TAL_RULE_VTL) {
    const double r = rho[x];
    f = -(*tr - c[x] / (var[x] + r * r));  // vtl.py:29-41, summed over the ROI
  } else if (rule == TAL_RULE_CTL) {
    const double r = rho[x];
    f = c[x] / (var[x] + r * r);  // marginal.py:150-153 summed over the ROI
  } else if (rule == TAL_RULE_MMITL) {
    f = c[x];
  } else {
    f = -c[x];  // tru_var/__init__.py:50
  }
  F[x] = accumulate ? F[x] + f : f;
}

__global__ void __launch_bounds__(256)
itl_directed_kernel(const double* __restrict__ var, const double* __restrict__ cs,
                    const double* __restrict__ rho, i64 n, int accumulate, double* __restrict__ F) {
  const i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const double r = rho[x];
  const double nz = r * r;
  // diag(cov - Kxt^T Sigma^-1 Kxt) = var - ||L^-1 Kxt[:, x]||^2 (gaussian_distribution.py:36)
  const double post_var = __dsub_rn(var[x], cs[x]);
  // 0.5 * max(log(predictive / posterior_predictive), 0)  (itl.py:53-58);
  // jnp.maximum propagates NaN
  const double f = 0.5 * max_nan(log((var[x] + nz) / (post_var + nz)), 0.0);
  F[x] = accumulate ? F[x] + f : f;
}

template <typename T, int RULE>
static int launch_colsum_rows(const T* cov, i64 ld, i64 row0, i64 n_loc, i64 N, const i64* roi_idx,
                              i64 m, const double* row_weight, const double* dinv, double p0,
                              double p1, double* part, int n_split, cudaStream_t st, i64 sym_block) {
  constexpr int VN = Vec16<T>::n;
  dim3 grid((unsigned)((ld + 256 * VN - 1) / (256 * VN)), (unsigned)n_split);
  colsum_rows_kernel<T, RULE, 4><<<grid, 256, 0, st>>>(cov, ld, row0, n_loc, N, roi_idx, m,
                                                       row_weight, dinv, p0, p1, part, sym_block);
  return 0;
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_score_itl_undirected(const double* var, const double* rho, int q, long long N, double* F,
                             void* stream) {
  TAL_ARG(var && rho && F && q >= 1 && N >= 1);
  itl_undirected_kernel<<<(unsigned)((N + 255) / 256), 256, 0, as_stream(stream)>>>(var, rho, q, N,
                                                                                   F);
  TAL_LAUNCH_CHECK("tal_score_itl_undirected");
  return TAL_OK;
}

static int colsum_rows_impl(int rule, const void* cov, long long ld, long long row0, long long n_loc,
                    long long N, const long long* roi_idx, long long m, const double* row_weight,
                    const double* dinv, double p0, double p1, double* part, int n_split,
                    double* c_out, int dtype, void* stream, i64 sym_block) {
  TAL_ARG(cov && roi_idx && part && c_out && m >= 0 && n_split >= 1 && n_split <= 65535);
  TAL_ARG(ld % 32 == 0 && ld >= N);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG(rule >= TAL_RULE_VTL && rule <= TAL_RULE_TRUVAR);
  cudaStream_t st = as_stream(stream);
  double* dst = n_split == 1 ? c_out : part;
#define TAL_ROWS(T, R)                                                                          \
  launch_colsum_rows<T, R>((const T*)cov, ld, row0, n_loc, N, roi_idx, m, row_weight, dinv, p0, \
                           p1, dst, n_split, st, sym_block)
  if (dtype == TAL_F64) {
    switch (rule) {
      case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_ROWS(double, TAL_RULE_VTL); break;
      case TAL_RULE_MMITL: TAL_ROWS(double, TAL_RULE_MMITL); break;
      default: TAL_ROWS(double, TAL_RULE_TRUVAR); break;
    }
  } else {
    switch (rule) {
      case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_ROWS(float, TAL_RULE_VTL); break;
      case TAL_RULE_MMITL: TAL_ROWS(float, TAL_RULE_MMITL); break;
      default: TAL_ROWS(float, TAL_RULE_TRUVAR); break;
    }
  }
#undef TAL_ROWS
  TAL_LAUNCH_CHECK("tal_colsum_rows");
  if (n_split > 1) {
    sum_parts_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(part, n_split, N, c_out);
    TAL_LAUNCH_CHECK("tal_colsum_rows/sum_parts");
  }
  return TAL_OK;
}

int tal_colsum_rows(int rule, const void* cov, long long ld, long long row0, long long n_loc,
                    long long N, const long long* roi_idx, long long m, const double* row_weight,
                    const double* dinv, double p0, double p1, double* part, int n_split,
                    double* c_out, int dtype, void* stream) {
  return colsum_rows_impl(rule, cov, ld, row0, n_loc, N, roi_idx, m, row_weight, dinv, p0, p1, part,
                          n_split, c_out, dtype, stream, 0);
}

static int colsum_cols_impl(int rule, const void* cov, long long ld, long long n_loc,
                    const long long* roi_idx, long long m, const double* row_weight,
                    const double* dinv, double p0, double p1, double* c_out, int dtype,
                    void* stream, i64 sym_block, i64 row0, int accumulate) {
  TAL_ARG(cov && roi_idx && c_out && m >= 0 && n_loc >= 0);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG(rule >= TAL_RULE_VTL && rule <= TAL_RULE_TRUVAR);
  if (n_loc == 0) return TAL_OK;
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)((n_loc + 7) / 8);
#define TAL_COLS(T, R)                                                                        \
  colsum_cols_kernel<T, R><<<grid, 256, 0, st>>>((const T*)cov, ld, n_loc, roi_idx, m,        \
                                                 row_weight, dinv, p0, p1, c_out, sym_block,  \
                                                 row0, accumulate)
  if (dtype == TAL_F64) {
    switch (rule) {
      case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_COLS(double, TAL_RULE_VTL); break;
      case TAL_RULE_MMITL: TAL_COLS(double, TAL_RULE_MMITL); break;
      default: TAL_COLS(double, TAL_RULE_TRUVAR); break;
    }
  } else {
    switch (rule) {
      case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_COLS(float, TAL_RULE_VTL); break;
      case TAL_RULE_MMITL: TAL_COLS(float, TAL_RULE_MMITL); break;
      default: TAL_COLS(float, TAL_RULE_TRUVAR); break;
    }
  }
#undef TAL_COLS
  TAL_LAUNCH_CHECK("tal_colsum_cols");
  return TAL_OK;
}

int tal_colsum_cols(int rule, const void* cov, long long ld, long long n_loc,
                    const long long* roi_idx, long long m, const double* row_weight,
                    const double* dinv, double p0, double p1, double* c_out, int dtype,
                    void* stream) {
  return colsum_cols_impl(rule, cov, ld, n_loc, roi_idx, m, row_weight, dinv, p0, p1, c_out, dtype,
                          stream, 0, 0, 0);
}

/* lower-block storage: c[x] = sum_{a in ROI} elem(cov[a][x]) over ALL x in [0, N), from what this
 * shard keeps: the kept part of its ROI rows (row reduce) plus, for its own rows x, the ROI entries
 * mirrored into row x (column gather, accumulated).  Every pair (a, x) is counted exactly once
 * across the shards; a sum over the shards completes c. */
int tal_colsum_lower(int rule, const void* cov, long long ld, long long row0, long long n_loc,
                     long long N, const long long* roi_idx, long long m, const double* row_weight,
                     const double* dinv, double p0, double p1, double* part, int n_split,
                     double* c_out, int dtype, void* stream) {
  const i64 blk = tal_sym_block(dtype);
  int rc = colsum_rows_impl(rule, cov, ld, row0, n_loc, N, roi_idx, m, row_weight, dinv, p0, p1, part,
                            n_split, c_out, dtype, stream, blk);
  if (rc != TAL_OK) return rc;
  return colsum_cols_impl(rule, cov, ld, n_loc, roi_idx, m, row_weight, dinv ? dinv + row0 : nullptr, p0,
                          p1, c_out + row0, dtype, stream, blk, row0, 1);
}

/* tal_colsum_lower for an ROI that covers a large part of the domain: the mirrored part is streamed and
 * masked (roi_mask: one byte per point; w_pt: by-point weights [N] or NULL = 1) instead of gathered. */
int tal_colsum_lower_dense(int rule, const void* cov, long long ld, long long row0, long long n_loc,
                           long long N, const long long* roi_idx, long long m, const unsigned char* roi_mask,
                           const double* row_weight, const double* w_pt, const double* dinv, double p0,
                           double p1, double* part, int n_split, double* c_out, int dtype, void* stream) {
  TAL_ARG(roi_mask && ((row_weight == nullptr) == (w_pt == nullptr)));
  TAL_ARG(cov && part && c_out && n_split >= 1 && n_split <= 65535 && ld % 32 == 0 && ld >= N);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG(rule >= TAL_RULE_VTL && rule <= TAL_RULE_TRUVAR);
  TAL_ARG(tal_sym_block(dtype) == 32);  // (the kernels below shift)
  (void)roi_idx; (void)m;
  const i64 blk = tal_sym_block(dtype);
  cudaStream_t st = as_stream(stream);
  {  // kept part of the ROI rows this shard holds: part[s][x], s < n_split, then a fixed-order sum
    const i64 per = (n_loc + n_split - 1) / n_split > 0 ? (n_loc + n_split - 1) / n_split : 1;
    const i64 vn = dtype == TAL_F64 ? 2 : 4;
    static int vpt = -1;  // 1 (default): 4 KB of a row per CTA, 1.27 ms per fp32 matrix at C4 = the copy peak;
                          // TAL_ROWS_DENSE_VPT=4: 16 KB per CTA (measured 1.69 ms: fewer, fatter CTAs)
    if (vpt < 0) {
      const char* e = getenv("TAL_ROWS_DENSE_VPT");
      vpt = (e && atoi(e) == 4) ? 4 : 1;
    }
    dim3 grid((unsigned)((ld + 256 * vn * vpt - 1) / (256 * vn * vpt)), (unsigned)n_split);
    double* dst = n_split == 1 ? c_out : part;
#define TAL_RD2(T, R, W, P)                                                                                    \
  colsum_rows_dense_kernel<T, R, W, P><<<grid, 256, 0, st>>>((const T*)cov, ld, row0, n_loc, N, roi_mask, w_pt, \
                                                              dinv, p0, p1, dst, per)
#define TAL_RD(T, R)                                                                                         \
  do {                                                                                                       \
    if (vpt == 1) { if (w_pt) TAL_RD2(T, R, true, 1); else TAL_RD2(T, R, false, 1); }                        \
    else { if (w_pt) TAL_RD2(T, R, true, 4); else TAL_RD2(T, R, false, 4); }                                 \
  } while (0)
    if (dtype == TAL_F64) {
      switch (rule) {
        case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_RD(double, TAL_RULE_VTL); break;
        case TAL_RULE_MMITL: TAL_RD(double, TAL_RULE_MMITL); break;
        default: TAL_RD(double, TAL_RULE_TRUVAR); break;
      }
    } else {
      switch (rule) {
        case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_RD(float, TAL_RULE_VTL); break;
        case TAL_RULE_MMITL: TAL_RD(float, TAL_RULE_MMITL); break;
        default: TAL_RD(float, TAL_RULE_TRUVAR); break;
      }
    }
#undef TAL_RD2
#undef TAL_RD
    TAL_LAUNCH_CHECK("tal_colsum_lower_dense/rows");
    if (n_split > 1) {
      sum_parts_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(part, n_split, N, c_out);
      TAL_LAUNCH_CHECK("tal_colsum_lower_dense/sum_parts");
    }
  }
  if (n_loc == 0) return TAL_OK;
  const unsigned grid = (unsigned)((n_loc + 7) / 8);
  const double* dv = dinv ? dinv + row0 : nullptr;
#define TAL_CD(T, R)                                                                                         \
  do {                                                                                                       \
    if (w_pt) colsum_cols_dense_kernel<T, R, true><<<grid, 256, 0, st>>>((const T*)cov, ld, n_loc, N, roi_mask, \
        w_pt, dv, p0, p1, c_out + row0, blk, row0, 1);                                                        \
    else colsum_cols_dense_kernel<T, R, false><<<grid, 256, 0, st>>>((const T*)cov, ld, n_loc, N, roi_mask,    \
        w_pt, dv, p0, p1, c_out + row0, blk, row0, 1);                                                        \
  } while (0)
  if (dtype == TAL_F64) {
    switch (rule) {
      case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_CD(double, TAL_RULE_VTL); break;
      case TAL_RULE_MMITL: TAL_CD(double, TAL_RULE_MMITL); break;
      default: TAL_CD(double, TAL_RULE_TRUVAR); break;
    }
  } else {
    switch (rule) {
      case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_CD(float, TAL_RULE_VTL); break;
      case TAL_RULE_MMITL: TAL_CD(float, TAL_RULE_MMITL); break;
      default: TAL_CD(float, TAL_RULE_TRUVAR); break;
    }
  }
#undef TAL_CD
  TAL_LAUNCH_CHECK("tal_colsum_lower_dense");
  return TAL_OK;
}

long long tal_update_colsum_scratch_bytes(long long ld, long long n_loc, int dtype) {
  const long long vn = base_dtype(dtype) == TAL_F64 ? 2 : 4;
  const long long tiles = (ld + 256 * vn - 1) / (256 * vn);
  return tiles * (n_loc > 0 ? n_loc : 1) * (long long)sizeof(double);
}

int tal_update_colsum_lower_dense(int rule, void* cov, long long ld, long long row0, long long n_loc, long long N,
                                  const void* kvec, const void* wvec, const unsigned char* roi_mask,
                                  const double* w_pt, const double* dinv, double p0, double p1, double* part,
                                  int n_split, double* rowpart, double* c_out, int dtype, void* stream) {
  TAL_ARG(cov && kvec && wvec && roi_mask && part && rowpart && c_out);
  TAL_ARG(n_split >= 1 && n_split <= 65535 && ld % 32 == 0 && ld >= N && n_loc >= 0 && row0 % 32 == 0);
  TAL_ARG(rule >= TAL_RULE_VTL && rule <= TAL_RULE_TRUVAR);
  const bool fma_mode = arith_fma(dtype);
  dtype = base_dtype(dtype);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG(tal_sym_block(dtype) == 32);
  cudaStream_t st = as_stream(stream);
  const i64 vn = dtype == TAL_F64 ? 2 : 4;
  const i64 s_el = dtype == TAL_F64 ? 8 : 4;
  const i64 per = (n_loc + n_split - 1) / n_split > 0 ? (n_loc + n_split - 1) / n_split : 1;
  const int tiles = (int)((ld + 256 * vn - 1) / (256 * vn));
  const size_t smem = (size_t)per * (10 * sizeof(double) + s_el + 1) + 16;
  TAL_ARG(smem <= 96 * 1024);
  dim3 grid((unsigned)tiles, (unsigned)n_split);
#define TAL_UC3(T, R, W, F)                                                                                   \
  do {                                                                                                       \
    auto kern = update_colsum_dense_kernel<T, R, W, F>;                                                      \
    static bool attr = false;                                                                                \
    if (!attr) {                                                                                             \
      TAL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));          \
      attr = true;                                                                                           \
    }                                                                                                        \
    kern<<<grid, 256, smem, st>>>((T*)cov, ld, row0, n_loc, N, (const T*)kvec, (const T*)wvec, roi_mask, w_pt, dinv, \
                                  p0, p1, part, rowpart, per);                                               \
  } while (0)
#define TAL_UC2(T, R, W) do { if (fma_mode) TAL_UC3(T, R, W, true); else TAL_UC3(T, R, W, false); } while (0)
#define TAL_UC(T, R) do { if (w_pt) TAL_UC2(T, R, true); else TAL_UC2(T, R, false); } while (0)
  if (n_loc > 0) {
    if (dtype == TAL_F64) {
      switch (rule) {
        case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_UC(double, TAL_RULE_VTL); break;
        case TAL_RULE_MMITL: TAL_UC(double, TAL_RULE_MMITL); break;
        default: TAL_UC(double, TAL_RULE_TRUVAR); break;
      }
    } else {
      switch (rule) {
        case TAL_RULE_VTL: case TAL_RULE_CTL: TAL_UC(float, TAL_RULE_VTL); break;
        case TAL_RULE_MMITL: TAL_UC(float, TAL_RULE_MMITL); break;
        default: TAL_UC(float, TAL_RULE_TRUVAR); break;
      }
    }
    TAL_LAUNCH_CHECK("tal_update_colsum_lower_dense");
  }
#undef TAL_UC
#undef TAL_UC2
#undef TAL_UC3
  sum_parts2_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(part, n_loc > 0 ? n_split : 0, rowpart, tiles, row0,
                                                                 n_loc, N, c_out);
  TAL_LAUNCH_CHECK("tal_update_colsum_lower_dense/sum");
  return TAL_OK;
}

int tal_roi_trace(const double* var, const long long* roi_idx, long long m, double* out,
                  void* stream) {
  TAL_ARG(var && roi_idx && out && m >= 0);
  roi_trace_kernel<<<1, 256, 0, as_stream(stream)>>>(var, roi_idx, m, out);
  TAL_LAUNCH_CHECK("tal_roi_trace");
  return TAL_OK;
}

int tal_roi_weights(const double* var, const long long* roi_idx, long long m, int reciprocal,
                    double* out, void* stream) {
  TAL_ARG(var && roi_idx && out && m >= 0);
  if (m == 0) return TAL_OK;
  roi_weights_kernel<<<(unsigned)((m + 255) / 256), 256, 0, as_stream(stream)>>>(var, roi_idx, m,
                                                                                reciprocal, out);
  TAL_LAUNCH_CHECK("tal_roi_weights");
  return TAL_OK;
}

int tal_pred_var_inv(const double* var, const double* rho, long long N, double* out,
                     void* stream) {
  TAL_ARG(var && rho && out && N >= 1);
  pred_var_inv_kernel<<<(unsigned)((N + 255) / 256), 256, 0, as_stream(stream)>>>(var, rho, N, out);
  TAL_LAUNCH_CHECK("tal_pred_var_inv");
  return TAL_OK;
}

int tal_score_rows_finish(int rule, const double* c, const double* var, const double* rho,
                          const double* tr, long long n, int accumulate, double* F,
                          void* stream) {
  TAL_ARG(c && F && n >= 0);
  TAL_ARG(rule >= TAL_RULE_VTL && rule <= TAL_RULE_TRUVAR);
  TAL_ARG(rule > TAL_RULE_CTL || (var && rho));
  TAL_ARG(rule != TAL_RULE_VTL || tr);
  if (n == 0) return TAL_OK;
  rows_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      rule, c, var, rho, tr, n, accumulate, F);
  TAL_LAUNCH_CHECK("tal_score_rows_finish");
  return TAL_OK;
}

int tal_score_itl_directed(const double* var, const double* cs, const double* rho,
                           long long n, int accumulate, double* F, void* stream) {
  TAL_ARG(var && cs && rho && F && n >= 0);
  if (n == 0) return TAL_OK;
  itl_directed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      var, cs, rho, n, accumulate, F);
  TAL_LAUNCH_CHECK("tal_score_itl_directed");
  return TAL_OK;
}

}  // extern "C"
