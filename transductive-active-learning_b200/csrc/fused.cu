// One BO step of directed ITL over a resident conditioning state in THREE kernels.
//
// Reference path: DiscreteGreedyAlgorithm._step (lib/algorithms/__init__.py:77-94) with
// ITL._directed (lib/algorithms/itl.py:33-60) as F and MarginalModel.step
// (lib/model/marginal.py:350-368) as the update.  The unfused sequence (stepctx.cu, round 1)
// was ~12 launches of 3-9 us kernels around the one streaming pass; at 8 GPUs that chain was
// 40 % of a step.  Here the O(N) work is folded into the kernels that already touch the data:
//
//   fused_observe   row idx of every GP (own rows / mirrored entries, or the peer mailbox after
//                   waiting for the G pieces), pending-update correction, w = (k/L)/L, mean, var,
//                   l, u, and the (k, w) pair queued for the deferred pass    [was 5 launches]
//   fused_partial   the read-only pass over the conditioning factor; its prologue gathers the
//                   column Wb[:, idx] itself (or reads the peer-supplied one) [was 3 launches]
//   fused_finish    appended rows g, h; cs += g^2 - k w; the directed-ITL score of every
//                   candidate; sample-region + safe-set masks; block / last-block arg-max;
//                   sharded: the last block publishes the record to every peer and tells them
//                   that this rank's row buffer is free                        [was 4-6 launches]
//
// Every value is computed by the same sequence of rounded operations as in the unfused kernels
// (rank1.cu step_vectors, rankb.cu pending_correct, condapp.cu, score.cu itl_directed, argmax.cu).
#include "argmax.cuh"
#include "fused.cuh"
#include "peer.cuh"

namespace tal {

__device__ __forceinline__ i64 fz_cend(i64 grow, i64 block, i64 ld) {
  if (block <= 0) return ld;
  const i64 e = (grow / block + 1) * block;
  return e < ld ? e : ld;
}

// grid (ceil(ld / 256), q).  G == 1: row0 = 0, n_loc = N (this GPU holds every row).
template <typename T, bool FMA>
__global__ void __launch_bounds__(256) fused_observe_kernel(FusedObs a) {
  __shared__ T kx_p[TAL_MAX_PENDING], wx_p[TAL_MAX_PENDING];
  __shared__ T kx_sh;
  const int gp = blockIdx.y;
  const i64 ld = a.ld, N = a.N;
  const i64 b = (i64)blockIdx.x * 256 + threadIdx.x;
  T* kbuf = (T*)a.kbuf + (i64)gp * ld;
  T* wbuf = (T*)a.wbuf + (i64)gp * ld;
  T* Kslot = a.Kslot ? (T*)a.Kslot + (i64)gp * ld : nullptr;
  T* Wslot = a.Wslot ? (T*)a.Wslot + (i64)gp * ld : nullptr;
  if (a.mailbox && a.wait) {  // sharded: the G pieces of the row land in kbuf (inside this rank's mailbox)
    PeerHeader* h = hdr(a.mailbox);
    const unsigned long long want = *(const volatile unsigned long long*)&h->row_step + 1;
    if ((int)threadIdx.x < a.G && !spin_until(&h->row_seq[threadIdx.x], want)) h->error = TAL_PEER_ERR_ROW_TIMEOUT;
    __syncthreads();
  }
  const i64 idx = a.idx_dev ? __ldcg(a.idx_dev) : a.idx_host;
  if (idx < 0 || idx >= N) {  // failed arg-max: the observation is a no-op (k = w = 0 is queued)
    if (b < ld) {
      if (!a.mailbox) kbuf[b] = T(0);
      wbuf[b] = T(0);
      if (Kslot) { Kslot[b] = T(0); Wslot[b] = T(0); }
    }
    return;
  }
  const T* cov = (const T*)a.cov + (i64)gp * a.stride_q;
  const T* Kp = (const T*)a.Kp + (i64)gp * ld;
  const T* Wp = (const T*)a.Wp + (i64)gp * ld;
  if ((int)threadIdx.x < a.pending) {
    kx_p[threadIdx.x] = Kp[(i64)threadIdx.x * a.pend_stride + idx];
    wx_p[threadIdx.x] = Wp[(i64)threadIdx.x * a.pend_stride + idx];
  }
  T kx_raw = T(0);
  // the stored diagonal element: local matrix, or -- sharded -- the copy its owner pushed next to the row (NOT
  // kbuf[idx]: the block that owns element idx rewrites it in place below, a late block would correct it twice)
  if (threadIdx.x == 0) kx_raw = a.mailbox ? T(hdr(a.mailbox)->kdiag[gp]) : cov[idx * ld + idx];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.pending; ++i) kx_raw = sub_prod_m<FMA>(kx_raw, kx_p[i], wx_p[i]);
    kx_sh = kx_raw;
  }
  __syncthreads();
  if (b >= ld) return;
  const i64 cend = fz_cend(idx, a.sym_block, ld);
  const bool mirrored = b >= cend;
  T kb;
  if (a.mailbox) kb = kbuf[b];
  else if (!mirrored) kb = cov[idx * ld + b];
  else kb = b < N ? cov[b * ld + idx] : T(0);
  for (int i = 0; i < a.pending; ++i)
    kb = mirrored ? sub_prod_m<FMA>(kb, Kp[(i64)i * a.pend_stride + b], wx_p[i])
                  : sub_prod_m<FMA>(kb, kx_p[i], Wp[(i64)i * a.pend_stride + b]);
  // Sigma_oo = cov[idx, idx] + noise_std^2 (gaussian_distribution.py:29-32), in T like the matrix
  const double r = a.rho[(i64)gp * N + idx];
  const T s = T((double)kx_sh + r * r);
  const T L = sqrt(s);
  const T wb = (kb / L) / L;  // solve(L.T, solve(L, k)) for 1x1 L (utils.py:23)
  kbuf[b] = kb;
  wbuf[b] = wb;
  if (Kslot) { Kslot[b] = kb; Wslot[b] = wb; }
  if (b >= N) return;
  const i64 o = (i64)gp * N + b;
  const T v_new = sub_prod_m<FMA>(T(a.var[o]), kb, wb);  // == diag of the matrix update
  a.var[o] = (double)v_new;
  if (!a.y) {  // the observation arrives later (fused_observe_y): keep mean[idx] of before it
    if (b == 0) a.scal[gp] = a.mean[(i64)gp * N + idx];
    return;
  }
  const double yv = a.y_gather ? a.y[(i64)gp * a.y_stride + idx] : a.y[(i64)gp * a.y_stride];
  const double delta = yv - a.scal[gp];  // obs - prior_mean[obs_idx] (:120); scal = mean[idx] before this step
  const double m_new = add_prod(a.mean[o], (double)wb, delta);
  a.mean[o] = m_new;
  const double sd = sqrt((double)v_new);  // NaN for negative variance, as the reference
  const double bt = a.beta.v[gp];
  a.l[o] = max_nan(a.l[o], __dsub_rn(m_new, __dmul_rn(bt, sd)));
  a.u[o] = min_nan(a.u[o], __dadd_rn(m_new, __dmul_rn(bt, sd)));
}

// mean, l, u of an observation whose k, w, var were already applied (fused_observe with y == null).
template <typename T>
__global__ void __launch_bounds__(256)
fused_observe_y_kernel(const T* __restrict__ wbuf, i64 ld, i64 N, const i64* __restrict__ idx_dev, i64 idx_host,
                       const double* __restrict__ y, i64 y_stride, int y_gather, BetaQ beta,
                       const double* __restrict__ scal, const double* __restrict__ var, double* __restrict__ mean,
                       double* __restrict__ l, double* __restrict__ u) {
  const int gp = blockIdx.y;
  const i64 b = (i64)blockIdx.x * 256 + threadIdx.x;
  const i64 idx = idx_dev ? *idx_dev : idx_host;
  if (b >= N || idx < 0 || idx >= N) return;
  const double yv = y_gather ? y[(i64)gp * y_stride + idx] : y[(i64)gp * y_stride];
  const double delta = yv - scal[gp];
  const i64 o = (i64)gp * N + b;
  const double m_new = add_prod(mean[o], (double)wbuf[(i64)gp * ld + b], delta);
  mean[o] = m_new;
  const double sd = sqrt(var[o]);
  const double bt = beta.v[gp];
  l[o] = max_nan(l[o], __dsub_rn(m_new, __dmul_rn(bt, sd)));
  u[o] = min_nan(u[o], __dadd_rn(m_new, __dmul_rn(bt, sd)));
}

__device__ __forceinline__ double fz_sign(i64 j, i64 m_pad) {
  return (j < m_pad || ((j - m_pad) & 1) == 0) ? -1.0 : 1.0;  // W_0 and g rows: -, h rows: + (condapp.cu)
}

// part[gp][s][x] = sum_{j in split s} sign_j col_j Wb[gp][j][x], colsq[gp][s] = sum_j sign_j col_j^2 with
// col = Wb[gp][:, idx - cand0] (gathered here) or the peer-supplied column.
// grid (ceil(ncols / 512), n_split, q); dynamic shared memory: `per` doubles.
__global__ void __launch_bounds__(256)
fused_partial_kernel(const double* __restrict__ Wb, i64 wb_stride_q, i64 ldw, i64 rows, i64 m_pad,
                     const i64* __restrict__ idx_dev, i64 idx_host, i64 cand0, i64 n_cand,
                     const double* __restrict__ pcol, i64 pcol_stride, i64 ncols, double* __restrict__ part,
                     double* __restrict__ colsq, int n_split, i64 per, void* wait_box, int G) {
  extern __shared__ double csm[];
  __shared__ double sh[33];
  const int gp = blockIdx.z;
  if (wait_box) {  // running beside fused_observe: wait for the pieces of this row exchange here as well
    PeerHeader* h = hdr(wait_box);
    // (row_step through L2: this kernel runs on a side stream beside other grids; a line of the previous step
    //  left in an L1 must not decide which exchange is waited for)
    const unsigned long long want = *(const volatile unsigned long long*)&h->row_step + 1;
    if ((int)threadIdx.x < G && !spin_until(&h->row_seq[threadIdx.x], want)) h->error = TAL_PEER_ERR_ROW_TIMEOUT;
    __syncthreads();
  }
  const i64 j0 = (i64)blockIdx.y * per;
  const i64 j1 = j0 + per < rows ? j0 + per : rows;
  const double* Wg = Wb + (i64)gp * wb_stride_q;
  const i64 xl = (idx_dev ? __ldcg(idx_dev) : idx_host) - cand0;  // (L2 reads: see above)
  const bool have = xl >= 0 && xl < n_cand;
  double sq = 0.0;
  for (i64 j = j0 + threadIdx.x; j < j1; j += 256) {
    const double c = pcol ? __ldcg(pcol + (i64)gp * pcol_stride + j) : (have ? __ldcg(Wg + j * ldw + xl) : 0.0);
    const double sc = fz_sign(j, m_pad) * c;
    csm[j - j0] = sc;
    sq += sc * c;
  }
  if (blockIdx.x == 0) {
    sq = block_sum(sq, sh);
    if (threadIdx.x == 0) colsq[(i64)gp * n_split + blockIdx.y] = sq;
  }
  __syncthreads();
  const i64 x = ((i64)blockIdx.x * 256 + threadIdx.x) * 2;
  if (x >= ncols) return;
  const double2* Wp = reinterpret_cast<const double2*>(Wg + x);
  const i64 ld2 = ldw / 2;
  double a0 = 0.0, a1 = 0.0;
  i64 j = j0;
  for (; j + 8 <= j1; j += 8) {
    double2 v[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) v[t] = ld_stream(Wp + (j + t) * ld2);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const double c = csm[j + t - j0];
      a0 += c * v[t].x;
      a1 += c * v[t].y;
    }
  }
  for (; j < j1; ++j) {
    const double2 v = ld_stream(Wp + j * ld2);
    const double c = csm[j - j0];
    a0 += c * v.x;
    a1 += c * v.y;
  }
  *reinterpret_cast<double2*>(part + ((i64)gp * n_split + blockIdx.y) * ncols + x) = make_double2(a0, a1);
}

// grid (min(ceil(ncols / 256), kFusedMaxBlocks)), block 256.
template <typename T>
__global__ void __launch_bounds__(256) fused_finish_kernel(FusedFin a) {
  __shared__ double sA[TAL_MAX_Q], s1[TAL_MAX_Q];
  __shared__ Rec sh[32];
  __shared__ bool is_last;
  const i64 idx = a.idx_dev ? *a.idx_dev : a.idx_host;
  const bool valid = idx >= 0 && idx < a.N;
  const T* kbuf = (const T*)a.kbuf;
  const T* wbuf = (const T*)a.wbuf;
  if ((int)threadIdx.x < a.q && valid) {
    const int gp = threadIdx.x;
    const double kx = (double)kbuf[(i64)gp * a.ld + idx];
    const double r = a.rho[(i64)gp * a.N + idx];
    double acc = 0.0;
    for (int s = 0; s < a.n_split; ++s) acc += a.colsq[(i64)gp * a.n_split + s];  // fixed order
    sA[gp] = kx + acc + r * r;
    s1[gp] = (double)T(kx + r * r);  // as fused_observe
  }
  __syncthreads();
  Rec rec = rec_empty();
  for (i64 x = (i64)blockIdx.x * 256 + threadIdx.x; x < a.ncols; x += (i64)gridDim.x * 256) {
    const bool cand = x < a.n_cand;
    const i64 gx = a.cand0 + x;
    double f = 0.0;
    for (int gp = 0; gp < a.q; ++gp) {
      double g = 0.0, h = 0.0;
      double* csp = a.cs + (i64)gp * a.ncols + x;
      double c = cand ? *csp : 0.0;
      if (cand && valid) {
        const double k = (double)kbuf[(i64)gp * a.ld + gx], w = (double)wbuf[(i64)gp * a.ld + gx];
        double kA = k;
        const double* pp = a.part + (i64)gp * a.n_split * a.ncols + x;
        for (int s = 0; s < a.n_split; ++s) kA += pp[(i64)s * a.ncols];  // fixed order: deterministic
        g = kA / sqrt(sA[gp]);
        h = k / sqrt(s1[gp]);
        c = c + (g * g - k * w);
        *csp = c;
      }
      double* Wg = a.Wb + (i64)gp * a.wb_stride_q;
      Wg[a.rows * a.ldw + x] = g;  // pad columns stay zero
      Wg[(a.rows + 1) * a.ldw + x] = h;
      if (cand && a.select) {  // 0.5 max(log(predictive / posterior predictive), 0)  (itl.py:53-58)
        const double v = a.var[(i64)gp * a.N + gx];
        const double r = a.rho[(i64)gp * a.N + gx];
        const double nz = r * r;
        const double pv = __dsub_rn(v, c);
        const double fi = 0.5 * max_nan(log((v + nz) / (pv + nz)), 0.0);
        f = gp ? f + fi : fi;
      }
    }
    if (cand && a.select) {
      if (a.sample && !a.sample[gx]) f = -INFINITY;  // F.at[~get_mask(domain, sample_region)].set(-inf)
      a.F[x] = f;
      bool is_safe;
      if (a.safe) {
        is_safe = a.safe[gx] != 0;
      } else {  // pessimistic safe set: all_{i >= fc} l_i[x] >= 0 (marginal.py:176-179)
        is_safe = true;
        for (int i = a.fc; i < a.q; ++i) is_safe = is_safe && (a.l[(i64)i * a.N + gx] >= 0.0);
      }
      rec = rec_combine(rec, rec_of(f, is_safe, gx));
    }
  }
  rec = block_reduce_rec(rec, sh);
  FusedScratch* sc = a.scratch;
  if (threadIdx.x == 0) {
    sc->part[blockIdx.x] = rec;
    __threadfence();
    const unsigned int t = atomicAdd(&sc->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  // ---- last block: every block of this pass has finished ----
  __threadfence();
  if (a.select) {
    Rec acc = rec_empty();
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) {
      const volatile Rec* p = &sc->part[i];
      Rec b;
      b.val = p->val; b.idx = p->idx; b.cnt = p->cnt; b.flags = p->flags;
      acc = rec_combine(acc, b);
    }
    __syncthreads();
    acc = block_reduce_rec(acc, sh);  // valid in warp 0
    if (threadIdx.x == 0) write_result(acc, a.rec);
  }
  if (threadIdx.x >= 32) return;
  if (threadIdx.x == 0) sc->ticket = 0;  // ready for the next launch on this stream
  __syncwarp();
  if (a.G <= 1) {
    if (a.select) {  // the record is final: stash mean[idx] of the NEXT observation (fused_observe reads it)
      const i64 nidx = a.rec->idx;
      if ((int)threadIdx.x < a.q && nidx >= 0 && nidx < a.N) a.scal[threadIdx.x] = a.mean[(i64)threadIdx.x * a.N + nidx];
    }
    return;
  }
  // sharded: this row exchange is consumed (every block has read kbuf / wbuf / pcol): tell the peers that
  // the row buffer is free for the next one, and publish the record of the next arg-max exchange
  PeerHeader* me = hdr(a.boxes[a.rank]);
  const unsigned long long s_row = me->row_step + 1;
  const unsigned long long s_rec = me->rec_step + 1;
  __syncwarp();
  const int g = threadIdx.x;
  if (g == 0) me->row_step = s_row;
  if (g < a.G) {
    PeerHeader* pg = hdr(a.boxes[g]);
    st_release_sys(&pg->ready[a.rank], s_row + 1);
    if (a.select) {
      PeerSlot* dst = &pg->slot[s_rec & 1][a.rank];
      const int4* src = reinterpret_cast<const int4*>(a.rec);
      int4* d4 = reinterpret_cast<int4*>(&dst->rec);
      d4[0] = src[0];
      d4[1] = src[1];
      __threadfence_system();
      st_release_sys(&dst->seq, s_rec);
    }
  }
}


__global__ void fused_stash_mean_kernel(const double* __restrict__ mean, i64 N, int q,
                                        const i64* __restrict__ idx_dev, i64 idx_host, double* __restrict__ scal) {
  const i64 idx = idx_dev ? *idx_dev : idx_host;
  if ((int)threadIdx.x < q && idx >= 0 && idx < N) scal[threadIdx.x] = mean[(i64)threadIdx.x * N + idx];
}

// ---- host-side launchers ------------------------------------------------------------------
int fused_observe(const FusedObs& a, int q, int dtype, cudaStream_t st) {
  const bool fma_mode = arith_fma(dtype);
  dtype = base_dtype(dtype);
  dim3 grid((unsigned)((a.ld + 255) / 256), q);
#define TAL_FO(T, F) fused_observe_kernel<T, F><<<grid, 256, 0, st>>>(a)
  if (dtype == TAL_F64) { if (fma_mode) TAL_FO(double, true); else TAL_FO(double, false); }
  else { if (fma_mode) TAL_FO(float, true); else TAL_FO(float, false); }
#undef TAL_FO
  TAL_LAUNCH_CHECK("fused_observe");
  return TAL_OK;
}

int fused_observe_y(const void* wbuf, i64 ld, i64 N, int q, const i64* idx_dev, i64 idx_host, const double* y,
                    i64 y_stride, int y_gather, const BetaQ& beta, const double* scal, const double* var,
                    double* mean, double* l, double* u, int dtype, cudaStream_t st) {
  dim3 grid((unsigned)((N + 255) / 256), q);
  if (base_dtype(dtype) == TAL_F64)
    fused_observe_y_kernel<double><<<grid, 256, 0, st>>>((const double*)wbuf, ld, N, idx_dev, idx_host, y, y_stride,
                                                         y_gather, beta, scal, var, mean, l, u);
  else
    fused_observe_y_kernel<float><<<grid, 256, 0, st>>>((const float*)wbuf, ld, N, idx_dev, idx_host, y, y_stride,
                                                        y_gather, beta, scal, var, mean, l, u);
  TAL_LAUNCH_CHECK("fused_observe_y");
  return TAL_OK;
}

int fused_partial(const double* Wb, i64 wb_stride_q, i64 ldw, i64 rows, i64 m_pad, const i64* idx_dev, i64 idx_host,
                  i64 cand0, i64 n_cand, const double* pcol, i64 pcol_stride, i64 ncols, double* part,
                  double* colsq, int n_split, int q, cudaStream_t st, void* wait_box, int G) {
  const i64 per = ((rows + n_split - 1) / n_split + 7) / 8 * 8;
  const size_t smem = (size_t)per * sizeof(double);
  if (smem > 200 * 1024) {
    set_error("fused_partial: %lld rows per split do not fit shared memory", (long long)per);
    return TAL_ERR_ARG;
  }
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    TAL_CUDA(cudaFuncSetAttribute(fused_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    smem_set = 200 * 1024;
  }
  dim3 grid((unsigned)((ncols / 2 + 255) / 256), (unsigned)n_split, q);
  fused_partial_kernel<<<grid, 256, smem, st>>>(Wb, wb_stride_q, ldw, rows, m_pad, idx_dev, idx_host, cand0, n_cand,
                                                pcol, pcol_stride, ncols, part, colsq, n_split, per, wait_box, G);
  TAL_LAUNCH_CHECK("fused_partial");
  return TAL_OK;
}

int fused_stash_mean(const double* mean, i64 N, int q, const i64* idx_dev, i64 idx_host, double* scal, cudaStream_t st) {
  fused_stash_mean_kernel<<<1, 32, 0, st>>>(mean, N, q, idx_dev, idx_host, scal);
  TAL_LAUNCH_CHECK("fused_stash_mean");
  return TAL_OK;
}

int fused_finish(const FusedFin& a, int dtype, cudaStream_t st) {
  i64 blocks = (a.ncols + 255) / 256;
  if (blocks > kFusedMaxBlocks) blocks = kFusedMaxBlocks;
  if (base_dtype(dtype) == TAL_F64) fused_finish_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(a);
  else fused_finish_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(a);
  TAL_LAUNCH_CHECK("fused_finish");
  return TAL_OK;
}

}  // namespace tal

extern "C" {

long long tal_fused_scratch_bytes(int q, long long ncols, int n_split) {
  return (long long)sizeof(tal::FusedScratch) + (long long)sizeof(double) * ((long long)q * n_split * (ncols + 1) + 16);
}

}  // extern "C"
