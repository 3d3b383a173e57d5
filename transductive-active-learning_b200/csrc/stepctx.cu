// One BO step as ONE host call, THREE kernels (sharded: four).
//
// Reference path: DiscreteGreedyAlgorithm._step (lib/algorithms/__init__.py:77-94):
// F -> masked arg-max -> observe -> MarginalModel.step.  The launch sequence of a step is issued
// here from a context structure that the host fills once (include/tal_b200.h: tal_step_ctx); the
// O(N) work of a step is fused into the kernels of fused.cu:
//
//   tal_ctx_select   makes `combined` hold the selection for the current state: nothing to do when
//                    the previous tal_ctx_observe already produced it (fused_finish), else the
//                    directed-ITL score from the resident conditioning state -> masked arg-max
//                    (-> peer all-gather + combine when sharded)
//   tal_ctx_observe  [sharded: push the pieces of row idx to every peer, combining the arg-max
//                    records of this step in the same kernel] -> fused_observe -> fused_partial ->
//                    fused_finish (which selects the NEXT candidate when `auto_select` is set) ->
//                    every update_block observations one pass over the matrix
//   tal_ctx_step     both, the observation taken from a device table
#include "fused.cuh"
#include "peer.cuh"
#include <cstdlib>
#include <algorithm>

using namespace tal;

static void ev_record(tal_step_ctx* c, bool start, void* stream, long long updates = 0) {
  if (!c->timing || c->n_ev >= TAL_CTX_MAX_EVENTS) return;
  void** slot = start ? &c->ev_start[c->n_ev] : &c->ev_stop[c->n_ev];
  if (!*slot) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    *slot = e;
  }
  cudaEventRecord((cudaEvent_t)*slot, as_stream(stream));
  if (!start) {
    c->ev_updates[c->n_ev] = updates;  // updates applied by the pass these events bracket
    c->n_ev += 1;
  }
}

// Per-phase breakdown of a fused step (diagnostic, TAL_BREAKDOWN=1): CUDA events between the launches of
// tal_ctx_observe for up to 256 steps, read by tal_ctx_breakdown (bench.py prints them as `breakdown_us`).
static int g_bd = -1;
static constexpr int kBdSteps = 256, kBdMarks = 6;
static cudaEvent_t g_bd_ev[kBdSteps][kBdMarks];
static bool g_bd_has[kBdSteps][kBdMarks];
static int g_bd_n = 0;
static bool bd_on() {
  if (g_bd < 0) {
    const char* e = getenv("TAL_BREAKDOWN");
    g_bd = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_bd == 1 && g_bd_n < kBdSteps;
}
static void bd_mark(int k, cudaStream_t st) {
  if (!bd_on()) return;
  if (!g_bd_ev[g_bd_n][k]) cudaEventCreate(&g_bd_ev[g_bd_n][k]);
  cudaEventRecord(g_bd_ev[g_bd_n][k], st);
  g_bd_has[g_bd_n][k] = true;
}

static inline int upd_dtype(const tal_step_ctx* c) { return (int)c->dtype | (c->arith ? TAL_ARITH_FMA : 0); }

static int ctx_flush(tal_step_ctx* c, void* stream) {
  if (c->pending == 0) return TAL_OK;
  ev_record(c, true, stream);
  int rc = tal_rankb_update(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, (int)c->q, c->Kp, c->Wp,
                            c->pend_stride, (int)c->pending, (int)c->lower, upd_dtype(c), stream);
  ev_record(c, false, stream, c->pending);
  if (rc != TAL_OK) return rc;
  // (slots >= pending are never read: nothing to clear)
  c->pending = 0;
  c->passes += 1;
  return TAL_OK;
}

static bool masks_match(const tal_step_ctx* c) {
  return c->sel_safe == c->safe && c->sel_sample == c->sample && c->sel_fc == c->first_constraint;
}

// consume an arg-max exchange that fused_finish published: wait for the G records, reduce them
// into `combined` (and advance the exchange counter)
static int ctx_combine(tal_step_ctx* c, void* stream) {
  int rc = tal_peer_combine_records(c->mailbox, (int)c->G, c->combined, stream);
  if (rc == TAL_OK) c->sel_state = 2;
  return rc;
}

extern "C" {

int tal_ctx_flush(tal_step_ctx* c, void* stream) {
  TAL_ARG(c);
  return ctx_flush(c, stream);
}

int tal_ctx_select(tal_step_ctx* c, void* stream) {
  TAL_ARG(c && c->F && c->cs && c->rec && c->combined);
  if (c->sel_state != 0 && !masks_match(c)) {  // selected under other masks: discard
    if (c->sel_state == 1) {
      int rc = ctx_combine(c, stream);
      if (rc != TAL_OK) return rc;
    }
    c->sel_state = 0;
  }
  if (c->sel_state == 2) return TAL_OK;
  if (c->sel_state == 1) return ctx_combine(c, stream);
  for (int i = 0; i < (int)c->q; ++i) {
    int rc = tal_score_itl_directed(c->var + i * c->N + c->cand0, c->cs + i * c->ncols, c->rho + i * c->N + c->cand0,
                                    c->n_cand, i > 0, c->F, stream);
    if (rc != TAL_OK) return rc;
  }
  const unsigned char* safe = c->safe ? c->safe + c->cand0 : nullptr;
  const unsigned char* sample = c->sample ? c->sample + c->cand0 : nullptr;
  tal_argmax_result* out = c->G > 1 ? c->rec : c->combined;
  int rc = tal_masked_argmax(c->F, safe, safe ? nullptr : c->l + c->cand0, (int)c->q, (int)c->first_constraint,
                             c->N, sample, c->n_cand, c->cand0, c->argmax_scratch, out, stream);
  if (rc != TAL_OK) return rc;
  if (c->G > 1) {
    TAL_ARG(c->mailboxes && c->mailbox);
    rc = tal_peer_publish_record(c->rec, c->mailboxes, (int)c->G, (int)c->rank, stream);
    if (rc != TAL_OK) return rc;
    rc = tal_peer_combine_records(c->mailbox, (int)c->G, c->combined, stream);
    if (rc != TAL_OK) return rc;
  }
  c->sel_state = 2;
  c->scal_ok = 0;
  c->sel_safe = c->safe; c->sel_sample = c->sample; c->sel_fc = c->first_constraint;
  return TAL_OK;
}

int tal_ctx_observe(tal_step_ctx* c, const long long* idx_dev, long long idx_host, const double* y_dev,
                    long long y_stride, int y_gather, void* stream) {
  TAL_ARG(c && c->cov && c->kbuf && c->wbuf);
  const int q = (int)c->q, dtype = (int)c->dtype;
  cudaStream_t st = as_stream(stream);
  bool cond = c->Wb != nullptr && c->cond_valid != 0;
  bool rebase = false;
  if (cond && c->rows + 2 > c->cap_rows) {  // the factor is full: this observation is applied without
    cond = false;                           // it and the caller re-conditions from scratch (TAL_CTX_REBASE)
    rebase = true;
    c->cond_valid = 0;
  }
  if (cond) TAL_ARG(c->fz_scratch && c->cs && (c->auto_select == 0 || (c->F && c->rec && c->combined)));
  if (cond && c->overlap != 0 && !c->side_stream) {  // (created before anything of this step is enqueued)
    cudaStream_t s2; cudaEvent_t e1, e2;
    TAL_CUDA(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    TAL_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
    TAL_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    c->side_stream = s2; c->ev_fork = e1; c->ev_join = e2;
  }
  const bool own = idx_dev != nullptr && c->combined != nullptr && idx_dev == &c->combined->idx;
  int rc;
  if (c->sel_state == 1 && !(own && masks_match(c))) {  // a published selection nobody will use: consume it
    rc = ctx_combine(c, stream);
    if (rc != TAL_OK) return rc;
    c->sel_state = 0;
  }
  const bool combine_in_push = c->G > 1 && c->sel_state == 1;
  bool scal_ok = own && c->scal_ok != 0;
  bd_mark(0, st);
  if (c->G > 1) {
    rc = peer_push_row_impl(c->mailboxes, (int)c->G, (int)c->rank, c->bounds, c->cov, c->cov_stride_q, c->ld, c->N, q,
                            (int)c->lower, idx_dev, idx_host, c->kbuf_off, cond ? c->Wb : nullptr,
                            c->cap_rows * c->ldw, c->ldw, cond ? c->rows : 0, c->pcol_off, c->pcol_stride, c->cand0,
                            c->n_cand, dtype, combine_in_push ? c->combined : nullptr, c->mean, c->scal, stream);
    if (rc != TAL_OK) return rc;
    if (combine_in_push) scal_ok = true;
    if (!cond) {  // nobody advances the row exchange later: the stand-alone wait does
      rc = tal_peer_wait_row(c->mailbox, (int)c->G, idx_dev, idx_host, c->mean, c->N, q, c->scal, stream);
      if (rc != TAL_OK) return rc;
      scal_ok = true;
    }
  }
  if (y_dev && !scal_ok) {
    rc = fused_stash_mean(c->mean, c->N, q, idx_dev, idx_host, c->scal, st);
    if (rc != TAL_OK) return rc;
  }
  bd_mark(1, st);
  const long long s = dtype == TAL_F64 ? 8 : 4;
  FusedObs o;
  o.cov = c->cov; o.stride_q = c->cov_stride_q; o.ld = c->ld; o.N = c->N;
  o.sym_block = c->lower ? tal_sym_block(dtype) : 0;
  o.idx_dev = idx_dev; o.idx_host = idx_host;
  o.kbuf = c->kbuf; o.wbuf = c->wbuf;
  o.Kp = c->Kp; o.Wp = c->Wp; o.pend_stride = c->pend_stride; o.pending = (int)c->pending;
  o.Kslot = o.Wslot = nullptr;
  if (c->update_block > 1) {
    o.Kslot = (char*)c->Kp + (size_t)c->pending * (size_t)c->pend_stride * (size_t)s;
    o.Wslot = (char*)c->Wp + (size_t)c->pending * (size_t)c->pend_stride * (size_t)s;
  }
  o.rho = c->rho; o.mean = c->mean; o.var = c->var; o.l = c->l; o.u = c->u; o.scal = c->scal;
  o.y = y_dev; o.y_stride = y_stride; o.y_gather = y_gather;
  for (int i = 0; i < TAL_MAX_Q; ++i) o.beta.v[i] = c->beta[i];
  o.mailbox = c->G > 1 ? c->mailbox : nullptr; o.G = (int)c->G; o.wait = cond ? 1 : 0;
  // fused_observe (needs the row) and fused_partial (needs the factor's column) are independent: with `overlap` the
  // scoring pass is forked onto a side stream before fused_observe and joined before fused_finish
  const bool forked = cond && c->overlap != 0;
  double* part = nullptr;
  double* colsq = nullptr;
  if (cond) {
    double* base = (double*)((char*)c->fz_scratch + sizeof(FusedScratch));
    colsq = base + 8;
    part = colsq + (((long long)q * c->n_split + 1) & ~1LL);  // (16-byte aligned: double2 stores)
  }
  if (forked) {
    cudaStream_t s2 = (cudaStream_t)c->side_stream;
    TAL_CUDA(cudaEventRecord((cudaEvent_t)c->ev_fork, st));
    TAL_CUDA(cudaStreamWaitEvent(s2, (cudaEvent_t)c->ev_fork, 0));
    rc = fused_partial(c->Wb, c->cap_rows * c->ldw, c->ldw, c->rows, c->m_pad, idx_dev, idx_host, c->cand0, c->n_cand,
                       c->G > 1 ? c->pcol_peer : nullptr, c->pcol_stride, c->ncols, part, colsq, (int)c->n_split, q,
                       s2, c->G > 1 ? c->mailbox : nullptr, (int)c->G);
    if (rc != TAL_OK) return rc;
    TAL_CUDA(cudaEventRecord((cudaEvent_t)c->ev_join, s2));
  }
  rc = fused_observe(o, q, upd_dtype(c), st);
  if (rc != TAL_OK) return rc;
  bd_mark(2, st);
  c->scal_ok = 0;
  c->sel_state = 0;
  if (cond) {
    if (forked) {
      TAL_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)c->ev_join, 0));
    } else {
      rc = fused_partial(c->Wb, c->cap_rows * c->ldw, c->ldw, c->rows, c->m_pad, idx_dev, idx_host, c->cand0, c->n_cand,
                         c->G > 1 ? c->pcol_peer : nullptr, c->pcol_stride, c->ncols, part, colsq, (int)c->n_split, q,
                         st);
      if (rc != TAL_OK) return rc;
    }
    bd_mark(3, st);
    const bool select = c->auto_select != 0 && y_dev != nullptr;  // (l, u are final only once y has been applied)
    FusedFin f;
    f.Wb = c->Wb; f.wb_stride_q = c->cap_rows * c->ldw; f.ldw = c->ldw; f.rows = c->rows; f.ncols = c->ncols;
    f.cand0 = c->cand0; f.n_cand = c->n_cand; f.N = c->N; f.ld = c->ld;
    f.part = part; f.colsq = colsq; f.n_split = (int)c->n_split; f.q = q;
    f.kbuf = c->kbuf; f.wbuf = c->wbuf; f.rho = c->rho; f.var = c->var; f.l = c->l; f.mean = c->mean;
    f.cs = c->cs; f.idx_dev = idx_dev; f.idx_host = idx_host;
    f.F = c->F; f.safe = c->safe; f.sample = c->sample; f.fc = (int)c->first_constraint; f.select = select ? 1 : 0;
    f.scratch = (FusedScratch*)c->fz_scratch; f.rec = c->G > 1 ? c->rec : c->combined; f.scal = c->scal;
    f.boxes = c->mailboxes; f.G = (int)c->G; f.rank = (int)c->rank;
    rc = fused_finish(f, dtype, st);
    if (rc != TAL_OK) return rc;
    bd_mark(4, st);
    c->rows += 2;
    if (select) {
      c->sel_state = c->G > 1 ? 1 : 2;
      c->scal_ok = c->G > 1 ? 0 : 1;
      c->sel_safe = c->safe; c->sel_sample = c->sample; c->sel_fc = c->first_constraint;
    }
  }
  // the covariance update of this observation is queued (fused_observe wrote the slot); every
  // update_block of them are applied in one pass
  if (c->update_block > 1) {
    c->pending += 1;
    if (c->pending >= c->update_block) {
      rc = ctx_flush(c, stream);
      if (rc != TAL_OK) return rc;
    }
    bd_mark(5, st);
    if (bd_on()) g_bd_n += 1;
    return rebase ? TAL_CTX_REBASE : TAL_OK;
  }
  c->passes += 1;
  ev_record(c, true, stream);
  if (c->lower)
    rc = tal_rank1_update_lower(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, q, c->kbuf, c->wbuf,
                                upd_dtype(c), stream);
  else
    rc = tal_rank1_update(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, q, c->kbuf, c->wbuf, upd_dtype(c),
                          0, stream);
  ev_record(c, false, stream, 1);
  if (rc != TAL_OK) return rc;
  return rebase ? TAL_CTX_REBASE : TAL_OK;
}

int tal_ctx_observe_y(tal_step_ctx* c, const long long* idx_dev, long long idx_host, const double* y_dev,
                      long long y_stride, int y_gather, void* stream) {
  TAL_ARG(c && y_dev);
  BetaQ beta;
  for (int i = 0; i < TAL_MAX_Q; ++i) beta.v[i] = c->beta[i];
  return fused_observe_y(c->wbuf, c->ld, c->N, (int)c->q, idx_dev, idx_host, y_dev, y_stride, y_gather, beta, c->scal,
                         c->var, c->mean, c->l, c->u, (int)c->dtype, as_stream(stream));
}

int tal_ctx_step(tal_step_ctx* c, void* stream) {
  TAL_ARG(c && c->y_table && c->combined);
  // sharded, selection already published by the previous step: the push kernel combines the records
  if (!(c->sel_state == 1 && masks_match(c))) {
    int rc = tal_ctx_select(c, stream);
    if (rc != TAL_OK) return rc;
  }
  return tal_ctx_observe(c, &c->combined->idx, 0, c->y_table, c->y_stride, 1, stream);
}

long long tal_step_ctx_bytes(void) { return (long long)sizeof(tal_step_ctx); }

/* Diagnostic (TAL_BREAKDOWN=1): mean microseconds of the phases of the last recorded fused steps --
 * out_us[0] peer push (+ record combine), [1] fused_observe (+ row wait), [2] fused_partial, [3] fused_finish
 * (+ arg-max + publish), [4] covariance pass (amortised), [5] gap between consecutive steps; *n = steps seen.
 * Synchronises the device and resets the recording. */
int tal_ctx_breakdown(double* out_us_host, long long* n_host) {
  TAL_ARG(out_us_host && n_host);
  TAL_CUDA(cudaDeviceSynchronize());
  static double v[6][kBdSteps];
  long long cnt = 0, gaps = 0;
  double pass_sum = 0.0;
  for (int s = 0; s < g_bd_n; ++s) {
    bool all = true;
    for (int k = 0; k < kBdMarks; ++k) all = all && g_bd_has[s][k];
    if (!all) continue;
    for (int k = 0; k < 5; ++k) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_bd_ev[s][k], g_bd_ev[s][k + 1]);
      v[k][cnt] = ms * 1e3;
    }
    pass_sum += v[4][cnt];
    cnt += 1;
    if (s + 1 < g_bd_n && g_bd_has[s + 1][0]) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_bd_ev[s][5], g_bd_ev[s + 1][0]);
      v[5][gaps++] = ms * 1e3;
    }
  }
  // medians (a few steps sit next to host-side set-up work); the pass is amortised: its mean
  for (int k = 0; k < 6; ++k) {
    const long long n = k == 5 ? gaps : cnt;
    if (n == 0) { out_us_host[k] = 0.0; continue; }
    std::sort(v[k], v[k] + n);
    out_us_host[k] = v[k][n / 2];
  }
  out_us_host[4] = cnt ? pass_sum / cnt : 0.0;
  *n_host = cnt;
  for (int s = 0; s < g_bd_n; ++s)
    for (int k = 0; k < kBdMarks; ++k) g_bd_has[s][k] = false;
  g_bd_n = 0;
  return TAL_OK;
}

int tal_ctx_timing_begin(tal_step_ctx* c) {
  TAL_ARG(c);
  c->n_ev = 0;
  c->timing = 1;
  return TAL_OK;
}

int tal_ctx_timing_read(tal_step_ctx* c, double* ms_out_host, long long* updates_out_host, long long* n_out_host,
                        void* stream) {
  TAL_ARG(c && ms_out_host && n_out_host);
  TAL_CUDA(cudaStreamSynchronize(as_stream(stream)));
  for (long long i = 0; i < c->n_ev; ++i) {
    float ms = 0.f;
    TAL_CUDA(cudaEventElapsedTime(&ms, (cudaEvent_t)c->ev_start[i], (cudaEvent_t)c->ev_stop[i]));
    ms_out_host[i] = ms;
    if (updates_out_host) updates_out_host[i] = c->ev_updates[i];
  }
  *n_out_host = c->n_ev;
  c->timing = 0;
  return TAL_OK;
}

}  // extern "C"
