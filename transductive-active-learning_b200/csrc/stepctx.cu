// One BO step as ONE host call.
//
// Reference path: DiscreteGreedyAlgorithm._step (lib/algorithms/__init__.py:77-94):
// F -> masked arg-max -> observe -> MarginalModel.step.  On the device a step is a
// fixed sequence of ~10 short kernels next to the two streaming passes; issued
// one by one from Python their launch cost (~0.5 ms) exceeds their run time, so
// the sequence is issued here from a context structure filled once by the host
// (include/tal_b200.h: tal_step_ctx).  Nothing new is computed: every launch goes
// through the library's own entry points.
//
//   tal_ctx_select   directed-ITL score from the resident conditioning state ->
//                    masked arg-max (-> peer all-gather + combine when sharded)
//   tal_ctx_observe  observed row (extract or peer exchange) -> pending-row
//                    correction -> step vectors -> conditioning update -> queue the
//                    covariance update (a pass over the matrix every update_block)
//   tal_ctx_step     both, the observation taken from a device table
#include "common.cuh"

using namespace tal;

static void ev_record(tal_step_ctx* c, bool start, void* stream) {
  if (!c->timing || c->n_ev >= TAL_CTX_MAX_EVENTS) return;
  void** slot = start ? &c->ev_start[c->n_ev] : &c->ev_stop[c->n_ev];
  if (!*slot) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    *slot = e;
  }
  cudaEventRecord((cudaEvent_t)*slot, as_stream(stream));
  if (!start) c->n_ev += 1;
}

static int ctx_flush(tal_step_ctx* c, void* stream) {
  if (c->pending == 0) return TAL_OK;
  ev_record(c, true, stream);
  int rc = tal_rankb_update(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, (int)c->q, c->Kp, c->Wp,
                            c->pend_stride, (int)c->pending, (int)c->lower, (int)c->dtype, stream);
  ev_record(c, false, stream);
  if (rc != TAL_OK) return rc;
  const size_t bytes = (size_t)c->pend_slots * (size_t)c->pend_stride * (c->dtype == TAL_F64 ? 8 : 4);
  TAL_CUDA(cudaMemsetAsync(c->Kp, 0, bytes, as_stream(stream)));
  TAL_CUDA(cudaMemsetAsync(c->Wp, 0, bytes, as_stream(stream)));
  c->pending = 0;
  c->passes += 1;
  return TAL_OK;
}

extern "C" {

int tal_ctx_flush(tal_step_ctx* c, void* stream) {
  TAL_ARG(c);
  return ctx_flush(c, stream);
}

int tal_ctx_select(tal_step_ctx* c, void* stream) {
  TAL_ARG(c && c->F && c->cs && c->rec && c->combined);
  const long long s = c->dtype == TAL_F64 ? 8 : 4;
  (void)s;
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
  return TAL_OK;
}

int tal_ctx_observe(tal_step_ctx* c, const long long* idx_dev, long long idx_host, const double* y_dev,
                    long long y_stride, int y_gather, void* stream) {
  TAL_ARG(c && c->cov && c->kbuf && c->wbuf && y_dev);
  const int q = (int)c->q, dtype = (int)c->dtype;
  const bool cond = c->Wb != nullptr && c->cond_valid != 0;
  if (cond && c->rows + 2 > c->cap_rows) return TAL_CTX_REBASE;  // the host re-conditions from scratch
  int rc;
  if (c->G > 1) {
    rc = tal_peer_push_row(c->mailboxes, (int)c->G, (int)c->rank, c->bounds, c->cov, c->cov_stride_q, c->ld, c->N, q,
                           (int)c->lower, idx_dev, idx_host, c->kbuf_off, cond ? c->Wb : nullptr,
                           c->cap_rows * c->ldw, c->ldw, cond ? c->rows : 0, c->pcol_off, c->cand0, c->n_cand, dtype,
                           stream);
    if (rc != TAL_OK) return rc;
    rc = tal_peer_wait_row(c->mailbox, (int)c->G, idx_dev, idx_host, c->mean, c->N, q, c->scal, stream);
  } else if (c->lower) {
    rc = tal_extract_row_lower(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, q, idx_dev, idx_host,
                               c->mean, c->kbuf, c->scal, dtype, stream);
  } else {
    rc = tal_extract_row(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, q, idx_dev, idx_host, c->mean,
                         c->kbuf, c->scal, dtype, stream);
  }
  if (rc != TAL_OK) return rc;
  if (c->pending > 0) {
    rc = tal_pending_correct_row(c->kbuf, c->ld, c->N, q, c->Kp, c->Wp, c->pend_stride, (int)c->pending, idx_dev,
                                 idx_host, (int)c->lower, dtype, stream);
    if (rc != TAL_OK) return rc;
  }
  rc = tal_step_vectors(c->kbuf, c->wbuf, c->N, q, idx_dev, idx_host, y_dev, y_stride, y_gather, c->rho, c->beta,
                        c->scal, c->mean, c->var, c->l, c->u, dtype, stream);
  if (rc != TAL_OK) return rc;
  const long long s = dtype == TAL_F64 ? 8 : 4;
  if (cond) {
    for (int i = 0; i < q; ++i) {
      const double* col;
      if (c->G > 1) {
        col = c->pcol_peer + (long long)i * c->pcol_stride;
      } else {
        rc = tal_cond_extract_col(c->Wb + (long long)i * c->cap_rows * c->ldw, c->ldw, c->rows, c->cand0, c->n_cand,
                                  idx_dev, idx_host, c->pcol, stream);
        if (rc != TAL_OK) return rc;
        col = c->pcol;
      }
      rc = tal_cond_append_update(c->Wb + (long long)i * c->cap_rows * c->ldw, c->ldw, c->cap_rows, c->m_pad, c->rows,
                                  c->ncols, c->n_cand, c->cand0, col, (const char*)c->kbuf + (long long)i * c->ld * s,
                                  (const char*)c->wbuf + (long long)i * c->ld * s, c->rho + (long long)i * c->N,
                                  idx_dev, idx_host, c->cs + (long long)i * c->ncols, c->app_scratch, (int)c->n_split,
                                  dtype, stream);
      if (rc != TAL_OK) return rc;
    }
    c->rows += 2;
  }
  // queue the covariance update of this observation; apply every update_block of them in one pass
  if (c->update_block > 1) {
    const size_t row_bytes = (size_t)q * (size_t)c->ld * (size_t)s;
    char* Kslot = (char*)c->Kp + (size_t)c->pending * (size_t)c->pend_stride * (size_t)s;
    char* Wslot = (char*)c->Wp + (size_t)c->pending * (size_t)c->pend_stride * (size_t)s;
    TAL_CUDA(cudaMemcpyAsync(Kslot, c->kbuf, row_bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    TAL_CUDA(cudaMemcpyAsync(Wslot, c->wbuf, row_bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    c->pending += 1;
    if (c->pending >= c->update_block) return ctx_flush(c, stream);
    return TAL_OK;
  }
  c->passes += 1;
  ev_record(c, true, stream);
  if (c->lower)
    rc = tal_rank1_update_lower(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, q, c->kbuf, c->wbuf, dtype,
                                stream);
  else
    rc = tal_rank1_update(c->cov, c->cov_stride_q, c->ld, c->row0, c->n_loc, c->N, q, c->kbuf, c->wbuf, dtype, 0,
                          stream);
  ev_record(c, false, stream);
  return rc;
}

int tal_ctx_step(tal_step_ctx* c, void* stream) {
  TAL_ARG(c && c->y_table);
  int rc = tal_ctx_select(c, stream);
  if (rc != TAL_OK) return rc;
  return tal_ctx_observe(c, &c->combined->idx, 0, c->y_table, c->y_stride, 1, stream);
}

long long tal_step_ctx_bytes(void) { return (long long)sizeof(tal_step_ctx); }

int tal_ctx_timing_begin(tal_step_ctx* c) {
  TAL_ARG(c);
  c->n_ev = 0;
  c->timing = 1;
  return TAL_OK;
}

int tal_ctx_timing_read(tal_step_ctx* c, double* ms_out_host, long long* n_out_host, void* stream) {
  TAL_ARG(c && ms_out_host && n_out_host);
  TAL_CUDA(cudaStreamSynchronize(as_stream(stream)));
  for (long long i = 0; i < c->n_ev; ++i) {
    float ms = 0.f;
    TAL_CUDA(cudaEventElapsedTime(&ms, (cudaEvent_t)c->ev_start[i], (cudaEvent_t)c->ev_stop[i]));
    ms_out_host[i] = ms;
  }
  *n_out_host = c->n_ev;
  c->timing = 0;
  return TAL_OK;
}

}  // extern "C"
