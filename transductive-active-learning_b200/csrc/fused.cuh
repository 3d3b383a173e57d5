// Argument blocks and host-side launchers of the fused step kernels (fused.cu), used by stepctx.cu.
#pragma once
#include "argmax.cuh"

namespace tal {

constexpr int kFusedMaxBlocks = 1024;

struct FusedScratch {  // device scratch of fused_finish (tal_fused_scratch_bytes)
  unsigned int ticket;
  unsigned int pad[7];
  Rec part[kFusedMaxBlocks];
};

struct FusedObs {
  const void* cov; i64 stride_q, ld, N, sym_block;      // G == 1: every row of the matrix is local
  const i64* idx_dev; i64 idx_host;
  void *kbuf, *wbuf;                                      // [q][ld]
  const void *Kp, *Wp; i64 pend_stride; int pending;      // pending (k, w) slots [0, pending)
  void *Kslot, *Wslot;                                    // slot `pending` (null: no deferral)
  const double* rho; double *mean, *var, *l, *u, *scal;
  const double* y; i64 y_stride; int y_gather;            // y == null: mean / l / u follow later
  BetaQ beta;
  void* mailbox; int G; int wait;                         // sharded: kbuf holds the exchanged row
};

struct FusedFin {
  double* Wb; i64 wb_stride_q, ldw, rows, ncols, cand0, n_cand, N, ld;
  const double *part, *colsq; int n_split, q;
  const void *kbuf, *wbuf; const double *rho, *var, *l, *mean;
  double* cs;
  const i64* idx_dev; i64 idx_host;
  double* F; const unsigned char *safe, *sample; int fc; int select;
  FusedScratch* scratch; tal_argmax_result* rec; double* scal;
  void* const* boxes; int G, rank;
};

int fused_observe(const FusedObs& a, int q, int dtype, cudaStream_t st);
int fused_observe_y(const void* wbuf, i64 ld, i64 N, int q, const i64* idx_dev, i64 idx_host, const double* y,
                    i64 y_stride, int y_gather, const BetaQ& beta, const double* scal, const double* var,
                    double* mean, double* l, double* u, int dtype, cudaStream_t st);
// wait_box != null: the kernel itself waits for the G pieces of the current row exchange (the peer-supplied column
// pcol travels with them) -- needed when it does not run behind fused_observe in stream order
int fused_partial(const double* Wb, i64 wb_stride_q, i64 ldw, i64 rows, i64 m_pad, const i64* idx_dev, i64 idx_host,
                  i64 cand0, i64 n_cand, const double* pcol, i64 pcol_stride, i64 ncols, double* part,
                  double* colsq, int n_split, int q, cudaStream_t st, void* wait_box = nullptr, int G = 1);
int fused_finish(const FusedFin& a, int dtype, cudaStream_t st);
int fused_stash_mean(const double* mean, i64 N, int q, const i64* idx_dev, i64 idx_host, double* scal, cudaStream_t st);

}  // namespace tal
