// Mailbox layout and release/acquire helpers of the NVLink peer-memory exchange, shared by
// peer.cu and fused.cu (see peer.cu for the protocol).
#pragma once
#include "common.cuh"

namespace tal {

constexpr int kPeerMax = TAL_PEER_MAX;
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct alignas(64) PeerSlot {
  tal_argmax_result rec;       // 32 B
  unsigned long long seq;      // exchange number this record belongs to
  unsigned long long pad[3];
};

struct alignas(256) PeerHeader {
  unsigned long long rec_step;  // arg-max exchanges completed by THIS rank
  unsigned long long row_step;  // row exchanges completed by THIS rank
  unsigned long long pad1;
  unsigned int ticket;          // last-block election of the push kernel
  int error;                    // != 0: a spin timed out (TAL_PEER_ERR_*)
  unsigned long long pad0[4];
  unsigned long long ready[kPeerMax];    // ready[g] = s: rank g may receive row exchange s
  unsigned long long row_seq[kPeerMax];  // row_seq[g] = s: rank g's piece of exchange s has landed here
  PeerSlot slot[2][kPeerMax];
  // the diagonal element cov_i[x*][x*] of the row being exchanged, as stored (no pending update applied), one per
  // GP: written by the owner of row x* together with its piece.  fused_observe corrects the row IN PLACE in kbuf,
  // so a block that starts late must not take the diagonal from kbuf[x*] (its owner block may already have
  // corrected it: the correction would be applied twice, and that block's weights w = k / s would be off)
  double kdiag[TAL_MAX_Q];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spin until *p >= want (the counters only grow; a peer without a piece may already
// have announced the next exchange); false on timeout
__device__ __forceinline__ bool spin_until(const unsigned long long* p, unsigned long long want) {
  if (ld_acquire_sys(p) >= want) return true;
  const unsigned long long t0 = now_ns();
  for (;;) {
    for (int i = 0; i < 64; ++i)
      if (ld_acquire_sys(p) >= want) return true;
    if (now_ns() - t0 > kSpinTimeoutNs) return false;
    __nanosleep(100);
  }
}

__device__ __forceinline__ PeerHeader* hdr(void* mailbox) { return reinterpret_cast<PeerHeader*>(mailbox); }


// One warp: lane g < G waits for rank g's record of exchange s in mailbox h, then the records are
// reduced (max value, lowest index among equals, tie counts add, flags OR); the result is valid
// in every lane.  A timed-out spin or an earlier exchange error of this mailbox sets flag bit 3
// (status TAL_ARGMAX_PEER_ERROR): the host raises instead of continuing on stale data.
__device__ __forceinline__ tal_argmax_result combine_records_warp(PeerHeader* h, unsigned long long s, int G) {
  const int g = threadIdx.x & 31;
  double val = -INFINITY;
  i64 idx = 0x7fffffffffffffffLL, cnt = 0;
  int flags = 0;
  bool ok = true;
  if (g < G) {
    PeerSlot* sl = &h->slot[s & 1][g];
    ok = spin_until(&sl->seq, s);
    const volatile tal_argmax_result* r = &sl->rec;
    val = r->val; idx = r->idx; cnt = r->n_ties; flags = r->flags;
    if (cnt == 0) { val = -INFINITY; idx = 0x7fffffffffffffffLL; }
  }
  if (!__all_sync(0xffffffffu, ok)) {
    if (g == 0) h->error = TAL_PEER_ERR_RECORD_TIMEOUT;
    flags |= 8;
  }
  if (*(volatile int*)&h->error != 0) flags |= 8;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, val, o);
    const i64 i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    const i64 c2 = __shfl_xor_sync(0xffffffffu, cnt, o);
    const int f2 = __shfl_xor_sync(0xffffffffu, flags, o);
    flags |= f2;
    if (c2 == 0 || (cnt != 0 && val > v2)) {
    } else if (cnt == 0 || v2 > val) {
      val = v2; idx = i2; cnt = c2;
    } else {
      idx = idx < i2 ? idx : i2; cnt += c2;
    }
  }
  tal_argmax_result out;
  out.status = (flags & 8) ? TAL_ARGMAX_PEER_ERROR
               : !(flags & 1) ? TAL_ARGMAX_EMPTY_SAFE_SET
               : !(flags & 2) ? TAL_ARGMAX_ALL_NEG_INF
               : !(flags & 4) ? TAL_ARGMAX_ALL_NAN : TAL_ARGMAX_OK;
  out.idx = (cnt && out.status == TAL_ARGMAX_OK) ? idx : -1;  // (a failed arg-max is never applied)
  out.val = val;
  out.n_ties = cnt;
  out.flags = flags;
  return out;
}

int peer_push_row_impl(void* const* mailboxes, int G, int rank, const long long* bounds_host, const void* cov,
                       long long cov_stride_q, long long ld, long long N, int q, int lower,
                       const long long* idx_dev, long long idx_host, long long kbuf_off, const double* V,
                       long long v_stride_q, long long ldv, long long m_pad, long long pcol_off,
                       long long pcol_stride, long long cand0, long long n_cand, int dtype,
                       tal_argmax_result* combined, const double* mean, double* scal, void* stream);

}  // namespace tal
