// Exchange steps of the row-sharded model over NVLink peer memory.
//
// One process per GPU; every rank owns a "mailbox" (one cudaMalloc, exported
// with cudaIpcGetMemHandle and mapped by the peers) and a device table of the G
// mailbox pointers.  The two per-step exchanges then need no NCCL call and no
// host-known root:
//
//   arg-max      every rank STORES its 32-byte record into slot[rank] of every
//                peer's mailbox (NVLink P2P stores), releases a sequence number;
//                every rank spins on its own slots and combines (same rule as
//                tal_argmax_combine) — an all-gather in one hop.
//   observed row every rank tells the owner of row x* "ready" (it has finished
//                the previous update, so its row buffer may be overwritten);
//                the owner waits for the G readies, STORES Sigma[x*, :] (and
//                V[:, x*] of the incremental conditioning) into every peer's row
//                buffer and releases a sequence number; the others spin on it.
//
// All counters live in device memory and advance inside the kernels, so a step
// is a fixed launch sequence (CUDA-graph capturable).  Spins are bounded
// (kSpinTimeoutNs): on expiry the mailbox error word is set and the kernel
// returns instead of hanging the GPU.
#include "peer.cuh"
#include <cstring>

namespace tal {

// ---- arg-max all-gather ----------------------------------------------------
// One warp: lane g publishes this rank's record to peer g.
__global__ void peer_publish_kernel(const tal_argmax_result* __restrict__ rec, void* const* __restrict__ boxes,
                                    int G, int rank) {
  const int g = threadIdx.x;
  if (g >= G) return;
  const unsigned long long s = hdr(boxes[rank])->rec_step + 1;  // local, stream-ordered
  PeerSlot* dst = &hdr(boxes[g])->slot[s & 1][rank];
  const int4* src = reinterpret_cast<const int4*>(rec);
  int4* d4 = reinterpret_cast<int4*>(&dst->rec);
  d4[0] = src[0];
  d4[1] = src[1];
  __threadfence_system();
  st_release_sys(&dst->seq, s);
}

// One warp: lane g waits for rank g's record, then the records are combined
// (max value, lowest index among equals, tie counts add, flags OR).
__global__ void peer_combine_kernel(void* mailbox, int G, tal_argmax_result* __restrict__ out) {
  PeerHeader* h = hdr(mailbox);
  const unsigned long long s = h->rec_step + 1;
  const tal_argmax_result r = combine_records_warp(h, s, G);
  if (threadIdx.x == 0) {
    *out = r;
    h->rec_step = s;
  }
}

// ---- observed-row broadcast --------------------------------------------------
// Row x* of every GP is assembled in every mailbox's kbuf from the pieces the
// shards hold.  Full storage: the owner of row x* holds all of it.  Lower-block
// storage (sym_block > 0): the owner holds columns [0, cend(x*)), and the shard
// that owns row c >= cend(x*) holds entry c as cov[c][x*].  Every rank
//   1. tells every peer "my row buffer is free for exchange s"   (ready[rank] = s),
//   2. stores its piece into every peer's kbuf once that peer is ready,
//   3. tells every peer "my piece of exchange s has landed"       (row_seq[rank] = s);
// tal_peer_wait_row then waits for the G row_seq entries.  The owner also stores
// the column V[:, x*] of the incremental conditioning (pcol).
// grid (ceil(max(ld / VN, m_pad) / 256), q, G): block z stores to ONE peer
// (staggered by rank so that the links are loaded evenly), 16 bytes per thread.
struct ShardBounds {
  i64 b[kPeerMax + 1];  // rank g holds global rows [b[g], b[g + 1])
};

// COMBINE: the arg-max records of this step were published (fused.cu, last block of the select
// pass) but not yet combined: every block waits for the G records and reduces them itself (the
// index of the row is their result), block (0,0,0) writes the combined record and scal[i] =
// mean[i][idx], the last block to finish advances rec_step.
template <typename T, bool COMBINE>
__global__ void __launch_bounds__(256)
peer_push_row_kernel(void* const* __restrict__ boxes, int G, int rank, ShardBounds bounds,
                     const T* __restrict__ cov, i64 stride_q, i64 ld, i64 N, i64 sym_block,
                     const i64* __restrict__ idx_dev, i64 idx_host, tal_argmax_result* __restrict__ combined,
                     const double* __restrict__ mean, double* __restrict__ scal, int q,
                     i64 kbuf_off,  // byte offset of kbuf [q][ld] inside a mailbox
                     const double* __restrict__ V, i64 v_stride_q, i64 ldv, i64 m_pad,
                     i64 pcol_off,  // byte offset of pcol [q][pcol_stride] inside a mailbox (V != null)
                     i64 pcol_stride,  // doubles between the GPs' columns in the mailbox (>= m_pad)
                     i64 cand0, i64 n_cand) {  // V holds columns (candidates) [cand0, cand0 + n_cand)
  typedef typename Vec16<T>::type V16;
  constexpr int VN = Vec16<T>::n;
  __shared__ bool s_last;
  __shared__ i64 s_idx;
  const int gp = blockIdx.y;
  PeerHeader* me = hdr(boxes[rank]);
  i64 idx;
  unsigned long long s_rec = 0;
  if constexpr (COMBINE) {
    s_rec = me->rec_step + 1;
    if (threadIdx.x < 32) {
      const tal_argmax_result r = combine_records_warp(me, s_rec, G);
      if (threadIdx.x == 0) {
        s_idx = r.idx;
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) *combined = r;
      }
    }
    __syncthreads();
    idx = s_idx;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (int)threadIdx.x < q && idx >= 0 && idx < N)
      scal[threadIdx.x] = mean[(i64)threadIdx.x * N + idx];
  } else {
    idx = idx_dev ? *idx_dev : idx_host;
  }
  if (idx < 0 || idx >= N) idx = -1;  // failed arg-max: zeros are exchanged, the protocol keeps moving
  const unsigned long long s = me->row_step + 1;
  if (blockIdx.x == 0 && gp == 0 && blockIdx.z == 0 && (int)threadIdx.x < G)
    st_release_sys(&hdr(boxes[threadIdx.x])->ready[rank], s);
  const i64 row0 = bounds.b[rank], n_loc = bounds.b[rank + 1] - row0;
  const i64 row_end = rank == G - 1 ? ld : row0 + n_loc;  // pad columns belong to the last shard (zeros)
  const bool own = idx >= row0 && idx < row0 + n_loc;
  const i64 cend = idx < 0 ? ld : (sym_block > 0 ? ((idx / sym_block + 1) * sym_block < ld
                                                        ? (idx / sym_block + 1) * sym_block : ld) : ld);
  const bool zero_owner = idx < 0 && rank == 0;  // invalid index: rank 0 fills zeros
  const int g = (rank + (int)blockIdx.z) % G;  // the peer this block stores to
  const i64 t = (i64)blockIdx.x * 256 + threadIdx.x;
  const i64 base = t * VN;
  bool mine = false;
  V16 v;
  memset(&v, 0, sizeof(v));
  if (base < ld) {
    if (base < cend) {
      mine = own || zero_owner;
      if (own) v = *reinterpret_cast<const V16*>(cov + gp * stride_q + (idx - row0) * ld + base);
    } else if (base >= row0 && base < row_end) {
      mine = true;
      T e[VN];
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        const i64 c = base + j;
        e[j] = c < N ? cov[gp * stride_q + (c - row0) * ld + idx] : T(0);
      }
      memcpy(&v, e, sizeof(v));
    }
  }
  const bool own_col = idx >= cand0 && idx < cand0 + n_cand;
  const bool col_mine = V != nullptr && t < m_pad && (own_col || zero_owner);
  // any data for this peer in this block?  (uniform per block except at piece boundaries)
  const int any = __syncthreads_or((mine || col_mine) ? 1 : 0);
  if (any) {
    // "peer g's row buffer is free": with COMBINE this block has already seen peer g's arg-max record of this
    // step, which the last block of g's fused_finish publishes AFTER every block of it has read kbuf / wbuf /
    // pcol of the previous exchange -- the record implies the buffer is free, no second wait over NVLink
    if constexpr (!COMBINE) {
      if (threadIdx.x == 0 && !spin_until(&me->ready[g], s)) me->error = TAL_PEER_ERR_READY_TIMEOUT;
      __syncthreads();
    }
    char* box = reinterpret_cast<char*>(boxes[g]);
    if (mine) *reinterpret_cast<V16*>(reinterpret_cast<T*>(box + kbuf_off) + (i64)gp * ld + base) = v;
    if (mine && own && idx >= base && idx < base + VN) {  // the thread that carries the diagonal element
      T e[VN];
      memcpy(e, &v, sizeof(v));
      hdr(boxes[g])->kdiag[gp] = (double)e[idx - base];
    }
    if (col_mine)
      reinterpret_cast<double*>(box + pcol_off)[(i64)gp * pcol_stride + t] =
          own_col ? V[gp * v_stride_q + t * ldv + (idx - cand0)] : 0.0;
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int tk = atomicAdd(&me->ticket, 1u);
    s_last = (tk == total - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if ((int)threadIdx.x < G) {
    __threadfence_system();
    if (threadIdx.x == 0) {
      me->ticket = 0;
      if constexpr (COMBINE) me->rec_step = s_rec;  // every block has read it by now
    }
    st_release_sys(&hdr(boxes[threadIdx.x])->row_seq[rank], s);
  }
}

// One warp: wait for the G pieces of exchange row_step + 1, then advance.
__global__ void peer_wait_row_kernel(void* mailbox, int G, const i64* __restrict__ idx_dev, i64 idx_host,
                                     const double* __restrict__ mean, i64 N, int q, double* __restrict__ scal) {
  PeerHeader* h = hdr(mailbox);
  const i64 idx = idx_dev ? *idx_dev : idx_host;
  const unsigned long long s = h->row_step + 1;
  bool ok = true;
  if ((int)threadIdx.x < G) ok = spin_until(&h->row_seq[threadIdx.x], s);
  ok = __all_sync(0xffffffffu, ok);
  if ((int)threadIdx.x < q && idx >= 0 && idx < N) scal[threadIdx.x] = mean[(i64)threadIdx.x * N + idx];
  if (threadIdx.x == 0) {
    if (!ok) h->error = TAL_PEER_ERR_ROW_TIMEOUT;
    h->row_step = s;
  }
}

static i64 align_up(i64 x, i64 a) { return (x + a - 1) / a * a; }

}  // namespace tal

namespace tal {

// combined != null: the records of this step are combined inside the kernel (see the kernel).
int peer_push_row_impl(void* const* mailboxes, int G, int rank, const long long* bounds_host, const void* cov,
                       long long cov_stride_q, long long ld, long long N, int q, int lower,
                       const long long* idx_dev, long long idx_host, long long kbuf_off, const double* V,
                       long long v_stride_q, long long ldv, long long m_pad, long long pcol_off,
                       long long pcol_stride, long long cand0, long long n_cand, int dtype,
                       tal_argmax_result* combined, const double* mean, double* scal, void* stream) {
  TAL_ARG(mailboxes && cov && bounds_host && G >= 1 && G <= kPeerMax && rank >= 0 && rank < G);
  TAL_ARG(q >= 1 && q <= TAL_MAX_Q && ld >= N && N >= 1 && ld % 4 == 0);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG(kbuf_off >= (long long)sizeof(PeerHeader) && kbuf_off % 16 == 0);
  TAL_ARG(V == nullptr || (m_pad >= 1 && pcol_off % 16 == 0 && pcol_stride >= m_pad));
  TAL_ARG(combined == nullptr || (mean && scal));
  ShardBounds bounds;
  for (int g = 0; g <= kPeerMax; ++g) bounds.b[g] = g <= G ? bounds_host[g] : bounds_host[G];
  TAL_ARG(bounds.b[0] == 0 && bounds.b[G] == N);
  for (int g = 0; g < G; ++g) TAL_ARG(bounds.b[g] <= bounds.b[g + 1] && bounds.b[g] % 4 == 0);
  const i64 vn = dtype == TAL_F64 ? 2 : 4;
  const i64 blk = lower ? tal_sym_block(dtype) : 0;
  i64 span = (ld + vn - 1) / vn;
  if (V && m_pad > span) span = m_pad;
  dim3 grid((unsigned)((span + 255) / 256), q, G);
#define TAL_PUSH(T, C)                                                                              \
  peer_push_row_kernel<T, C><<<grid, 256, 0, as_stream(stream)>>>(                                  \
      mailboxes, G, rank, bounds, (const T*)cov, cov_stride_q, ld, N, blk, idx_dev, idx_host, combined, mean, \
      scal, q, kbuf_off, V, v_stride_q, ldv, m_pad, pcol_off, pcol_stride, cand0, n_cand)
  if (dtype == TAL_F64) { if (combined) TAL_PUSH(double, true); else TAL_PUSH(double, false); }
  else { if (combined) TAL_PUSH(float, true); else TAL_PUSH(float, false); }
#undef TAL_PUSH
  TAL_LAUNCH_CHECK("tal_peer_push_row");
  return TAL_OK;
}

}  // namespace tal

using namespace tal;

extern "C" {

long long tal_peer_header_bytes(void) { return (long long)align_up(sizeof(PeerHeader), 256); }

int tal_peer_alloc(long long bytes, void** ptr_host) {
  TAL_ARG(bytes > 0 && ptr_host);
  void* p = nullptr;
  TAL_CUDA(cudaMalloc(&p, (size_t)bytes));
  TAL_CUDA(cudaMemset(p, 0, (size_t)bytes));
  TAL_CUDA(cudaDeviceSynchronize());
  *ptr_host = p;
  return TAL_OK;
}

int tal_peer_free(void* ptr) {
  if (ptr) TAL_CUDA(cudaFree(ptr));
  return TAL_OK;
}

int tal_peer_export(void* ptr, unsigned char* handle_host) {
  TAL_ARG(ptr && handle_host);
  static_assert(sizeof(cudaIpcMemHandle_t) == TAL_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  TAL_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_host, &h, sizeof(h));
  return TAL_OK;
}

int tal_peer_import(const unsigned char* handle_host, void** ptr_host) {
  TAL_ARG(handle_host && ptr_host);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  TAL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_host = p;
  return TAL_OK;
}

int tal_peer_unimport(void* ptr) {
  if (ptr) TAL_CUDA(cudaIpcCloseMemHandle(ptr));
  return TAL_OK;
}

int tal_peer_enable_access(int peer_device) {
  int dev = 0, can = 0;
  TAL_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev) return TAL_OK;
  TAL_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) {
    tal::set_error("tal_peer_enable_access: device %d cannot access device %d", dev, peer_device);
    return TAL_ERR_CUDA;
  }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
    tal::set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    return TAL_ERR_CUDA;
  }
  cudaGetLastError();
  return TAL_OK;
}

int tal_peer_publish_record(const tal_argmax_result* rec, void* const* mailboxes, int G, int rank,
                            void* stream) {
  TAL_ARG(rec && mailboxes && G >= 1 && G <= kPeerMax && rank >= 0 && rank < G);
  peer_publish_kernel<<<1, 32, 0, as_stream(stream)>>>(rec, mailboxes, G, rank);
  TAL_LAUNCH_CHECK("tal_peer_publish_record");
  return TAL_OK;
}

int tal_peer_combine_records(void* mailbox, int G, tal_argmax_result* out, void* stream) {
  TAL_ARG(mailbox && out && G >= 1 && G <= kPeerMax);
  peer_combine_kernel<<<1, 32, 0, as_stream(stream)>>>(mailbox, G, out);
  TAL_LAUNCH_CHECK("tal_peer_combine_records");
  return TAL_OK;
}

int tal_peer_push_row(void* const* mailboxes, int G, int rank, const long long* bounds_host, const void* cov,
                      long long cov_stride_q, long long ld, long long N, int q, int lower,
                      const long long* idx_dev, long long idx_host, long long kbuf_off,
                      const double* V, long long v_stride_q, long long ldv, long long m_pad,
                      long long pcol_off, long long pcol_stride, long long cand0, long long n_cand, int dtype,
                      void* stream) {
  return tal::peer_push_row_impl(mailboxes, G, rank, bounds_host, cov, cov_stride_q, ld, N, q, lower, idx_dev,
                                 idx_host, kbuf_off, V, v_stride_q, ldv, m_pad, pcol_off, pcol_stride, cand0,
                                 n_cand, dtype, nullptr, nullptr, nullptr, stream);
}

int tal_peer_wait_row(void* mailbox, int G, const long long* idx_dev, long long idx_host, const double* mean,
                      long long N, int q, double* scal, void* stream) {
  TAL_ARG(mailbox && mean && scal && q >= 1 && q <= TAL_MAX_Q && N >= 1 && G >= 1 && G <= kPeerMax);
  peer_wait_row_kernel<<<1, 32, 0, as_stream(stream)>>>(mailbox, G, idx_dev, idx_host, mean, N, q, scal);
  TAL_LAUNCH_CHECK("tal_peer_wait_row");
  return TAL_OK;
}

int tal_peer_error(const void* mailbox, int* error_host, void* stream) {
  TAL_ARG(mailbox && error_host);
  const PeerHeader* h = reinterpret_cast<const PeerHeader*>(mailbox);
  TAL_CUDA(cudaMemcpyAsync(error_host, &h->error, sizeof(int), cudaMemcpyDeviceToHost, as_stream(stream)));
  TAL_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return TAL_OK;
}

}  // extern "C"
