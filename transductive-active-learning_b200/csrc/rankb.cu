// Temporally blocked posterior update: b rank-1 updates applied in ONE pass.
//
// Reference path: MarginalModel.step (lib/model/marginal.py:350-368) rewrites the
// N x N covariance after every observation, Sigma' = Sigma - k^T w
// (lib/gp/gaussian_distribution.py:36): 2 N^2 s bytes of HBM traffic per step for
// 2 N^2 flops.  The updates of consecutive steps commute with everything a step
// reads except the ONE row it observes, so they can be deferred: the (k_i, w_i) of
// the last p <= b observations are kept ("pending"), the observed row of a new
// step is read from the not-yet-updated matrix and corrected on the fly,
//
//   Sigma_t[x, c] = ((Sigma_base[x, c] - k_0[x] w_0[c]) - k_1[x] w_1[c]) - ...
//
// (tal_pending_correct_row; O(p N) work), and after b observations all of them are
// applied to every entry in one streaming pass (tal_rankb_update): the matrix
// traffic per step drops by b while each entry sees exactly the same sequence of
// rounded operations c <- c - rn(k_i[row] * w_i[col]) as b separate rank-1 passes:
// the result is bit-identical.  The pass stays HBM-bound up to b ~ 16
// (2 b flops per 16 bytes against 33.8 TFLOP/s of fp64 FMA issue).
#include "common.cuh"
#include <cstdlib>

namespace tal {

__device__ __forceinline__ i64 symb_cend(i64 grow, i64 block, i64 ld) {
  const i64 e = (grow / block + 1) * block;
  return e < ld ? e : ld;
}

template <typename T>
struct VecB;  // 16-byte vectors: P of them stay in registers per thread
template <>
struct VecB<double> {
  typedef double2 type;
  static constexpr int n = 2;
};
template <>
struct VecB<float> {
  typedef float4 type;
  static constexpr int n = 4;
};

template <typename T, int VN>
struct Lanes {
  T v[VN];
};

// grid: full storage (row blocks, column tiles, q); lower storage (column tiles, row blocks, q).
// K, W: [P][q][ld] pending vectors of which the first p are applied (slots >= p are not read).  One
// CTA walks ROWS rows of a 256-vector column tile: its P w-vectors stay in registers for the whole
// walk (they depend on the column only), the K values of its rows are staged in shared memory
// ([P][ROWS]), UNROLL rows of loads are in flight per thread.  Lower storage: a thread starts at the
// first row that keeps its column (the 32-row block that contains the column index).
template <typename T, int P, int UNROLL, bool FMA, int ROWS>
__global__ void __launch_bounds__(256, 2)
rankb_kernel(T* __restrict__ cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc, const T* __restrict__ K,
             const T* __restrict__ W, i64 pend_stride, int p, i64 sym_block) {
  typedef typename VecB<T>::type V;
  constexpr int VN = VecB<T>::n;
  __shared__ T ksm[P][ROWS];
  const int gp = blockIdx.z;
  const i64 rb = sym_block > 0 ? blockIdx.y : blockIdx.x;
  const i64 ct = sym_block > 0 ? blockIdx.x : blockIdx.y;
  const i64 col = (ct * 256 + threadIdx.x) * VN;
  const i64 r0 = rb * ROWS;
  const i64 r1 = min(n_loc, r0 + (i64)ROWS);
  if (sym_block > 0 && ct * 256 * VN >= symb_cend(row0 + r1 - 1, sym_block, ld)) return;  // whole tile above
  for (int e = threadIdx.x; e < P * ROWS; e += 256) {
    const int i = e / ROWS, r = e % ROWS;
    ksm[i][r] = (i < p && r0 + r < r1) ? K[(i64)i * pend_stride + (i64)gp * ld + row0 + r0 + r] : T(0);
  }
  __syncthreads();
  if (col >= ld) return;
  i64 rs = r0;
  if (sym_block > 0) {
    const i64 first = (col / sym_block) * sym_block - row0;  // rows before it keep this column in the mirror
    if (first > rs) rs = first;
  }
  if (rs >= r1) return;
  Lanes<T, VN> wv[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    if (i < p) {
      const V t = *reinterpret_cast<const V*>(W + (i64)i * pend_stride + (i64)gp * ld + col);
      memcpy(&wv[i], &t, sizeof(V));
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) wv[i].v[e] = T(0);
    }
  }
  T* C = cov + gp * stride_q + col;
  for (i64 r = rs; r < r1; r += UNROLL) {
    Lanes<T, VN> c[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j)
      if (r + j < r1) {
        const V t = ld_stream(reinterpret_cast<const V*>(C + (r + j) * ld));
        memcpy(&c[j], &t, sizeof(V));
      }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j)
      if (r + j < r1) {
#pragma unroll
        for (int i = 0; i < P; ++i) {
          const T kk = ksm[i][r + j - r0];
#pragma unroll
          for (int e = 0; e < VN; ++e) c[j].v[e] = sub_prod_m<FMA>(c[j].v[e], kk, wv[i].v[e]);
        }
        V t;
        memcpy(&t, &c[j], sizeof(V));
        st_stream(reinterpret_cast<V*>(C + (r + j) * ld), t);
      }
  }
}

// The same pass with the row loads decoupled from the registers: every thread streams ITS OWN 16-byte
// column slice through a private ring of D shared-memory slots with cp.async (no barrier: a thread only
// ever reads what it copied itself), so D rows per thread stay in flight while the fp64 chain of the
// current row runs.  At P = 16 the register-resident variant above holds only 4 rows in flight next to
// its 64 w-registers and is latency-bound (ncu: 62 % DRAM, 72 % of the cycles no warp eligible, long
// scoreboard); here the bytes in flight per SM are D x 4 KB x 2 CTAs regardless of the arithmetic.
__device__ __forceinline__ void rb_cp_async16(void* smem, const void* gmem) {
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void rb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void rb_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T, int P, bool FMA, int ROWS, int D>
__global__ void __launch_bounds__(256, (P > 16 ? 1 : 2))  // (32 w-vectors = 128 registers: one CTA per SM)
rankb_async_kernel(T* __restrict__ cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc, const T* __restrict__ K,
                   const T* __restrict__ W, i64 pend_stride, int p, i64 sym_block) {
  typedef typename VecB<T>::type V;
  constexpr int VN = VecB<T>::n;
  extern __shared__ __align__(16) unsigned char rb_smem[];
  V* ring = reinterpret_cast<V*>(rb_smem);                       // [D][256]
  T (*ksm)[ROWS] = reinterpret_cast<T (*)[ROWS]>(rb_smem + (size_t)D * 256 * sizeof(V));  // [P][ROWS]
  const int gp = blockIdx.z;
  const i64 rb = sym_block > 0 ? blockIdx.y : blockIdx.x;
  const i64 ct = sym_block > 0 ? blockIdx.x : blockIdx.y;
  const i64 col = (ct * 256 + threadIdx.x) * VN;
  const i64 r0 = rb * ROWS;
  const i64 r1 = min(n_loc, r0 + (i64)ROWS);
  if (sym_block > 0 && ct * 256 * VN >= symb_cend(row0 + r1 - 1, sym_block, ld)) return;  // whole tile above
  i64 rs = r0;
  if (sym_block > 0) {
    const i64 first = (col / sym_block) * sym_block - row0;
    if (first > rs) rs = first;
  }
  const bool active = col < ld && rs < r1;
  T* C = cov + gp * stride_q + col;
  V* mine = ring + threadIdx.x;
  if (active) {  // the first D rows are on their way while K and W are fetched
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (rs + d < r1) rb_cp_async16(mine + d * 256, C + (rs + d) * ld);
      rb_cp_async_commit();
    }
  }
  for (int e = threadIdx.x; e < P * ROWS; e += 256) {
    const int i = e / ROWS, r = e % ROWS;
    ksm[i][r] = (i < p && r0 + r < r1) ? K[(i64)i * pend_stride + (i64)gp * ld + row0 + r0 + r] : T(0);
  }
  __syncthreads();
  if (!active) return;
  Lanes<T, VN> wv[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    if (i < p) {
      const V t = *reinterpret_cast<const V*>(W + (i64)i * pend_stride + (i64)gp * ld + col);
      memcpy(&wv[i], &t, sizeof(V));
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) wv[i].v[e] = T(0);
    }
  }
  int slot = 0;
  for (i64 r = rs; r < r1; ++r) {
    rb_cp_async_wait<D - 1>();  // this thread's copy of row r has landed
    Lanes<T, VN> c;
    {
      const V t = mine[slot * 256];
      memcpy(&c, &t, sizeof(V));
    }
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const T kk = ksm[i][r - r0];
#pragma unroll
      for (int e = 0; e < VN; ++e) c.v[e] = sub_prod_m<FMA>(c.v[e], kk, wv[i].v[e]);
    }
    V t;
    memcpy(&t, &c, sizeof(V));
    st_stream(reinterpret_cast<V*>(C + r * ld), t);
    // the slot is free again (its value went through the arithmetic above): fetch row r + D into it
    if (r + D < r1) rb_cp_async16(mine + slot * 256, C + (r + D) * ld);
    rb_cp_async_commit();
    slot = slot + 1 == D ? 0 : slot + 1;
  }
}

// k[c] <- (((k[c] - K_0[a] W_0[b]) - K_1[a] W_1[b]) - ...) with (a, b) = (x, c) for the entries the row
// x keeps and (c, x) for the mirrored ones (lower-block storage): exactly what the deferred passes
// will do to those entries.  grid (ceil(ld / 256), q).
template <typename T, bool FMA>
__global__ void __launch_bounds__(256)
pending_correct_kernel(T* __restrict__ kbuf, i64 ld, i64 N, const T* __restrict__ K, const T* __restrict__ W,
                       i64 pend_stride, int p, const i64* __restrict__ idx_dev, i64 idx_host, i64 sym_block) {
  const int gp = blockIdx.y;
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ld) return;
  const i64 x = idx_dev ? *idx_dev : idx_host;
  if (x < 0 || x >= N) return;
  const bool mirrored = sym_block > 0 && c >= symb_cend(x, sym_block, ld);
  T v = kbuf[(i64)gp * ld + c];
  for (int i = 0; i < p; ++i) {
    const T* Ki = K + (i64)i * pend_stride + (i64)gp * ld;
    const T* Wi = W + (i64)i * pend_stride + (i64)gp * ld;
    v = mirrored ? sub_prod_m<FMA>(v, Ki[c], Wi[x]) : sub_prod_m<FMA>(v, Ki[x], Wi[c]);
  }
  kbuf[(i64)gp * ld + c] = v;
}

// Tuning knob (microbenchmarks only): TAL_RANKB_VARIANT = -1 (default, chosen from the measurements in
// profiles/r02_rankb_variants.md: the cp.async ring at P = 16, where the register-resident kernel is
// latency-bound -- ring of 16 rows with FMA arithmetic, of 8 with two roundings -- and the register-resident
// kernel with 128 rows per CTA below, which already runs at the copy roofline), 0 (register-resident, 128 rows
// per CTA, UNROLL 4 at P = 16 and 8 below), 1 (128 rows, UNROLL 8), 2 (256 rows, default UNROLL), 3 (256 rows,
// UNROLL 8), 4 (32 rows per CTA: the round-1 shape), 5 / 6 / 8 (cp.async ring of 8 / 16 / 24 rows per thread,
// 128 rows per CTA), 7 (ring of 16, 256 rows per CTA).
static int rankb_variant() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("TAL_RANKB_VARIANT");
    v = e ? atoi(e) : -1;
  }
  return v;
}

template <typename T, int P, bool FMA>
static int launch_rankb(T* cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc, int q, const T* K, const T* W,
                        i64 pend_stride, int p, i64 sym_block, cudaStream_t st) {
  constexpr int VN = VecB<T>::n;
  constexpr int UD = P >= 16 ? 4 : 8;  // P w-vectors + UNROLL rows in registers: 2 CTAs per SM
  const unsigned col_tiles = (unsigned)((ld + 256 * VN - 1) / (256 * VN));
#define TAL_RB(U, ROWS)                                                                              \
  do {                                                                                               \
    const unsigned row_blocks = (unsigned)((n_loc + ROWS - 1) / ROWS);                               \
    dim3 grid = sym_block > 0 ? dim3(col_tiles, row_blocks, q) : dim3(row_blocks, col_tiles, q);     \
    rankb_kernel<T, P, U, FMA, ROWS><<<grid, 256, 0, st>>>(cov, stride_q, ld, row0, n_loc, K, W, pend_stride, p, \
                                                           sym_block);                               \
  } while (0)
#define TAL_RBA(ROWS, D)                                                                             \
  do {                                                                                               \
    typedef typename VecB<T>::type V_;                                                               \
    const size_t smem = (size_t)D * 256 * sizeof(V_) + (size_t)P * ROWS * sizeof(T);                 \
    auto kern = rankb_async_kernel<T, P, FMA, ROWS, D>;                                              \
    static bool attr = false;                                                                        \
    if (!attr) {                                                                                     \
      TAL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      attr = true;                                                                                   \
    }                                                                                                \
    const unsigned row_blocks = (unsigned)((n_loc + ROWS - 1) / ROWS);                               \
    dim3 grid = sym_block > 0 ? dim3(col_tiles, row_blocks, q) : dim3(row_blocks, col_tiles, q);     \
    kern<<<grid, 256, smem, st>>>(cov, stride_q, ld, row0, n_loc, K, W, pend_stride, p, sym_block);  \
  } while (0)
  int v = rankb_variant();
  if (v < 0) v = P >= 16 ? (FMA ? 6 : 5) : 0;
  if constexpr (P > 16) {  // 32 updates per pass: ring variants only (the w-vectors alone fill the register file)
    if (v == 8) TAL_RBA(128, 24);
    else TAL_RBA(128, 16);
  } else {
    if (v == 5) TAL_RBA(128, 8);
    else if (v == 6) TAL_RBA(128, 16);
    else if (v == 7) TAL_RBA(256, 16);
    else if (v == 8) TAL_RBA(128, 24);
    else if (v == 1) TAL_RB(8, 128);
    else if (v == 2) TAL_RB(UD, 256);
    else if (v == 3) TAL_RB(8, 256);
    else if (v == 4) TAL_RB(UD, 32);
    else TAL_RB(UD, 128);
  }
#undef TAL_RB
#undef TAL_RBA
  return 0;
}

template <typename T, bool FMA>
static int rankb_dispatch(T* cov, i64 stride_q, i64 ld, i64 row0, i64 n_loc, int q, const T* K, const T* W,
                          i64 pend_stride, int p, i64 sym_block, cudaStream_t st) {
  if (p <= 2) return launch_rankb<T, 2, FMA>(cov, stride_q, ld, row0, n_loc, q, K, W, pend_stride, p, sym_block, st);
  if (p <= 4) return launch_rankb<T, 4, FMA>(cov, stride_q, ld, row0, n_loc, q, K, W, pend_stride, p, sym_block, st);
  if (p <= 8) return launch_rankb<T, 8, FMA>(cov, stride_q, ld, row0, n_loc, q, K, W, pend_stride, p, sym_block, st);
  if (p <= 16) return launch_rankb<T, 16, FMA>(cov, stride_q, ld, row0, n_loc, q, K, W, pend_stride, p, sym_block, st);
  return launch_rankb<T, 32, FMA>(cov, stride_q, ld, row0, n_loc, q, K, W, pend_stride, p, sym_block, st);
}

}  // namespace tal

using namespace tal;

extern "C" {

int tal_rankb_update(void* cov, long long cov_stride_q, long long ld, long long row0, long long n_loc,
                     long long N, int q, const void* K, const void* W, long long pend_stride, int p,
                     int lower, int dtype, void* stream) {
  TAL_ARG(cov && K && W);
  TAL_ARG(q >= 1 && q <= TAL_MAX_Q && N >= 1 && ld == (N + 31) / 32 * 32 && n_loc >= 0);
  TAL_ARG(p >= 1 && p <= TAL_MAX_PENDING && pend_stride >= (long long)q * ld);
  const bool fma_mode = arith_fma(dtype);
  dtype = base_dtype(dtype);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  TAL_ARG(row0 % 32 == 0);
  if (n_loc == 0) return TAL_OK;
  const i64 blk = lower ? tal_sym_block(dtype) : 0;
#define TAL_RBD(T, F)                                                                             \
  rankb_dispatch<T, F>((T*)cov, cov_stride_q, ld, row0, n_loc, q, (const T*)K, (const T*)W, pend_stride, p, blk, \
                       as_stream(stream))
  if (dtype == TAL_F64) { if (fma_mode) TAL_RBD(double, true); else TAL_RBD(double, false); }
  else { if (fma_mode) TAL_RBD(float, true); else TAL_RBD(float, false); }
#undef TAL_RBD
  TAL_LAUNCH_CHECK("tal_rankb_update");
  return TAL_OK;
}

int tal_pending_correct_row(void* kbuf, long long ld, long long N, int q, const void* K, const void* W,
                            long long pend_stride, int p, const long long* idx_dev, long long idx_host,
                            int lower, int dtype, void* stream) {
  TAL_ARG(kbuf && K && W && q >= 1 && q <= TAL_MAX_Q && N >= 1 && ld >= N);
  TAL_ARG(p >= 0 && p <= TAL_MAX_PENDING);
  const bool fma_mode = arith_fma(dtype);
  dtype = base_dtype(dtype);
  TAL_ARG(dtype == TAL_F64 || dtype == TAL_F32);
  if (p == 0) return TAL_OK;
  const i64 blk = lower ? tal_sym_block(dtype) : 0;
  dim3 grid((unsigned)((ld + 255) / 256), q);
#define TAL_PC(T, F)                                                                              \
  pending_correct_kernel<T, F><<<grid, 256, 0, as_stream(stream)>>>((T*)kbuf, ld, N, (const T*)K, (const T*)W, \
                                                                   pend_stride, p, idx_dev, idx_host, blk)
  if (dtype == TAL_F64) { if (fma_mode) TAL_PC(double, true); else TAL_PC(double, false); }
  else { if (fma_mode) TAL_PC(float, true); else TAL_PC(float, false); }
#undef TAL_PC
  TAL_LAUNCH_CHECK("tal_pending_correct_row");
  return TAL_OK;
}

}  // extern "C"
