// Shared helpers for the sm_100a kernels behind include/tal_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include "../../include/tal_b200.h"

typedef long long i64;

namespace tal {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Check the launch that was just enqueued.
#define TAL_LAUNCH_CHECK(name)                                                  \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      tal::set_error("%s: %s", name, cudaGetErrorString(e__));                  \
      return TAL_ERR_CUDA;                                                      \
    }                                                                           \
    tal::count_launch();                                                        \
  } while (0)

#define TAL_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      tal::set_error("%s: %s", #call, cudaGetErrorString(e__));                 \
      return TAL_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

#define TAL_ARG(cond)                                                           \
  do {                                                                          \
    if (!(cond)) {                                                              \
      tal::set_error("%s: invalid argument: %s", __func__, #cond);              \
      return TAL_ERR_ARG;                                                       \
    }                                                                           \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// dtype argument of the update entry points: TAL_F64 / TAL_F32, optionally | TAL_ARITH_FMA
static inline int base_dtype(int dtype) { return dtype & ~TAL_ARITH_FMA; }
static inline bool arith_fma(int dtype) { return (dtype & TAL_ARITH_FMA) != 0; }

constexpr int kNumSM = 148;  // B200

struct BetaQ {
  double v[TAL_MAX_Q];
};

// ---- arithmetic that must not be contracted into FMA ---------------------
// The reference subtracts a rounded product (XLA dot_general with K=1, then
// subtract: gaussian_distribution.py:36).  Using the same two roundings in
// the matrix kernel and in the variance vector keeps var == diag(cov) bit
// for bit and matches a non-fused CPU evaluation.
__device__ __forceinline__ double sub_prod(double c, double a, double b) {
  return __dsub_rn(c, __dmul_rn(a, b));
}
__device__ __forceinline__ float sub_prod(float c, float a, float b) {
  return __fsub_rn(c, __fmul_rn(a, b));
}
// Arithmetic mode of the covariance update (TAL_ARITH_FMA in the dtype argument of the update
// entry points): FMA = false keeps the two roundings above; FMA = true is c - a*b with ONE
// rounding (what a fused multiply-subtract loop computes), half the fp64 issue slots.  All
// kernels that touch the matrix, the pending rows and the variance vector use the same mode,
// so var == diag(cov) and deferred == per-step hold bit for bit in either mode.
template <bool FMA>
__device__ __forceinline__ double sub_prod_m(double c, double a, double b) {
  if constexpr (FMA) return fma(-a, b, c);
  else return __dsub_rn(c, __dmul_rn(a, b));
}
template <bool FMA>
__device__ __forceinline__ float sub_prod_m(float c, float a, float b) {
  if constexpr (FMA) return fmaf(-a, b, c);
  else return __fsub_rn(c, __fmul_rn(a, b));
}
__device__ __forceinline__ double add_prod(double c, double a, double b) {
  return __dadd_rn(c, __dmul_rn(a, b));
}

// jnp.maximum / jnp.minimum propagate NaN.
__device__ __forceinline__ double max_nan(double a, double b) {
  return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000LL) : (a > b ? a : b);
}
__device__ __forceinline__ double min_nan(double a, double b) {
  return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000LL) : (a < b ? a : b);
}

// ---- streaming global access (read-once / write-once data) ---------------
template <typename T>
struct Vec16;  // 16-byte vector of T
template <>
struct Vec16<double> {
  typedef double2 type;
  static constexpr int n = 2;
};
template <>
struct Vec16<float> {
  typedef float4 type;
  static constexpr int n = 4;
};

__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 r;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(double2* p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y)
               : "memory");
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// ---- reductions ------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum, deterministic; result valid in every thread.  `sh` holds
// at least 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

template <typename T>
__device__ __forceinline__ T load_as(const void* p, i64 i);
template <>
__device__ __forceinline__ double load_as<double>(const void* p, i64 i) {
  return reinterpret_cast<const double*>(p)[i];
}
template <>
__device__ __forceinline__ float load_as<float>(const void* p, i64 i) {
  return reinterpret_cast<const float*>(p)[i];
}

}  // namespace tal
