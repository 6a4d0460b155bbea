// common.cuh — shared device/host helpers for libpdlb200 (sm_100a only).
//
// Layout contract (lib/PDL/Core/pdlbroadcast.h:64): element of
// parameter p at broadcast index (i0,i1,...) and named-dim index n lives at
//     base_p + (offs_p + sum_k i_k*incs[k][p] + n*inc_n) elements.
// The host planner (plan.cpp) drops size-1 dims and merges adjacent dims whose
// strides chain for every parameter, so kernels see <= MAXD "collapsed" dims.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>
#include "../../include/pdlb200.h"

namespace pdlb200 {

constexpr int MAXD = 8;       // collapsed broadcast dims a kernel walks
constexpr int EW_THREADS = 256;

// ---- type traits -----------------------------------------------------------
template <class T> struct tt;
#define PDLB200_TT(T, ID, ISINT, ISUNS, WIDE, PLUS) \
  template <> struct tt<T> { static constexpr int id = ID; static constexpr bool is_int = ISINT; \
    static constexpr bool is_uns = ISUNS; using wide_u = WIDE; using plus = PLUS; };
// wide_u: unsigned type of width max(32, width(T)) — integer + - * << are done in it
// so that signed overflow wraps exactly like the reference's -fwrapv build.
// plus: the "int+" output type, max(long, T) in PDL's type order (PdlParObj.pm:149-159).
PDLB200_TT(int8_t,   PDLB200_SB,  true,  false, uint32_t, int32_t)
PDLB200_TT(uint8_t,  PDLB200_B,   true,  true,  uint32_t, int32_t)
PDLB200_TT(int16_t,  PDLB200_S,   true,  false, uint32_t, int32_t)
PDLB200_TT(uint16_t, PDLB200_US,  true,  true,  uint32_t, int32_t)
PDLB200_TT(int32_t,  PDLB200_L,   true,  false, uint32_t, int32_t)
PDLB200_TT(uint32_t, PDLB200_UL,  true,  true,  uint32_t, uint32_t)
PDLB200_TT(int64_t,  PDLB200_LL,  true,  false, uint64_t, int64_t)   // IND shares the C type
PDLB200_TT(uint64_t, PDLB200_ULL, true,  true,  uint64_t, uint64_t)
PDLB200_TT(float,    PDLB200_F,   false, false, uint32_t, float)
PDLB200_TT(double,   PDLB200_D,   false, false, uint64_t, double)
#undef PDLB200_TT

template <class T> __host__ __device__ __forceinline__ bool t_isnan(T v) {
  if constexpr (tt<T>::is_int) return false; else return v != v;
}
template <class T> __host__ __device__ __forceinline__ T from_bits(uint64_t bits) {
  if constexpr (sizeof(T) == 8) { union { uint64_t u; T t; } x; x.u = bits; return x.t; }
  else if constexpr (sizeof(T) == 4) { union { uint32_t u; T t; } x; x.u = (uint32_t)bits; return x.t; }
  else if constexpr (sizeof(T) == 2) { union { uint16_t u; T t; } x; x.u = (uint16_t)bits; return x.t; }
  else { union { uint8_t u; T t; } x; x.u = (uint8_t)bits; return x.t; }
}
// PDL_ISBAD2 (lib/PDL/Core/pdl.h.PL:261-262)
template <class T> __device__ __forceinline__ bool is_bad(T v, T badval, bool badnan) {
  if constexpr (tt<T>::is_int) return v == badval;   // PDL_ISNAN_<int type> is constant 0
  else return badnan ? t_isnan(v) : (v == badval);
}

// NaN results the way the reference's x86-64 SSE code produces them (IEEE 754 leaves the bits
// open; the GPU would return its canonical 0x7fffffff): a NaN operand is propagated, first
// operand first, with the quiet bit set; an invalid operation (inf-inf, 0*inf, 0/0, sqrt(-1))
// yields the x86 "default NaN", which is NEGATIVE quiet NaN.
template <class T> __device__ __forceinline__ T quiet_nan(T v) {
  if constexpr (sizeof(T) == 4) return __uint_as_float(__float_as_uint(v) | 0x00400000u);
  else return __longlong_as_double(__double_as_longlong(v) | 0x0008000000000000ll);
}
template <class T> __device__ __forceinline__ T x86_default_nan() {
  if constexpr (sizeof(T) == 4) return __uint_as_float(0xffc00000u);
  else return __longlong_as_double((long long)0xfff8000000000000ull);
}
template <class T> __device__ __forceinline__ T x86_nan2(T a, T b, T r) {
  if (r == r) return r;
  if (a != a) return quiet_nan(a);
  if (b != b) return quiet_nan(b);
  return x86_default_nan<T>();
}
template <class T> __device__ __forceinline__ T x86_nan1(T a, T r) {
  if (r == r) return r;
  if (a != a) return quiet_nan(a);
  return x86_default_nan<T>();
}

// 16-byte register image of VEC consecutive elements
template <class T> union alignas(16) Pack {
  uint4 q;
  T e[16 / sizeof(T)];
};

// ---- launch bookkeeping (api.cu) -------------------------------------------
void note_launch(const char *kernel_name);
int  sm_count();

struct Err {
  char *buf; size_t len;
  int fail(int code, const char *fmt, ...) const;
};
#define PDLB200_CUDA_OK(call, E) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return (E).fail(PDLB200_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

// ---- host-side collapsed view of a descriptor (plan.cpp) --------------------
struct Collapsed {
  int nd;                       // >= 1 (a lone size-1 dim when everything is scalar)
  int64_t dims[PDLB200_MAXDIMS];
  int64_t st[PDLB200_MAXPDLS][PDLB200_MAXDIMS]; // elements
  int64_t total;                // product of dims (0 if any dim is 0)
};
// Drop size-1 dims, merge dims d,d+1 when st[p][d+1] == st[p][d]*dims[d] for all p.
void collapse_dims(const pdlb200_trans *t, Collapsed *c);

// family launchers (one translation unit each)
int launch_elementwise(const pdlb200_trans *t, const Err &E);
int launch_convert(const pdlb200_trans *t, const Err &E);
int launch_ipow(const pdlb200_trans *t, const Err &E);
int launch_badops(const pdlb200_trans *t, const Err &E);
int launch_axisvalues(const pdlb200_trans *t, const Err &E);
int launch_inner(const pdlb200_trans *t, const Err &E);
int launch_magnover(const pdlb200_trans *t, const Err &E);
int launch_minmaximum(const pdlb200_trans *t, const Err &E);
int launch_reduce(const pdlb200_trans *t, const Err &E);
int launch_scan(const pdlb200_trans *t, const Err &E);
int launch_partial(const pdlb200_trans *t, const Err &E);
int launch_collapse(const pdlb200_trans *t, const Err &E);
int launch_nind(const pdlb200_trans *t, const Err &E);
int launch_complex(const pdlb200_trans *t, const Err &E);
int launch_matmult(const pdlb200_trans *t, const Err &E);

// per-device scratch for two-stage reductions (api.cu); grows, never shrinks
void *scratch(size_t nbytes, cudaStream_t s);

}  // namespace pdlb200
