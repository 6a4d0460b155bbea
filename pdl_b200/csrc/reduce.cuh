// reduce.cuh — PDL::Ufunc `a(n); [o]b()` reductions on the device.
//
// Replaces the generated loop(n) inside the two broadcast `for`s of
// pdl_<op>_readdata (lib/PDL/Ufunc.pd:88-118 sumover/prodover, :143-187 and/or/..over,
// :413-444 average, :446-500 minimum/maximum/_ind).  Roofline: HBM; algorithmic
// bytes per row = n*sizeof(T) read + sizeof(O) written.
//
// A "row" is one broadcast position; its n elements sit inc_n apart.  Three
// cooperation widths, picked by the host planner (reduce_plan.cu):
//   MODE 2  one CTA per row     (long rows: 128-bit loads, UNROLL in flight, smem tree)
//   MODE 1  one warp per row    (medium rows: shuffle tree only, no barrier)
//   MODE 0  one thread per row  (short rows, and "column" reductions where a broadcast
//                                dim is the unit-stride one: adjacent threads read
//                                adjacent rows, so the loads coalesce across the warp)
// When there are too few rows to fill 148 SMs each row is cut into chunks
// (blockIdx.y); partial accumulators go to scratch and a second tiny kernel merges
// them in chunk order.  All reducers are order-independent restatements of the
// reference's sequential loop (see each reducer), so integer, min/max and index
// results are bit-exact however the row is cut; float sums differ only by
// summation order.
#pragma once
#include "common.cuh"

namespace pdlb200 {

constexpr int RD_THREADS = 256;
constexpr int RD_UNROLL = 4;

struct RdPlan {
  const char *a; char *b;       // bases with offs applied
  int64_t n, inc_n;             // reduced dim: size, stride (elements)
  int64_t nrows;
  int64_t dims[MAXD];
  int64_t sa[MAXD], sb[MAXD];   // broadcast strides of a and b (elements)
  int64_t chunk;                // elements of n per chunk (== n when nchunks == 1)
  char *partial;                // scratch for nchunks > 1
  uint64_t abad, bbad;
  int nd;
  int nchunks;
  int abadnan;
  int badmode;                  // trans->bvalflag
};

// ---- accumulator shuffles ----------------------------------------------------
template <class A> __device__ __forceinline__ A shfl_xor_acc(const A &x, int mask) {
  static_assert(sizeof(A) % 4 == 0, "accumulator must be a multiple of 4 bytes");
  A r;
  const uint32_t *s = reinterpret_cast<const uint32_t *>(&x);
  uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(A) / 4); i++) d[i] = __shfl_xor_sync(0xffffffffu, s[i], mask);
  return r;
}

template <class O> __device__ __forceinline__ O wrap_add(O a, O b) {
  if constexpr (tt<O>::is_int) { using U = typename tt<O>::wide_u; return (O)((U)a + (U)b); } else return a + b;
}
template <class O> __device__ __forceinline__ O wrap_mul(O a, O b) {
  if constexpr (tt<O>::is_int) { using U = typename tt<O>::wide_u; return (O)((U)a * (U)b); } else return a * b;
}

// ---- reducers -------------------------------------------------------------------
// Each: Acc, init(), push(acc, value, n-index), merge(l, r) [commutative], finish(acc, n_good-known?, out...)

// sumover / dsumover: tmp += a over good elements; no good element in bad mode -> BAD (Ufunc.pd:102-110)
template <class T, class O> struct RSum {
  struct Acc { O s; int32_t any; int32_t pad; };
  static __device__ __forceinline__ Acc init() { Acc x; x.s = O(0); x.any = 0; x.pad = 0; return x; }
  static __device__ __forceinline__ void push(Acc &x, T v, int64_t) { x.s = wrap_add<O>(x.s, (O)v); x.any = 1; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { Acc x; x.s = wrap_add<O>(l.s, r.s); x.any = l.any | r.any; x.pad = 0; return x; }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    *out = (p.badmode && !x.any) ? from_bits<O>(p.bbad) : x.s;
  }
};
// prodover / dprodover (Ufunc.pd:91,102-110).  The reference leaves the loop once tmp == 0; for
// integers and finite floats the product is the same with or without the early exit.
template <class T, class O> struct RProd {
  struct Acc { O s; int32_t any; int32_t pad; };
  static __device__ __forceinline__ Acc init() { Acc x; x.s = O(1); x.any = 0; x.pad = 0; return x; }
  static __device__ __forceinline__ void push(Acc &x, T v, int64_t) { x.s = wrap_mul<O>(x.s, (O)v); x.any = 1; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { Acc x; x.s = wrap_mul<O>(l.s, r.s); x.any = l.any | r.any; x.pad = 0; return x; }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    *out = (p.badmode && !x.any) ? from_bits<O>(p.bbad) : x.s;
  }
};
// average / daverage (Ufunc.pd:417-430): tmp / cnt evaluated with C's usual arithmetic
// conversions (cnt is PDL_Indx = int64); cnt == 0 -> BAD (bad mode) or 0 / NaN (good mode).
template <class T, class O> struct RAvg {
  struct Acc { O s; int64_t cnt; };
  static __device__ __forceinline__ Acc init() { Acc x; x.s = O(0); x.cnt = 0; return x; }
  static __device__ __forceinline__ void push(Acc &x, T v, int64_t) { x.s = wrap_add<O>(x.s, (O)v); x.cnt++; }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) { Acc x; x.s = wrap_add<O>(l.s, r.s); x.cnt = l.cnt + r.cnt; return x; }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    if (x.cnt == 0) {
      if (p.badmode) *out = from_bits<O>(p.bbad);
      else if constexpr (tt<O>::is_int) *out = O(0);
      else *out = (O)__longlong_as_double(0x7ff8000000000000ll);  // NAN
      return;
    }
    if constexpr (!tt<O>::is_int) *out = x.s / (O)x.cnt;
    else if constexpr (sizeof(O) == 8 && tt<O>::is_uns) *out = (O)((uint64_t)x.s / (uint64_t)x.cnt);
    else *out = (O)((int64_t)x.s / x.cnt);
  }
};
// minimum / maximum / _ind (Ufunc.pd:455-465,481-491).  Sequential rule: cur is replaced when
// (a OP cur) or cur is NaN.  Closed form: the first-in-index-order extreme of the non-NaN good
// values; if every good value is NaN, the LAST NaN; if there is no good value, BAD.
template <class T, class O, bool ISMAX, bool WANT_IND> struct RMinMax {
  struct Acc { T cur; int64_t idx; int32_t state; int32_t pad; };  // state 0 empty, 1 non-NaN, 2 NaN only
  static __device__ __forceinline__ Acc init() { Acc x; x.cur = T(0); x.idx = -1; x.state = 0; x.pad = 0; return x; }
  static __device__ __forceinline__ bool better(T v, int64_t i, T cur, int64_t idx) {
    return (ISMAX ? (v > cur) : (v < cur)) || (v == cur && i < idx);
  }
  static __device__ __forceinline__ void push(Acc &x, T v, int64_t i) {
    if (t_isnan(v)) { if (x.state == 0 || (x.state == 2 && i > x.idx)) { x.cur = v; x.idx = i; x.state = 2; } }
    else if (x.state != 1 || better(v, i, x.cur, x.idx)) { x.cur = v; x.idx = i; x.state = 1; }
  }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) {
    if (r.state == 0) return l;
    if (l.state == 0) return r;
    if (l.state == 1 && r.state == 1) return better(r.cur, r.idx, l.cur, l.idx) ? r : l;
    if (l.state == 1) return l;
    if (r.state == 1) return r;
    return (r.idx > l.idx) ? r : l;
  }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, O *out) {
    if (x.state == 0) *out = from_bits<O>(p.bbad);
    else if constexpr (WANT_IND) *out = (O)x.idx;
    else *out = (O)x.cur;
  }
};
// andover orover zcover xorover (logical) and bandover borover bxorover (bitwise), Ufunc.pd:143-187.
// KIND: 0 and, 1 or, 2 zc, 3 xor, 4 band, 5 bor, 6 bxor.  Output type == input type.
template <class T, int KIND> struct RBits {
  struct Acc { typename tt<T>::wide_u v; int32_t any; };
  using U = typename tt<T>::wide_u;
  static __device__ __forceinline__ Acc init() {
    Acc x; x.any = 0;
    x.v = (KIND == 0 || KIND == 2) ? U(1) : (KIND == 4) ? ~U(0) : U(0);
    return x;
  }
  static __device__ __forceinline__ U bits(T a) { if constexpr (tt<T>::is_int) return (U)a; else return U(0); }
  static __device__ __forceinline__ void push(Acc &x, T a, int64_t) {
    x.any = 1;
    if constexpr (KIND == 0) x.v &= U(a != 0);
    else if constexpr (KIND == 1) x.v |= U(a != 0);
    else if constexpr (KIND == 2) x.v &= U(a == 0);
    else if constexpr (KIND == 3) x.v ^= U(a != 0);
    else if constexpr (KIND == 4) x.v &= bits(a);
    else if constexpr (KIND == 5) x.v |= bits(a);
    else x.v ^= bits(a);
  }
  static __device__ __forceinline__ Acc merge(const Acc &l, const Acc &r) {
    Acc x; x.any = l.any | r.any;
    if constexpr (KIND == 0 || KIND == 2 || KIND == 4) x.v = l.v & r.v;
    else if constexpr (KIND == 1 || KIND == 5) x.v = l.v | r.v;
    else x.v = l.v ^ r.v;
    return x;
  }
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, T *out) {
    *out = (p.badmode && !x.any) ? from_bits<T>(p.bbad) : (T)x.v;
  }
};

// ---- row walk -------------------------------------------------------------------
template <class R, class T, bool BAD>
__device__ __forceinline__ void rd_push(typename R::Acc &acc, T v, int64_t i, T abad, bool abadnan) {
  if constexpr (BAD) { if (is_bad(v, abad, abadnan)) return; }
  R::push(acc, v, i);
}

// Accumulate elements [lo, hi) of one row; `lane` of `width` cooperating threads.
template <class R, class T, bool BAD>
__device__ __forceinline__ void rd_row(typename R::Acc &acc, const T *row, int64_t lo, int64_t hi, int64_t inc,
                                       int lane, int width, T abad, bool abadnan) {
  constexpr int VEC = 16 / sizeof(T);
  if (inc == 1) {
    // peel to 16-byte alignment, then 128-bit loads with RD_UNROLL in flight, then the tail
    const uintptr_t addr = (uintptr_t)(row + lo);
    int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > hi - lo) head = hi - lo;
    for (int64_t i = lo + lane; i < lo + head; i += width) rd_push<R, T, BAD>(acc, row[i], i, abad, abadnan);
    const int64_t v0 = lo + head;
    const int64_t nv = (hi - v0) / VEC;
    const uint4 *vp = reinterpret_cast<const uint4 *>(row + v0);
    int64_t j = lane;
    for (; j + (int64_t)(RD_UNROLL - 1) * width < nv; j += (int64_t)RD_UNROLL * width) {
      Pack<T> r[RD_UNROLL];
#pragma unroll
      for (int u = 0; u < RD_UNROLL; u++) r[u].q = vp[j + (int64_t)u * width];
#pragma unroll
      for (int u = 0; u < RD_UNROLL; u++) {
        const int64_t e0 = v0 + (j + (int64_t)u * width) * VEC;
#pragma unroll
        for (int k = 0; k < VEC; k++) rd_push<R, T, BAD>(acc, r[u].e[k], e0 + k, abad, abadnan);
      }
    }
    for (; j < nv; j += width) {
      Pack<T> r; r.q = vp[j];
      const int64_t e0 = v0 + j * VEC;
#pragma unroll
      for (int k = 0; k < VEC; k++) rd_push<R, T, BAD>(acc, r.e[k], e0 + k, abad, abadnan);
    }
    for (int64_t i = v0 + nv * VEC + lane; i < hi; i += width) rd_push<R, T, BAD>(acc, row[i], i, abad, abadnan);
  } else {
    int64_t i = lo + lane;
    for (; i + (int64_t)(RD_UNROLL - 1) * width < hi; i += (int64_t)RD_UNROLL * width) {
      T v[RD_UNROLL];
#pragma unroll
      for (int u = 0; u < RD_UNROLL; u++) v[u] = row[(i + (int64_t)u * width) * inc];
#pragma unroll
      for (int u = 0; u < RD_UNROLL; u++) rd_push<R, T, BAD>(acc, v[u], i + (int64_t)u * width, abad, abadnan);
    }
    for (; i < hi; i += width) rd_push<R, T, BAD>(acc, row[i * inc], i, abad, abadnan);
  }
}

__device__ __forceinline__ void rd_row_offsets(const RdPlan &p, int64_t row, int64_t &oa, int64_t &ob) {
  oa = 0; ob = 0;
  for (int d = 0; d < p.nd; d++) {
    int64_t q, i;
    if (d == p.nd - 1) { i = row; q = 0; }
    else if ((uint64_t)row <= 0xffffffffull && (uint64_t)p.dims[d] <= 0xffffffffull) {
      const uint32_t q32 = (uint32_t)row / (uint32_t)p.dims[d]; q = q32; i = (uint32_t)row - q32 * (uint32_t)p.dims[d];
    } else { q = row / p.dims[d]; i = row - q * p.dims[d]; }
    oa += i * p.sa[d]; ob += i * p.sb[d];
    row = q;
  }
}

// MODE: 0 thread/row, 1 warp/row, 2 CTA/row.  blockIdx.y = chunk of n.
template <class R, class T, class O, bool BAD, int MODE>
__global__ void __launch_bounds__(RD_THREADS)
reduce_rows_kernel(const __grid_constant__ RdPlan p) {
  using Acc = typename R::Acc;
  const T abad = from_bits<T>(p.abad);
  const bool abadnan = p.abadnan != 0;
  const int chunk_id = blockIdx.y;
  const int64_t lo = (int64_t)chunk_id * p.chunk;
  const int64_t hi = (lo + p.chunk < p.n) ? lo + p.chunk : p.n;
  __shared__ Acc smem[RD_THREADS / 32];

  int64_t row, row_step; int lane, width;
  if (MODE == 0) { row = (int64_t)blockIdx.x * RD_THREADS + threadIdx.x; row_step = (int64_t)gridDim.x * RD_THREADS; lane = 0; width = 1; }
  else if (MODE == 1) { row = (int64_t)blockIdx.x * (RD_THREADS / 32) + (threadIdx.x >> 5); row_step = (int64_t)gridDim.x * (RD_THREADS / 32); lane = threadIdx.x & 31; width = 32; }
  else { row = blockIdx.x; row_step = gridDim.x; lane = threadIdx.x; width = RD_THREADS; }

  for (; row < p.nrows; row += row_step) {
    int64_t oa, ob;
    rd_row_offsets(p, row, oa, ob);
    Acc acc = R::init();
    rd_row<R, T, BAD>(acc, reinterpret_cast<const T *>(p.a) + oa, lo, hi, p.inc_n, lane, width, abad, abadnan);
    bool writer = true;
    if (MODE >= 1) {
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) acc = R::merge(acc, shfl_xor_acc(acc, m));
      writer = (lane & 31) == 0;
    }
    if (MODE == 2) {
      const int w = threadIdx.x >> 5;
      if ((threadIdx.x & 31) == 0) smem[w] = acc;
      __syncthreads();
      if (w == 0) {
        acc = (threadIdx.x < RD_THREADS / 32) ? smem[threadIdx.x] : R::init();
#pragma unroll
        for (int m = (RD_THREADS / 64); m >= 1; m >>= 1) acc = R::merge(acc, shfl_xor_acc(acc, m));
      }
      writer = threadIdx.x == 0;
      __syncthreads();  // smem reused by the next row
    }
    if (writer) {
      if (p.nchunks == 1) R::finish(acc, p, reinterpret_cast<O *>(p.b) + ob);
      else reinterpret_cast<Acc *>(p.partial)[row * p.nchunks + chunk_id] = acc;
    }
  }
}

// second stage: one warp per row merges that row's partials
template <class R, class O>
__global__ void __launch_bounds__(RD_THREADS)
reduce_finish_kernel(const __grid_constant__ RdPlan p) {
  using Acc = typename R::Acc;
  const int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * (RD_THREADS / 32) + (threadIdx.x >> 5);
  const int64_t row_step = (int64_t)gridDim.x * (RD_THREADS / 32);
  for (; row < p.nrows; row += row_step) {
    int64_t oa, ob;
    rd_row_offsets(p, row, oa, ob);
    Acc acc = R::init();
    const Acc *part = reinterpret_cast<const Acc *>(p.partial) + row * p.nchunks;
    for (int c = lane; c < p.nchunks; c += 32) acc = R::merge(acc, part[c]);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc = R::merge(acc, shfl_xor_acc(acc, m));
    if (lane == 0) R::finish(acc, p, reinterpret_cast<O *>(p.b) + ob);
  }
}

// host planner (reduce_plan.cu)
struct RdLaunch { int mode; dim3 grid; };
int rd_build_plan(const pdlb200_trans *t, size_t in_size, size_t out_size, size_t acc_size,
                  RdPlan *p, RdLaunch *l, const Err &E);

template <class R, class T, class O>
int rd_launch_typed(const pdlb200_trans *t, const char *name, const Err &E) {
  RdPlan p; RdLaunch l;
  int rc = rd_build_plan(t, sizeof(T), sizeof(O), sizeof(typename R::Acc), &p, &l, E);
  if (rc) return rc;
  if (p.nrows == 0) return PDLB200_OK;
  cudaStream_t s = (cudaStream_t)t->stream;
#define PDLB200_RD_GO(BADF, MODE) reduce_rows_kernel<R, T, O, BADF, MODE><<<l.grid, RD_THREADS, 0, s>>>(p)
  if (t->bvalflag) { if (l.mode == 0) PDLB200_RD_GO(true, 0); else if (l.mode == 1) PDLB200_RD_GO(true, 1); else PDLB200_RD_GO(true, 2); }
  else             { if (l.mode == 0) PDLB200_RD_GO(false, 0); else if (l.mode == 1) PDLB200_RD_GO(false, 1); else PDLB200_RD_GO(false, 2); }
#undef PDLB200_RD_GO
  note_launch(name);
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  if (p.nchunks > 1) {
    int64_t g = (p.nrows + RD_THREADS / 32 - 1) / (RD_THREADS / 32);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    reduce_finish_kernel<R, O><<<(int)g, RD_THREADS, 0, s>>>(p);
    note_launch("reduce_finish");
    PDLB200_CUDA_OK(cudaGetLastError(), E);
  }
  return PDLB200_OK;
}

}  // namespace pdlb200
