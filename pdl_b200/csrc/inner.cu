// inner.cu — inner, lib/PDL/Primitive.pd:48-70: a(n); b(n); [o]c();  c = sum_n a*b.
// This is the fused form of cfg3's `($a * $b)->sumover`: the [N,M] product never
// exists, so the traffic is the two operands, not 17 GB.  Roofline: HBM when the operands are big
// (a dot product of two long vectors reads 2*sizeof(T) per element); with dummy-dim operands it is
// L1/L2-resident and bound by the multiply-add rate.
//
// Reference arithmetic: each product is formed in the C type of the operands (float*float rounds to
// float, sub-int types multiply as int, 32/64-bit integers wrap under -fwrapv) and ADDED INTO A
// `long double` (x87 80-bit), which is converted to T once at the end.  The device has no 80-bit
// type: float rows accumulate in double (every float product is exact in double and the 53-bit sum
// then differs from the 64-bit one only beyond float precision), double rows accumulate in a
// compensated (two-sum) double pair, which carries more than the x87's 64 bits; integer rows
// accumulate in 64-bit integers, exact whenever the reference's sum is (|sum| < 2^63).
// BAD: any BAD a or b in the row makes c BAD (Primitive.pd:56-60), unlike sumover which skips them.
//
// Work split: a row is cut into `nchunks` chunks, one warp per (row, chunk); partial sums go through
// scratch and a small finishing kernel when nchunks > 1.  Many short rows take one thread per row.
#include <cstring>
#include "common.cuh"

namespace pdlb200 {

struct InPlan {
  const char *a, *b; char *c;
  char *part;                    // [nrows][nchunks] partials (Acc) when nchunks > 1
  int64_t dims[MAXD], sa[MAXD], sb[MAXD], sc[MAXD];
  int64_t n, inc_a, inc_b, nrows, nchunks, chunk;
  uint64_t abad, bbad, cbad;
  int nd, badmode, abadnan, bbadnan;
};

// accumulator per element type
template <class T, class = void> struct InAcc;
template <class T> struct InAcc<T, std::enable_if_t<tt<T>::is_int && !tt<T>::is_uns>> {
  long long s; int bad;
  __device__ void init() { s = 0; bad = 0; }
  __device__ void add(T a, T b) {
    if constexpr (sizeof(T) < 8) s += (long long)(int)((unsigned)(int)a * (unsigned)(int)b);   // product in int, wrapping
    else s = (long long)((unsigned long long)s + (unsigned long long)a * (unsigned long long)b);
  }
  __device__ void merge(const InAcc &o) { s = (long long)((unsigned long long)s + (unsigned long long)o.s); bad |= o.bad; }
  __device__ T result() const { return (T)s; }
};
template <class T> struct InAcc<T, std::enable_if_t<tt<T>::is_int && tt<T>::is_uns>> {
  unsigned long long s; int bad;
  __device__ void init() { s = 0; bad = 0; }
  __device__ void add(T a, T b) {
    if constexpr (sizeof(T) < 4) s += (unsigned long long)(long long)((int)a * (int)b);          // promoted to int
    else if constexpr (sizeof(T) == 4) s += (unsigned long long)((unsigned)a * (unsigned)b);     // unsigned wrap
    else s += (unsigned long long)a * (unsigned long long)b;
  }
  __device__ void merge(const InAcc &o) { s += o.s; bad |= o.bad; }
  __device__ T result() const { return (T)s; }
};
template <> struct InAcc<float> {
  double s; int bad;
  __device__ void init() { s = 0; bad = 0; }
  __device__ void add(float a, float b) { s += (double)(a * b); }
  __device__ void merge(const InAcc &o) { s += o.s; bad |= o.bad; }
  __device__ float result() const { return (float)s; }
};
template <> struct InAcc<double> {
  double s, c; int bad;           // s + c is the running sum (two-sum compensation)
  __device__ void init() { s = 0; c = 0; bad = 0; }
  __device__ void addv(double x) {
    const double t = s + x;
    const double bp = t - s;
    c += (s - (t - bp)) + (x - bp);
    s = t;
  }
  __device__ void add(double a, double b) { addv(a * b); }
  __device__ void merge(const InAcc &o) { addv(o.s); c += o.c; bad |= o.bad; }
  // a non-finite running sum (an inf element, or overflow) leaves NaN in the compensation term (inf - inf): the
  // reference's long double accumulator just carries the inf/NaN, so does `s`
  __device__ double result() const { return isfinite(s) ? s + c : s; }
};

template <class A> __device__ __forceinline__ A shfl_down_acc(const A &v, int d) {
  A r;
  constexpr int NW = (sizeof(A) + 3) / 4;
  unsigned w[NW], o[NW];
  memcpy(w, &v, sizeof(A));
#pragma unroll
  for (int k = 0; k < NW; k++) o[k] = __shfl_down_sync(0xffffffffu, w[k], d);
  memcpy(&r, o, sizeof(A));
  return r;
}

template <class T>
__device__ __forceinline__ void in_row_offsets(const InPlan &p, int64_t row, int64_t &oa, int64_t &ob, int64_t &oc) {
  oa = ob = oc = 0;
  for (int d = 0; d < p.nd; d++) {
    const int64_t q = (d == p.nd - 1) ? 0 : row / p.dims[d];
    const int64_t i = row - q * p.dims[d];
    oa += i * p.sa[d]; ob += i * p.sb[d]; oc += i * p.sc[d];
    row = q;
  }
}

// MAGN = magnover (lib/PDL/Ufunc.pd:1235-1256): b = a, BAD elements are SKIPPED (acc.bad then counts the good
// ones), a row without a good element is BAD, and the result is sum == 0 ? 0 : sqrt(sum).
template <class T, bool MAGN>
__device__ __forceinline__ void in_write(const InPlan &p, T *out, const InAcc<T> &acc) {
  if constexpr (MAGN) {
    if (p.badmode && !acc.bad) { *out = from_bits<T>(p.cbad); return; }
    if constexpr (tt<T>::is_int) *out = acc.result();
    else {
      double sum;
      if constexpr (sizeof(T) == 8) sum = acc.result(); else sum = acc.s;
      *out = sum == 0 ? T(0) : (T)sqrt(sum);
    }
  } else {
    *out = (p.badmode && acc.bad) ? from_bits<T>(p.cbad) : acc.result();
  }
}
template <class T, bool MAGN>
__device__ __forceinline__ void in_elem(const InPlan &p, InAcc<T> &acc, T va, T vb, T abad, T bbad) {
  if constexpr (MAGN) {
    if (p.badmode && is_bad(va, abad, p.abadnan != 0)) return;
    acc.bad = 1; acc.add(va, va);
  } else {
    if (p.badmode && (is_bad(va, abad, p.abadnan != 0) || is_bad(vb, bbad, p.bbadnan != 0))) acc.bad = 1;
    else acc.add(va, vb);
  }
}

// one warp per (row, chunk)
template <class T, bool MAGN>
__global__ void __launch_bounds__(256) inner_warp_kernel(const __grid_constant__ InPlan p) {
  const T abad = from_bits<T>(p.abad), bbad = from_bits<T>(p.bbad);
  const int lane = threadIdx.x & 31;
  const int64_t nwork = p.nrows * p.nchunks;
  for (int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); w < nwork; w += (int64_t)gridDim.x * 8) {
    const int64_t row = w / p.nchunks, chunk = w - row * p.nchunks;
    int64_t oa, ob, oc;
    in_row_offsets<T>(p, row, oa, ob, oc);
    const T *pa = reinterpret_cast<const T *>(p.a) + oa;
    const T *pb = reinterpret_cast<const T *>(p.b) + ob;
    const int64_t lo = chunk * p.chunk, hi = (lo + p.chunk < p.n) ? lo + p.chunk : p.n;
    InAcc<T> acc; acc.init();
    constexpr int U = 4;
    int64_t n = lo + lane;
    // unit-stride, 16-byte aligned chunk: 128-bit loads, U per operand in flight per lane
    constexpr int VEC = 16 / sizeof(T);
    if (p.inc_a == 1 && p.inc_b == 1 && ((((uintptr_t)(pa + lo)) | ((uintptr_t)(pb + lo))) & 15) == 0) {
      const int64_t nvec = (hi - lo) / VEC;
      const uint4 *qa = reinterpret_cast<const uint4 *>(pa + lo), *qb = reinterpret_cast<const uint4 *>(pb + lo);
      for (int64_t v0 = 0; v0 < nvec; v0 += 32 * U) {
        Pack<T> ra[U], rb[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const int64_t j = v0 + u * 32 + lane; if (j < nvec) { ra[u].q = qa[j]; if (!MAGN) rb[u].q = qb[j]; } }
#pragma unroll
        for (int u = 0; u < U; u++) {
          if (v0 + u * 32 + lane < nvec) {
#pragma unroll
            for (int k = 0; k < VEC; k++) in_elem<T, MAGN>(p, acc, ra[u].e[k], MAGN ? ra[u].e[k] : rb[u].e[k], abad, bbad);
          }
        }
      }
      n = lo + nvec * VEC + lane;     // tail elements go through the scalar loop below
    }
    // an operand that does not move along n (dummy dim: the cfg3 outer-product shape) is read once per row
    const bool a_const = p.inc_a == 0, b_const = p.inc_b == 0;
    const T a0 = (a_const && p.n > 0) ? pa[0] : T(0), b0 = (b_const && p.n > 0) ? pb[0] : T(0);
    for (; n + (U - 1) * 32 < hi; n += U * 32) {
      T va[U], vb[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        va[u] = a_const ? a0 : pa[(n + u * 32) * p.inc_a];
        vb[u] = MAGN ? va[u] : (b_const ? b0 : pb[(n + u * 32) * p.inc_b]);
      }
#pragma unroll
      for (int u = 0; u < U; u++) in_elem<T, MAGN>(p, acc, va[u], vb[u], abad, bbad);
    }
    for (; n < hi; n += 32) {
      const T va = a_const ? a0 : pa[n * p.inc_a], vb = MAGN ? va : (b_const ? b0 : pb[n * p.inc_b]);
      in_elem<T, MAGN>(p, acc, va, vb, abad, bbad);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { const InAcc<T> o = shfl_down_acc(acc, d); acc.merge(o); }
    if (lane == 0) {
      if (p.nchunks == 1) in_write<T, MAGN>(p, reinterpret_cast<T *>(p.c) + oc, acc);
      else reinterpret_cast<InAcc<T> *>(p.part)[w] = acc;
    }
  }
}

// finishing pass: one CTA per row; threads stride over the chunk partials, then warp shuffles + shared memory
template <class T, bool MAGN>
__global__ void __launch_bounds__(256) inner_finish_kernel(const __grid_constant__ InPlan p) {
  __shared__ InAcc<T> sh[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t row = blockIdx.x; row < p.nrows; row += gridDim.x) {
    const InAcc<T> *part = reinterpret_cast<const InAcc<T> *>(p.part) + row * p.nchunks;
    InAcc<T> acc; acc.init();
    for (int64_t k = threadIdx.x; k < p.nchunks; k += 256) acc.merge(part[k]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { const InAcc<T> o = shfl_down_acc(acc, d); acc.merge(o); }
    if (lane == 0) sh[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < 8; k++) acc.merge(sh[k]);
      int64_t oa, ob, oc;
      in_row_offsets<T>(p, row, oa, ob, oc);
      in_write<T, MAGN>(p, reinterpret_cast<T *>(p.c) + oc, acc);
    }
    __syncthreads();
  }
}

// many short rows: one thread per row
template <class T, bool MAGN>
__global__ void __launch_bounds__(256) inner_thread_kernel(const __grid_constant__ InPlan p) {
  const T abad = from_bits<T>(p.abad), bbad = from_bits<T>(p.bbad);
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.nrows; row += (int64_t)gridDim.x * blockDim.x) {
    int64_t oa, ob, oc;
    in_row_offsets<T>(p, row, oa, ob, oc);
    const T *pa = reinterpret_cast<const T *>(p.a) + oa;
    const T *pb = reinterpret_cast<const T *>(p.b) + ob;
    InAcc<T> acc; acc.init();
    for (int64_t n = 0; n < p.n; n++) {
      const T va = pa[n * p.inc_a], vb = MAGN ? va : pb[n * p.inc_b];
      in_elem<T, MAGN>(p, acc, va, vb, abad, bbad);
    }
    in_write<T, MAGN>(p, reinterpret_cast<T *>(p.c) + oc, acc);
  }
}

template <class T, bool MAGN = false>
static int inner_go(InPlan &p, cudaStream_t s, const Err &E) {
  const char *name = MAGN ? "magnover" : "inner";
  const int64_t cap = (int64_t)sm_count() * 8;
  if (p.n < 64 && p.nrows >= 1024) {
    int64_t g = (p.nrows + 255) / 256;
    if (g > cap) g = cap;
    inner_thread_kernel<T, MAGN><<<(int)g, 256, 0, s>>>(p);
  } else {
    // about one resident wave of warps (64 per SM) over all rows; chunks are multiples of 1024 elements
    int64_t want = cap * 8 / (p.nrows > 0 ? p.nrows : 1);
    if (want < 1) want = 1;
    int64_t chunk = (p.n + want - 1) / want;
    chunk = (chunk + 1023) / 1024 * 1024;
    if (chunk < 1024) chunk = 1024;
    p.chunk = chunk;
    p.nchunks = p.n > 0 ? (p.n + chunk - 1) / chunk : 1;
    if (p.nchunks > 1) {
      p.part = (char *)scratch((size_t)(p.nrows * p.nchunks) * sizeof(InAcc<T>), s);
      if (!p.part) return E.fail(PDLB200_ECUDA, "inner: cannot allocate scratch");
    }
    int64_t g = (p.nrows * p.nchunks + 7) / 8;
    if (g > cap * 4) g = cap * 4;
    inner_warp_kernel<T, MAGN><<<(int)g, 256, 0, s>>>(p);
    if (p.nchunks > 1) {
      int64_t g2 = p.nrows < cap * 4 ? p.nrows : cap * 4;
      inner_finish_kernel<T, MAGN><<<(int)g2, 256, 0, s>>>(p);
      note_launch(name);
    }
  }
  note_launch(name);
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

int launch_inner(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 3) return E.fail(PDLB200_EINVAL, "inner: expected 3 parameters");
  const size_t sz = pdlb200_type_size(t->datatype);
  if (!sz) return E.fail(PDLB200_EUNSUPPORTED, "inner: type %d is not on the device path", t->datatype);
  Collapsed c;
  collapse_dims(t, &c);
  if (c.total == 0) return PDLB200_OK;
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "inner: %d non-mergeable broadcast dims exceed the device walker's %d", c.nd, MAXD);
  InPlan p{};
  p.n = t->ind[0]; p.inc_a = t->rinc[0]; p.inc_b = t->rinc[1];
  if (p.n < 0) return E.fail(PDLB200_EINVAL, "inner: n = %lld", (long long)p.n);
  for (int k = 0; k < 3; k++)
    if (!t->pdls[k].data && (k == 2 || p.n > 0)) return E.fail(PDLB200_EINVAL, "inner: parameter %d got NULL data", k);
  p.a = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)sz;
  p.b = (const char *)t->pdls[1].data + t->pdls[1].offs * (int64_t)sz;
  p.c = (char *)t->pdls[2].data + t->pdls[2].offs * (int64_t)sz;
  p.nd = c.nd; p.nrows = c.total; p.nchunks = 1;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sa[d] = c.st[0][d]; p.sb[d] = c.st[1][d]; p.sc[d] = c.st[2][d]; }
  p.abad = t->pdls[0].badval; p.bbad = t->pdls[1].badval; p.cbad = t->pdls[2].badval;
  p.badmode = t->bvalflag != 0;
  p.abadnan = (t->pdls[0].flags & PDLB200_PAR_BADNAN) != 0;
  p.bbadnan = (t->pdls[1].flags & PDLB200_PAR_BADNAN) != 0;
  cudaStream_t s = (cudaStream_t)t->stream;
  switch (t->datatype) {
    case PDLB200_SB: return inner_go<int8_t>(p, s, E);   case PDLB200_B:  return inner_go<uint8_t>(p, s, E);
    case PDLB200_S:  return inner_go<int16_t>(p, s, E);  case PDLB200_US: return inner_go<uint16_t>(p, s, E);
    case PDLB200_L:  return inner_go<int32_t>(p, s, E);  case PDLB200_UL: return inner_go<uint32_t>(p, s, E);
    case PDLB200_IND: case PDLB200_LL: return inner_go<int64_t>(p, s, E);
    case PDLB200_ULL: return inner_go<uint64_t>(p, s, E);
    case PDLB200_F:  return inner_go<float>(p, s, E);    case PDLB200_D:  return inner_go<double>(p, s, E);
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "inner: type %d is not on the device path", t->datatype);
}

// magnover, lib/PDL/Ufunc.pd:1235-1256 : a(n); real [o]b().  GenericTypes D LD C* F ("F last"), so integer input
// is converted to float by the caller; device types F and D.  The reference sums a*a in long double and takes
// sqrtl; here the float sum is carried in double and the double sum in a compensated pair (see above).
int launch_magnover(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 2) return E.fail(PDLB200_EINVAL, "magnover: expected 2 parameters");
  if (t->datatype != PDLB200_F && t->datatype != PDLB200_D)
    return E.fail(PDLB200_EUNSUPPORTED, "magnover: type %d is not on the device path (float and double are)", t->datatype);
  const size_t sz = pdlb200_type_size(t->datatype);
  Collapsed c;
  collapse_dims(t, &c);
  if (c.total == 0) return PDLB200_OK;
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "magnover: %d non-mergeable broadcast dims exceed the device walker's %d", c.nd, MAXD);
  InPlan p{};
  p.n = t->ind[0]; p.inc_a = p.inc_b = t->rinc[0];
  if (p.n < 0) return E.fail(PDLB200_EINVAL, "magnover: n = %lld", (long long)p.n);
  if ((!t->pdls[0].data && p.n > 0) || !t->pdls[1].data) return E.fail(PDLB200_EINVAL, "magnover: NULL data");
  p.a = p.b = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)sz;
  p.c = (char *)t->pdls[1].data + t->pdls[1].offs * (int64_t)sz;
  p.nd = c.nd; p.nrows = c.total; p.nchunks = 1;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sa[d] = p.sb[d] = c.st[0][d]; p.sc[d] = c.st[1][d]; }
  p.abad = p.bbad = t->pdls[0].badval; p.cbad = t->pdls[1].badval;
  p.badmode = t->bvalflag != 0;
  p.abadnan = p.bbadnan = (t->pdls[0].flags & PDLB200_PAR_BADNAN) != 0;
  cudaStream_t s = (cudaStream_t)t->stream;
  if (t->datatype == PDLB200_F) return inner_go<float, true>(p, s, E);
  return inner_go<double, true>(p, s, E);
}
}  // namespace pdlb200
