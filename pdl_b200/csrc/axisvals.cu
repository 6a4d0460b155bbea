// axisvals.cu — axisvalues, lib/PDL/Primitive.pd:1468-1474: i(n); [o]a(n); `loop(n) %{ $a() = n; %}`.
// The body of sequence / xvals / yvals / zvals (lib/PDL/Basic.pm:117-129,479-485): keeps the
// constructors of every benchmark script on the device.  Write-only, HBM-bound:
// algorithmic bytes = sizeof(T) per element.
//
// The named dim n is put in front of the broadcast dims and the whole thing collapsed like any other
// walk; collapsing can only merge HIGHER dims into dim 0, so the value at collapsed position p of dim 0
// is p mod n.  A thread owns 16 bytes of consecutive positions: one modulo, then count-and-wrap.
#include "common.cuh"

namespace pdlb200 {

struct AxPlan {
  char *a;
  int64_t dims[MAXD], st[MAXD];
  int64_t n;            // size of the named dim
  int64_t vpr, n_units; // units per row of collapsed dim 0, total units
  int nd, vec;
};

template <class T>
__global__ void __launch_bounds__(256) axisvals_kernel(const __grid_constant__ AxPlan p) {
  constexpr int VEC = 16 / sizeof(T);
  T *base = reinterpret_cast<T *>(p.a);
  constexpr int U = 4;   // independent 128-bit stores per thread per trip
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t u0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u0 < p.n_units; u0 += nthreads * U) {
#pragma unroll
   for (int j = 0; j < U; j++) {
    const int64_t u = u0 + j * nthreads;
    if (u >= p.n_units) break;
    int64_t row, v;
    if (p.nd == 1) { row = 0; v = u; }
    else if ((uint64_t)p.n_units <= 0xffffffffull) { const uint32_t r = (uint32_t)u / (uint32_t)p.vpr; row = r; v = (uint32_t)u - r * (uint32_t)p.vpr; }
    else { row = u / p.vpr; v = u - row * p.vpr; }
    const int64_t i0 = v * VEC;
    int64_t off = i0 * p.st[0];
    for (int d = 1; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : row / p.dims[d];
      off += (row - q * p.dims[d]) * p.st[d];
      row = q;
    }
    const int64_t left = p.dims[0] - i0;
    const int cnt = left < VEC ? (int)left : VEC;
    int64_t m = (p.n == p.dims[0]) ? i0     // nothing was merged into dim 0: the position IS n
              : ((uint64_t)i0 <= 0xffffffffull && (uint64_t)p.n <= 0xffffffffull) ? (int64_t)((uint32_t)i0 % (uint32_t)p.n) : i0 % p.n;
    Pack<T> r;
    // floating point: while the values are exactly representable consecutive integers, one conversion and
    // VEC-1 exact additions replace VEC (quarter-rate) int->float conversions
    constexpr int64_t EXACT = tt<T>::is_int ? 0 : (sizeof(T) == 4 ? (1ll << 24) : (1ll << 53));
    if (!tt<T>::is_int && m + VEC <= p.n && m + VEC <= EXACT) {
      const T b0 = (m <= 0x7fffffff) ? (T)(int)m : (T)m;
#pragma unroll
      for (int k = 0; k < VEC; k++) r.e[k] = b0 + (T)k;
    } else {
#pragma unroll
      for (int k = 0; k < VEC; k++) { r.e[k] = (T)m; if (++m == p.n) m = 0; }
    }
    if (p.vec && cnt == VEC) *reinterpret_cast<uint4 *>(base + off) = r.q;
    else {
#pragma unroll
      for (int k = 0; k < VEC; k++) if (k < cnt) base[off + k * p.st[0]] = r.e[k];
    }
   }
  }
}

template <class T>
static int ax_go(const AxPlan &p, cudaStream_t s, const Err &E) {
  int64_t g = (p.n_units + 256 * 4 - 1) / (256 * 4);
  if (g < 1) g = 1;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (g > cap) g = cap;
  axisvals_kernel<T><<<(int)g, 256, 0, s>>>(p);
  note_launch("axisvalues");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

int launch_axisvalues(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 2) return E.fail(PDLB200_EINVAL, "axisvalues: expected 2 parameters");
  if (t->ndims + 1 > PDLB200_MAXDIMS) return E.fail(PDLB200_EUNSUPPORTED, "axisvalues: too many broadcast dims");
  const int64_t n = t->ind[0];
  if (n < 0) return E.fail(PDLB200_EINVAL, "axisvalues: n = %lld", (long long)n);
  // one-parameter walk over [n, broadcast dims...] of the output
  pdlb200_trans w = *t;
  w.npdls = 1; w.ndims = t->ndims + 1;
  w.dims[0] = n; w.incs[0] = t->rinc[1];
  for (int d = 0; d < t->ndims; d++) { w.dims[d + 1] = t->dims[d]; w.incs[d + 1] = t->incs[d * t->npdls + 1]; }
  w.pdls[0] = t->pdls[1];
  Collapsed c;
  collapse_dims(&w, &c);
  if (c.total == 0) return PDLB200_OK;
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "axisvalues: %d non-mergeable dims exceed the device walker's %d", c.nd, MAXD);
  if (!w.pdls[0].data) return E.fail(PDLB200_EINVAL, "axisvalues: output got NULL data");
  const size_t sz = pdlb200_type_size(t->datatype);
  if (!sz) return E.fail(PDLB200_EUNSUPPORTED, "axisvalues: type %d is not on the device path", t->datatype);
  AxPlan p{};
  p.a = (char *)w.pdls[0].data + w.pdls[0].offs * (int64_t)sz;
  p.nd = c.nd; p.n = n;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.st[d] = c.st[0][d]; }
  const int VEC = (int)(16 / sz);
  // dim 0 of the collapsed walk starts with the named dim unless n == 1 was dropped (then every value is 0 = p mod 1)
  bool ok = (p.st[0] == 1) && (((uintptr_t)p.a) % 16 == 0);
  for (int d = 1; d < c.nd && ok; d++) if ((p.st[d] * (int64_t)sz) % 16 != 0) ok = false;
  p.vec = ok;
  p.vpr = (p.dims[0] + VEC - 1) / VEC;
  p.n_units = p.vpr * (c.total / p.dims[0]);
  cudaStream_t s = (cudaStream_t)t->stream;
  switch (t->datatype) {
    case PDLB200_SB: return ax_go<int8_t>(p, s, E);   case PDLB200_B:  return ax_go<uint8_t>(p, s, E);
    case PDLB200_S:  return ax_go<int16_t>(p, s, E);  case PDLB200_US: return ax_go<uint16_t>(p, s, E);
    case PDLB200_L:  return ax_go<int32_t>(p, s, E);  case PDLB200_UL: return ax_go<uint32_t>(p, s, E);
    case PDLB200_IND: case PDLB200_LL: return ax_go<int64_t>(p, s, E);
    case PDLB200_ULL: return ax_go<uint64_t>(p, s, E);
    case PDLB200_F:  return ax_go<float>(p, s, E);    case PDLB200_D:  return ax_go<double>(p, s, E);
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "axisvalues: type %d is not on the device path", t->datatype);
}
}  // namespace pdlb200
