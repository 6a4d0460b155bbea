// complex.cu — plus minus mult divide on complex float / complex double (lib/PDL/Ops.pd:104-153,288-291 for the
// C and G generic types).  The reference's arithmetic is whatever gcc emits for C99 `_Complex` operands on
// x86-64 (-O2, no FMA):
//   a + b, a - b : component-wise;
//   a * b        : inline  x = ac - bd, y = ad + bc;  only if BOTH come out NaN, libgcc's __mulsc3 / __muldc3
//                  recover infinities (C99 Annex G);
//   a / b        : always a libgcc call.  __divsc3 (float) works in DOUBLE: x = (ac + bd) / (cc + dd),
//                  y = (bc - ad) / (cc + dd), rounded to float once.  __divdc3 (double) is the scaled Smith
//                  algorithm of GCC >= 11 (Baudin & Smith): scale by 1/2 near overflow, by 2^52 near underflow,
//                  ratio = c/d or d/c, with the small-ratio variant; then the same Annex G recovery.
// Every step below is one IEEE operation in the same order (read off the compiled code of GCC 13.3 / its libgcc),
// so finite results are bit-exact; NaN payloads are not tracked.  BAD: an element is BAD when BOTH parts equal the
// badvalue's parts (complex ==), Ops.pd:144-148.  One element per thread per trip (a complex double is a 128-bit
// load); plus / minus without BAD values run as the REAL op over one more leading dim of size 2.
#include <cstring>
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {

template <class R> struct Cx { R re, im; };

template <class R> __device__ __forceinline__ R cx_inf() { if constexpr (sizeof(R) == 4) return __uint_as_float(0x7f800000u); else return __longlong_as_double(0x7ff0000000000000ll); }
template <class R> __device__ __forceinline__ R cx_copysign(R m, R s) { if constexpr (sizeof(R) == 4) return copysignf(m, s); else return copysign(m, s); }
template <class R> __device__ __forceinline__ R cx_box(bool inf, R s) { return cx_copysign(inf ? R(1) : R(0), s); }

// libgcc2.c __mulsc3 / __muldc3
template <class R> __device__ __forceinline__ Cx<R> cx_mul(Cx<R> p, Cx<R> q) {
  R a = p.re, b = p.im, c = q.re, d = q.im;
  const R ac = a * c, bd = b * d, ad = a * d, bc = b * c;
  R x = ac - bd, y = ad + bc;
  if (x != x && y != y) {
    bool recalc = false;
    if (isinf(a) || isinf(b)) {
      a = cx_box<R>(isinf(a), a); b = cx_box<R>(isinf(b), b);
      if (c != c) c = cx_copysign<R>(R(0), c);
      if (d != d) d = cx_copysign<R>(R(0), d);
      recalc = true;
    }
    if (isinf(c) || isinf(d)) {
      c = cx_box<R>(isinf(c), c); d = cx_box<R>(isinf(d), d);
      if (a != a) a = cx_copysign<R>(R(0), a);
      if (b != b) b = cx_copysign<R>(R(0), b);
      recalc = true;
    }
    if (!recalc && (isinf(ac) || isinf(bd) || isinf(ad) || isinf(bc))) {
      if (a != a) a = cx_copysign<R>(R(0), a);
      if (b != b) b = cx_copysign<R>(R(0), b);
      if (c != c) c = cx_copysign<R>(R(0), c);
      if (d != d) d = cx_copysign<R>(R(0), d);
      recalc = true;
    }
    if (recalc) { x = cx_inf<R>() * (a * c - b * d); y = cx_inf<R>() * (a * d + b * c); }
  }
  return Cx<R>{x, y};
}

// the Annex G recovery shared by __divsc3 and __divdc3 (in the operand type R)
template <class R> __device__ __forceinline__ void cx_div_recover(R a, R b, R c, R d, R &x, R &y) {
  if (!(x != x && y != y)) return;
  if (c == R(0) && d == R(0) && (a == a || b == b)) {
    x = cx_copysign<R>(cx_inf<R>(), c) * a; y = cx_copysign<R>(cx_inf<R>(), c) * b;
  } else if ((isinf(a) || isinf(b)) && isfinite(c) && isfinite(d)) {
    a = cx_box<R>(isinf(a), a); b = cx_box<R>(isinf(b), b);
    x = cx_inf<R>() * (a * c + b * d); y = cx_inf<R>() * (b * c - a * d);
  } else if ((isinf(c) || isinf(d)) && isfinite(a) && isfinite(b)) {
    c = cx_box<R>(isinf(c), c); d = cx_box<R>(isinf(d), d);
    x = R(0) * (a * c + b * d); y = R(0) * (b * c - a * d);
  }
}
__device__ __forceinline__ Cx<float> cx_div(Cx<float> p, Cx<float> q) {
  const double a = p.re, b = p.im, c = q.re, d = q.im;
  const double denom = c * c + d * d;
  float x = (float)((a * c + b * d) / denom), y = (float)((b * c - a * d) / denom);
  cx_div_recover<float>(p.re, p.im, q.re, q.im, x, y);
  return Cx<float>{x, y};
}
__device__ __forceinline__ Cx<double> cx_div(Cx<double> p, Cx<double> q) {
  const double RBIG = __longlong_as_double(0x7fdfffffffffffffll);     // DBL_MAX / 2
  const double RMIN = __longlong_as_double(0x0010000000000000ll);     // DBL_MIN
  const double RMIN2 = __longlong_as_double(0x3cb0000000000000ll);    // DBL_EPSILON
  const double RMINSCAL = __longlong_as_double(0x4330000000000000ll); // 1 / DBL_EPSILON
  const double RMAX2 = __longlong_as_double(0x7c9fffffffffffffll);    // RBIG * RMIN2
  double a = p.re, b = p.im, c = q.re, d = q.im, x, y, ratio, denom;
  if (fabs(c) < fabs(d)) {
    if (fabs(d) >= RBIG) { a = a / 2; b = b / 2; c = c / 2; d = d / 2; }
    if (fabs(d) < RMIN2) { a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL; }
    else if (((fabs(a) < RMIN) && (fabs(b) < RMAX2) && (fabs(d) < RMAX2)) ||
             ((fabs(b) < RMIN) && (fabs(a) < RMAX2) && (fabs(d) < RMAX2))) {
      a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL;
    }
    ratio = c / d;
    denom = (c * ratio) + d;
    if (fabs(ratio) > RMIN) { x = ((a * ratio) + b) / denom; y = ((b * ratio) - a) / denom; }
    else { x = ((c * (a / d)) + b) / denom; y = ((c * (b / d)) - a) / denom; }
  } else {
    if (fabs(c) >= RBIG) { a = a / 2; b = b / 2; c = c / 2; d = d / 2; }
    if (fabs(c) < RMIN2) { a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL; }
    else if (((fabs(a) < RMIN) && (fabs(b) < RMAX2) && (fabs(c) < RMAX2)) ||
             ((fabs(b) < RMIN) && (fabs(a) < RMAX2) && (fabs(c) < RMAX2))) {
      a = a * RMINSCAL; b = b * RMINSCAL; c = c * RMINSCAL; d = d * RMINSCAL;
    }
    ratio = d / c;
    denom = (d * ratio) + c;
    if (fabs(ratio) > RMIN) { x = ((b * ratio) + a) / denom; y = (b - (a * ratio)) / denom; }
    else { x = ((d * (b / c)) + a) / denom; y = (b - (d * (a / c))) / denom; }
  }
  cx_div_recover<double>(a, b, c, d, x, y);
  return Cx<double>{x, y};
}

template <class R, int OP> __device__ __forceinline__ Cx<R> cx_apply(Cx<R> a, Cx<R> b) {
  if constexpr (OP == PDLB200_OP_PLUS) return Cx<R>{a.re + b.re, a.im + b.im};
  else if constexpr (OP == PDLB200_OP_MINUS) return Cx<R>{a.re - b.re, a.im - b.im};
  else if constexpr (OP == PDLB200_OP_MULT) return cx_mul<R>(a, b);
  else return cx_div(a, b);
}

// one complex element per thread per trip; the EwPlan walker (strides in complex elements)
template <class R, int OP, bool BAD>
__global__ void __launch_bounds__(EW_THREADS) cx_kernel(const __grid_constant__ EwPlan p, Cx<R> abad, Cx<R> bbad, Cx<R> cbad) {
  const int64_t n0 = p.dims[0];
  int64_t total = 1;
  for (int d = 0; d < p.nd; d++) total *= p.dims[d];
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e / n0, i0 = e - r * n0;
    int64_t oa = i0 * p.st[0][0], ob = i0 * p.st[1][0], oc = i0 * p.st[2][0];
    for (int d = 1; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : r / p.dims[d];
      const int64_t i = r - q * p.dims[d];
      oa += i * p.st[0][d]; ob += i * p.st[1][d]; oc += i * p.st[2][d];
      r = q;
    }
    const Cx<R> a = reinterpret_cast<const Cx<R> *>(p.ptr[0])[oa], b = reinterpret_cast<const Cx<R> *>(p.ptr[1])[ob];
    Cx<R> c;
    bool bad = false;
    if constexpr (BAD) {
      // PDL_ISBAD2 on a complex type: `==` on both parts, or (NaN badvalue) isnan of either part (Types.pm:209-232)
      const bool ia = p.badnan[0] ? (a.re != a.re || a.im != a.im) : (a.re == abad.re && a.im == abad.im);
      const bool ib = p.badnan[1] ? (b.re != b.re || b.im != b.im) : (b.re == bbad.re && b.im == bbad.im);
      bad = (p.badchk[0] && ia) || (p.badchk[1] && ib);
    }
    if (bad) c = cbad; else c = cx_apply<R, OP>(a, b);
    reinterpret_cast<Cx<R> *>(p.ptr[2])[oc] = c;
  }
}

template <class R> static Cx<R> cx_badval(const pdlb200_par &par) {
  Cx<R> v;
  if constexpr (sizeof(R) == 4) memcpy(&v, &par.badval, 8);               // both parts
  else { memcpy(&v.re, &par.badval, 8); v.im = v.re; }                     // real part; imaginary part equal
  return v;
}

template <class R> static int cx_go(const pdlb200_trans *t, const Err &E) {
  EwPlan p;
  int rc = ew_build_plan(t, 2, sizeof(Cx<R>), sizeof(Cx<R>), true, &p, E, sizeof(Cx<R>));
  if (rc) return rc;
  if (p.n_units == 0) return PDLB200_OK;
  int64_t total = 1;
  for (int d = 0; d < p.nd; d++) total *= p.dims[d];
  int64_t g = (total + EW_THREADS - 1) / EW_THREADS;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (g > cap) g = cap;
  cudaStream_t s = (cudaStream_t)t->stream;
  const Cx<R> ab = cx_badval<R>(t->pdls[0]), bb = cx_badval<R>(t->pdls[1]), cb = cx_badval<R>(t->pdls[2]);
#define CX_GO(OP) do { if (t->bvalflag) cx_kernel<R, OP, true><<<(int)g, EW_THREADS, 0, s>>>(p, ab, bb, cb); \
                       else cx_kernel<R, OP, false><<<(int)g, EW_THREADS, 0, s>>>(p, ab, bb, cb); } while (0)
  switch (t->op) {
    case PDLB200_OP_PLUS: CX_GO(PDLB200_OP_PLUS); break;
    case PDLB200_OP_MINUS: CX_GO(PDLB200_OP_MINUS); break;
    case PDLB200_OP_MULT: CX_GO(PDLB200_OP_MULT); break;
    case PDLB200_OP_DIVIDE: CX_GO(PDLB200_OP_DIVIDE); break;
    default: return E.fail(PDLB200_EUNSUPPORTED, "%s: not defined for complex types on the device", pdlb200_op_name(t->op));
  }
#undef CX_GO
  note_launch(sizeof(R) == 4 ? "ew_complex_float" : "ew_complex_double");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

int launch_complex(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 3) return E.fail(PDLB200_EINVAL, "%s: expected 3 parameters", pdlb200_op_name(t->op));
  for (int k = 0; k < 3; k++)
    if (t->pdls[k].type != t->datatype) return E.fail(PDLB200_EINVAL, "%s: complex parameters must all have the operation's type", pdlb200_op_name(t->op));
  // plus / minus without BAD values: the real op over one more leading dim of size 2 (re, im), at the real kernels' speed
  if (!t->bvalflag && (t->op == PDLB200_OP_PLUS || t->op == PDLB200_OP_MINUS) && t->ndims < PDLB200_MAXDIMS) {
    pdlb200_trans w = *t;
    w.datatype = t->datatype == PDLB200_CF ? PDLB200_F : PDLB200_D;
    w.ndims = t->ndims + 1;
    w.dims[0] = 2;
    for (int k = 0; k < 3; k++) { w.incs[k] = 1; w.pdls[k].type = w.datatype; w.pdls[k].offs = t->pdls[k].offs * 2; w.pdls[k].flags = 0; }
    for (int d = 0; d < t->ndims; d++) {
      w.dims[d + 1] = t->dims[d];
      for (int k = 0; k < 3; k++) w.incs[(d + 1) * 3 + k] = t->incs[d * 3 + k] * 2;
    }
    return launch_elementwise(&w, E);
  }
  return t->datatype == PDLB200_CF ? cx_go<float>(t, E) : cx_go<double>(t, E);
}

}  // namespace pdlb200
