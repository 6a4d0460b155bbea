// ew_ipow.cu — ipow, lib/PDL/Ops.pd:443-476: a(); longlong b(); [o]ans().
// The exponent parameter is ALWAYS longlong whatever the generic type, so this op does not fit the
// same-type walker of elementwise.cuh; it is compute-bound (a multiply chain per element), so a plain
// one-element-per-thread kernel over the collapsed broadcast dims is enough.  The loop is the
// reference's exponentiation by squaring, multiplication for multiplication (-fmad=false), so float
// and double results are bit-exact.  GenericTypes P Q + floats -> device types ULL, LL, F, D.
#include <cstring>
#include "common.cuh"

namespace pdlb200 {

struct IpPlan {
  const char *a, *b; char *c;
  int64_t dims[MAXD];
  int64_t sa[MAXD], sb[MAXD], sc[MAXD];
  int64_t total;
  int nd;
};

template <class T>
__global__ void __launch_bounds__(256) ipow_kernel(const __grid_constant__ IpPlan p) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t oa = 0, ob = 0, oc = 0, r = e;
    for (int d = 0; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : r / p.dims[d];
      const int64_t i = r - q * p.dims[d];
      oa += i * p.sa[d]; ob += i * p.sb[d]; oc += i * p.sc[d];
      r = q;
    }
    const T a = reinterpret_cast<const T *>(p.a)[oa];
    long long n = reinterpret_cast<const long long *>(p.b)[ob];
    T *out = reinterpret_cast<T *>(p.c) + oc;
    if (n == 0) { *out = T(1); continue; }
    T y = T(1), x = a;
    if (n < 0) {
      if constexpr (tt<T>::is_int) x = (x == T(0)) ? T(0) : (T)(T(1) / x);   // 1/0 kills the reference (SIGFPE)
      else x = T(1) / x;
      n = (long long)(0ull - (unsigned long long)n);
    }
    while (n > 1) {
      if (n % 2) {
        if constexpr (tt<T>::is_int) y = (T)((unsigned long long)y * (unsigned long long)x); else y = y * x;
        n -= 1;
      }
      if constexpr (tt<T>::is_int) x = (T)((unsigned long long)x * (unsigned long long)x); else x = x * x;
      n /= 2;
    }
    if constexpr (tt<T>::is_int) *out = (T)((unsigned long long)x * (unsigned long long)y); else *out = x * y;
  }
}

template <class T>
static int ipow_go(const pdlb200_trans *t, const IpPlan &p, const Err &E) {
  int64_t g = (p.total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (g > cap) g = cap;
  ipow_kernel<T><<<(int)g, 256, 0, (cudaStream_t)t->stream>>>(p);
  note_launch("ew_ipow");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

int launch_ipow(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 3) return E.fail(PDLB200_EINVAL, "ipow: expected 3 parameters, got %d", t->npdls);
  if (t->pdls[1].type != PDLB200_LL && t->pdls[1].type != PDLB200_IND)
    return E.fail(PDLB200_EINVAL, "ipow: parameter b must be longlong, got type %d", t->pdls[1].type);
  Collapsed c;
  collapse_dims(t, &c);
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "ipow: %d non-mergeable broadcast dims exceed %d", c.nd, MAXD);
  IpPlan p;
  memset(&p, 0, sizeof p);
  p.nd = c.nd; p.total = c.total;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sa[d] = c.st[0][d]; p.sb[d] = c.st[1][d]; p.sc[d] = c.st[2][d]; }
  if (p.total == 0) return PDLB200_OK;
  const size_t sz = pdlb200_type_size(t->datatype);
  for (int k = 0; k < 3; k++) if (!t->pdls[k].data) return E.fail(PDLB200_EINVAL, "ipow: parameter %d got NULL data", k);
  p.a = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)sz;
  p.b = (const char *)t->pdls[1].data + t->pdls[1].offs * 8;
  p.c = (char *)t->pdls[2].data + t->pdls[2].offs * (int64_t)sz;
  switch (t->datatype) {
    case PDLB200_ULL: return ipow_go<uint64_t>(t, p, E);
    case PDLB200_IND: case PDLB200_LL: return ipow_go<int64_t>(t, p, E);
    case PDLB200_F: return ipow_go<float>(t, p, E);
    case PDLB200_D: return ipow_go<double>(t, p, E);
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "ipow: type %d is not in its GenericTypes on the device (ulonglong longlong float double)", t->datatype);
}

}  // namespace pdlb200
