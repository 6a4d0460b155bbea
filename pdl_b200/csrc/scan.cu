// scan.cu — cumusumover / cumuprodover (+d variants), lib/PDL/Ufunc.pd:120-141: a(n); [o]b(n).
// Two kernels:
//  * scan_rows_kernel: one thread per row walks n sequentially — the reference's own order, so float
//    results are bit-exact; coalesced when a broadcast dim is the unit-stride one (many short rows).
//  * scan_warp_kernel: one warp per row for long rows: 32 consecutive elements per step (coalesced),
//    Kogge-Stone scan by shuffles, carry to the next step.  Integer results are bit-exact (wrap-around
//    add/multiply is associative); float results differ from the sequential order only in rounding.
// A single very long row still runs on one warp (a multi-CTA look-back scan is future work).
#include <cstring>
#include "common.cuh"
namespace pdlb200 {

struct ScPlan {
  const char *a; char *b;
  int64_t n, inc_a, inc_b, nrows;
  int64_t dims[MAXD], sa[MAXD], sb[MAXD];
  uint64_t abad, bbad;
  int nd, abadnan, badmode;
};

template <class T, class O, bool PROD>
__global__ void __launch_bounds__(256) scan_rows_kernel(const __grid_constant__ ScPlan p) {
  const T abad = from_bits<T>(p.abad);
  const O bbad = from_bits<O>(p.bbad);
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.nrows; row += (int64_t)gridDim.x * blockDim.x) {
    int64_t oa = 0, ob = 0, r = row;
    for (int d = 0; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : r / p.dims[d];
      const int64_t i = r - q * p.dims[d];
      oa += i * p.sa[d]; ob += i * p.sb[d]; r = q;
    }
    const T *pa = reinterpret_cast<const T *>(p.a) + oa;
    O *pb = reinterpret_cast<O *>(p.b) + ob;
    O tmp = PROD ? O(1) : O(0);
    for (int64_t n = 0; n < p.n; n++) {
      const T v = pa[n * p.inc_a];
      if (p.badmode && is_bad(v, abad, p.abadnan != 0)) { pb[n * p.inc_b] = bbad; continue; }
      if constexpr (tt<O>::is_int) {
        using U = typename tt<O>::wide_u;
        tmp = PROD ? (O)((U)tmp * (U)(O)v) : (O)((U)tmp + (U)(O)v);
      } else {
        tmp = PROD ? x86_nan2(tmp, (O)v, tmp * (O)v) : x86_nan2(tmp, (O)v, tmp + (O)v);
      }
      pb[n * p.inc_b] = tmp;
    }
  }
}

template <class O, bool PROD> __device__ __forceinline__ O scan_op(O a, O b) {
  if constexpr (tt<O>::is_int) {
    using U = typename tt<O>::wide_u;
    return PROD ? (O)((U)a * (U)b) : (O)((U)a + (U)b);
  } else return PROD ? a * b : a + b;
}
template <class O> __device__ __forceinline__ O shfl_up_t(O v, int d) {
  if constexpr (sizeof(O) == 8) {
    unsigned long long u; memcpy(&u, &v, 8);
    u = __shfl_up_sync(0xffffffffu, u, d);
    memcpy(&v, &u, 8); return v;
  } else {
    unsigned u; memcpy(&u, &v, 4);
    u = __shfl_up_sync(0xffffffffu, u, d);
    memcpy(&v, &u, 4); return v;
  }
}
template <class O> __device__ __forceinline__ O shfl_idx_t(O v, int src) {
  if constexpr (sizeof(O) == 8) {
    unsigned long long u; memcpy(&u, &v, 8);
    u = __shfl_sync(0xffffffffu, u, src);
    memcpy(&v, &u, 8); return v;
  } else {
    unsigned u; memcpy(&u, &v, 4);
    u = __shfl_sync(0xffffffffu, u, src);
    memcpy(&v, &u, 4); return v;
  }
}

template <class T, class O, bool PROD>
__global__ void __launch_bounds__(256) scan_warp_kernel(const __grid_constant__ ScPlan p) {
  const T abad = from_bits<T>(p.abad);
  const O bbad = from_bits<O>(p.bbad);
  const O ident = PROD ? O(1) : O(0);
  const int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t row_step = (int64_t)gridDim.x * 8;
  for (; row < p.nrows; row += row_step) {
    int64_t oa = 0, ob = 0, r = row;
    for (int d = 0; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : r / p.dims[d];
      const int64_t i = r - q * p.dims[d];
      oa += i * p.sa[d]; ob += i * p.sb[d]; r = q;
    }
    const T *pa = reinterpret_cast<const T *>(p.a) + oa;
    O *pb = reinterpret_cast<O *>(p.b) + ob;
    O carry = ident;
    for (int64_t n0 = 0; n0 < p.n; n0 += 32) {
      const int64_t n = n0 + lane;
      const bool in = n < p.n;
      const T v = in ? pa[n * p.inc_a] : T(0);
      const bool bad = in && p.badmode && is_bad(v, abad, p.abadnan != 0);
      O x = (in && !bad) ? (O)v : ident;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const O y = shfl_up_t(x, d);
        if (lane >= d) x = scan_op<O, PROD>(y, x);
      }
      const O res = scan_op<O, PROD>(carry, x);
      if (in) pb[n * p.inc_b] = bad ? bbad : res;
      carry = shfl_idx_t(res, 31);
    }
  }
}

template <class T, class O, bool PROD>
static int scan_go(const ScPlan &p, cudaStream_t s, const char *name, const Err &E) {
  const int64_t cap = (int64_t)sm_count() * 8;
  // long rows that are not laid out column-wise: one warp per row
  const bool column = (p.nd >= 1) && (p.sa[0] == 1 || p.sa[0] == -1) && p.inc_a != 1 && p.dims[0] >= 32;
  if (p.n >= 128 && !column) {
    int64_t g = (p.nrows + 7) / 8;
    if (g > cap) g = cap;
    scan_warp_kernel<T, O, PROD><<<(int)g, 256, 0, s>>>(p);
  } else {
    int64_t g = (p.nrows + 255) / 256;
    if (g > cap) g = cap;
    scan_rows_kernel<T, O, PROD><<<(int)g, 256, 0, s>>>(p);
  }
  note_launch(name);
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

template <bool PROD, bool DBL>
static int scan_types(const pdlb200_trans *t, const ScPlan &p, const char *name, const Err &E) {
  cudaStream_t s = (cudaStream_t)t->stream;
#define SC(T) if constexpr (DBL) return scan_go<T, double, PROD>(p, s, name, E); else return scan_go<T, typename tt<T>::plus, PROD>(p, s, name, E);
  switch (t->datatype) {
    case PDLB200_SB: SC(int8_t) case PDLB200_B: SC(uint8_t) case PDLB200_S: SC(int16_t) case PDLB200_US: SC(uint16_t)
    case PDLB200_L: SC(int32_t) case PDLB200_UL: SC(uint32_t) case PDLB200_IND: case PDLB200_LL: SC(int64_t)
    case PDLB200_ULL: SC(uint64_t) case PDLB200_F: SC(float) case PDLB200_D: SC(double)
    default: break;
  }
#undef SC
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}

int launch_scan(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 2) return E.fail(PDLB200_EINVAL, "%s: expected 2 parameters", pdlb200_op_name(t->op));
  Collapsed c;
  collapse_dims(t, &c);
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "%s: too many non-mergeable broadcast dims", pdlb200_op_name(t->op));
  ScPlan p;
  memset(&p, 0, sizeof p);
  p.n = t->ind[0]; p.inc_a = t->rinc[0]; p.inc_b = t->rinc[1]; p.nrows = c.total; p.nd = c.nd;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sa[d] = c.st[0][d]; p.sb[d] = c.st[1][d]; }
  if (p.nrows == 0 || p.n == 0) return PDLB200_OK;
  const size_t isz = pdlb200_type_size(t->pdls[0].type), osz = pdlb200_type_size(t->pdls[1].type);
  if (!t->pdls[0].data || !t->pdls[1].data) return E.fail(PDLB200_EINVAL, "%s: parameter got NULL data", pdlb200_op_name(t->op));
  p.a = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)isz;
  p.b = (char *)t->pdls[1].data + t->pdls[1].offs * (int64_t)osz;
  p.abad = t->pdls[0].badval; p.bbad = t->pdls[1].badval;
  p.abadnan = (t->pdls[0].flags & PDLB200_PAR_BADNAN) != 0;
  p.badmode = t->bvalflag != 0;
  switch (t->op) {
    case PDLB200_OP_CUMUSUMOVER:   return scan_types<false, false>(t, p, "scan_cumusumover", E);
    case PDLB200_OP_CUMUPRODOVER:  return scan_types<true,  false>(t, p, "scan_cumuprodover", E);
    case PDLB200_OP_DCUMUSUMOVER:  return scan_types<false, true>(t, p, "scan_dcumusumover", E);
    case PDLB200_OP_DCUMUPRODOVER: return scan_types<true,  true>(t, p, "scan_dcumuprodover", E);
    default: break;
  }
  return E.fail(PDLB200_EINVAL, "%s is not a scan", pdlb200_op_name(t->op));
}
}  // namespace pdlb200
