// scan.cu — cumusumover / cumuprodover (+d variants), lib/PDL/Ufunc.pd:120-141: a(n); [o]b(n).
// Kernels:
//  * scan_rows_kernel: one thread per row walks n sequentially — the reference's own order, so float
//    results are bit-exact; coalesced when a broadcast dim is the unit-stride one (many short rows).
//  * scan_chunk_kernel<APPLY>: one warp per (row, chunk).  Unit-stride 16-byte-aligned chunks move as
//    128-bit vectors, 4 in flight per lane: each lane scans its vector, the lane totals are scanned
//    across the warp by shuffles (Kogge-Stone), the carry moves on to the next step; other layouts
//    take 32 consecutive elements per step.  Integer results are bit-exact (wrap-around add/multiply
//    is associative); float results differ from the sequential order only in rounding.
//    - long rows, many of them: one chunk = the whole row, no carry-in.
//    - few very long rows (a 1-D cumusumover is the common case): scan_onepass_kernel (scan_onepass.cuh) — one
//      pass, decoupled look-back, bulk-async tiles — when T and O have the same size and the rows are 16-byte
//      aligned; otherwise three passes over chunks of a multiple of SC_CHUNK elements sized for one resident
//      wave of warps: (1) chunk totals, (2) scan_chunk_prefix_kernel: exclusive scan of the totals per row,
//      (3) the scan again with the chunk's carry-in.  3 passes of traffic instead of the ideal 2.
#include <cstdio>
#include <cstring>
#include "common.cuh"
namespace pdlb200 {

constexpr int64_t SC_CHUNK = 4096;

struct ScPlan {
  const char *a; char *b;
  char *carry;                  // chunked mode: one O per (row, chunk)
  int64_t nchunks, chunk;
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
template <class O> __device__ __forceinline__ O shfl_xor_t(O v, int m) {
  if constexpr (sizeof(O) == 8) {
    unsigned long long u; memcpy(&u, &v, 8);
    u = __shfl_xor_sync(0xffffffffu, u, m);
    memcpy(&v, &u, 8); return v;
  } else {
    unsigned u; memcpy(&u, &v, 4);
    u = __shfl_xor_sync(0xffffffffu, u, m);
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

// chunked mode, pass 1 and 3.  APPLY=false: write the chunk's total (good elements only) to carry[row][chunk];
// APPLY=true: scan the chunk starting from carry[row][chunk] (already turned into an exclusive prefix by pass 2).
template <class T, class O, bool PROD, bool APPLY>
__global__ void __launch_bounds__(256) scan_chunk_kernel(const __grid_constant__ ScPlan p) {
  const T abad = from_bits<T>(p.abad);
  const O bbad = from_bits<O>(p.bbad);
  const O ident = PROD ? O(1) : O(0);
  const int lane = threadIdx.x & 31;
  const int64_t nwork = p.nrows * p.nchunks;
  int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (; w < nwork; w += (int64_t)gridDim.x * 8) {
    const int64_t row = w / p.nchunks, chunk = w - row * p.nchunks;
    int64_t oa = 0, ob = 0, r = row;
    for (int d = 0; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : r / p.dims[d];
      const int64_t i = r - q * p.dims[d];
      oa += i * p.sa[d]; ob += i * p.sb[d]; r = q;
    }
    const T *pa = reinterpret_cast<const T *>(p.a) + oa;
    O *pb = reinterpret_cast<O *>(p.b) + ob;
    O *cw = reinterpret_cast<O *>(p.carry) + w;     // not dereferenced when p.carry is NULL (whole rows, no carry-in)
    const int64_t lo = chunk * p.chunk, hi = (lo + p.chunk < p.n) ? lo + p.chunk : p.n;
    O carry = (APPLY && p.carry) ? *cw : ident;
    constexpr int VEC = 16 / sizeof(T);
    // vector path: same-size in/out, unit strides, 16-byte aligned chunk starts (p.chunk is a multiple of SC_CHUNK, hence of every VEC)
    const bool vec = sizeof(T) == sizeof(O) && p.inc_a == 1 && (!APPLY || p.inc_b == 1) &&
                     (((uintptr_t)(pa + lo)) & 15) == 0 && (!APPLY || (((uintptr_t)(pb + lo)) & 15) == 0);
    if (vec) {
      constexpr int U = 4;                                   // 128-bit loads in flight per lane
      const int64_t nvec = (hi - lo) / VEC;                  // full vectors; the tail goes through the scalar loop
      const uint4 *vp = reinterpret_cast<const uint4 *>(pa + lo);
      O tot = ident;
      for (int64_t v0 = 0; v0 < nvec; v0 += 32 * U) {
        Pack<T> in[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const int64_t j = v0 + u * 32 + lane; if (j < nvec) in[u].q = vp[j]; }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int64_t j = v0 + u * 32 + lane;
          const bool inr = j < nvec;
          O x[VEC]; bool bd[VEC];
          O run = ident;
#pragma unroll
          for (int k = 0; k < VEC; k++) {
            const T v = inr ? in[u].e[k] : T(0);
            bd[k] = inr && p.badmode && is_bad(v, abad, p.abadnan != 0);
            if (inr && !bd[k]) run = scan_op<O, PROD>(run, (O)v);
            x[k] = run;                                      // inclusive scan inside the lane's vector
          }
          if (!APPLY) { tot = scan_op<O, PROD>(tot, run); continue; }
          O pre = run;                                       // exclusive scan of the lane totals across the warp
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const O y = shfl_up_t(pre, d);
            if (lane >= d) pre = scan_op<O, PROD>(y, pre);
          }
          const O warp_total = shfl_idx_t(pre, 31);
          O excl = shfl_up_t(pre, 1);
          if (lane == 0) excl = ident;
          const O base = scan_op<O, PROD>(carry, excl);
          if (inr) {
            Pack<O> out;
#pragma unroll
            for (int k = 0; k < VEC; k++) out.e[k] = bd[k] ? bbad : scan_op<O, PROD>(base, x[k]);
            reinterpret_cast<uint4 *>(pb + lo)[j] = out.q;
          }
          carry = scan_op<O, PROD>(carry, warp_total);
        }
      }
      if (!APPLY) {
        for (int64_t n = lo + nvec * VEC + lane; n < hi; n += 32) {
          const T v = pa[n];
          if (!(p.badmode && is_bad(v, abad, p.abadnan != 0))) tot = scan_op<O, PROD>(tot, (O)v);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) tot = scan_op<O, PROD>(tot, shfl_xor_t(tot, d));
        if (lane == 0) *cw = tot;
      } else {
        for (int64_t n0 = lo + nvec * VEC; n0 < hi; n0 += 32) {
          const int64_t n = n0 + lane;
          const bool in1 = n < hi;
          const T v = in1 ? pa[n] : T(0);
          const bool bad = in1 && p.badmode && is_bad(v, abad, p.abadnan != 0);
          O x = (in1 && !bad) ? (O)v : ident;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) { const O y = shfl_up_t(x, d); if (lane >= d) x = scan_op<O, PROD>(y, x); }
          const O res = scan_op<O, PROD>(carry, x);
          if (in1) pb[n] = bad ? bbad : res;
          carry = shfl_idx_t(res, 31);
        }
      }
      continue;
    }
    if (!APPLY) {
      O tot = ident;
#pragma unroll 4
      for (int64_t n = lo + lane; n < hi; n += 32) {
        const T v = pa[n * p.inc_a];
        if (!(p.badmode && is_bad(v, abad, p.abadnan != 0))) tot = scan_op<O, PROD>(tot, (O)v);
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) tot = scan_op<O, PROD>(tot, shfl_xor_t(tot, d));
      if (lane == 0) *cw = tot;
    } else {
      for (int64_t n0 = lo; n0 < hi; n0 += 32) {
        const int64_t n = n0 + lane;
        const bool in = n < hi;
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
}

// chunked mode, pass 2: carry[row][c] = op over totals of chunks < c (one warp per row, 4 loads in flight per lane)
template <class O, bool PROD>
__global__ void __launch_bounds__(256) scan_chunk_prefix_kernel(const __grid_constant__ ScPlan p) {
  const O ident = PROD ? O(1) : O(0);
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < p.nrows; row += (int64_t)gridDim.x * 8) {
    O *c = reinterpret_cast<O *>(p.carry) + row * p.nchunks;
    O run = ident;
    for (int64_t k0 = 0; k0 < p.nchunks; k0 += 128) {
      O t[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int64_t k = k0 + u * 32 + lane; t[u] = k < p.nchunks ? c[k] : ident; }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int64_t k = k0 + u * 32 + lane;
        O x = t[u];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const O y = shfl_up_t(x, d); if (lane >= d) x = scan_op<O, PROD>(y, x); }
        O excl = shfl_up_t(x, 1);
        if (lane == 0) excl = ident;
        if (k < p.nchunks) c[k] = scan_op<O, PROD>(run, excl);
        run = scan_op<O, PROD>(run, shfl_idx_t(x, 31));
      }
    }
  }
}

}  // namespace pdlb200
#include "scan_onepass.cuh"
namespace pdlb200 {

template <class T, class O, bool PROD>
static int scan_go(const ScPlan &p, cudaStream_t s, const char *name, const Err &E) {
  const int64_t cap = (int64_t)sm_count() * 8;
  {
    int rc = PDLB200_OK;
    if (scan_onepass_try<T, O, PROD>(p, s, name, E, &rc)) return rc;
  }
  // long rows that are not laid out column-wise: one warp per row
  const bool column = (p.nd >= 1) && (p.sa[0] == 1 || p.sa[0] == -1) && p.inc_a != 1 && p.dims[0] >= 32;
  // chunk length: a multiple of SC_CHUNK giving about one resident wave of warps (64 per SM) over all rows
  int64_t want = cap * 8 / (p.nrows > 0 ? p.nrows : 1);
  if (want < 1) want = 1;
  int64_t chunk = (p.n + want - 1) / want;
  chunk = (chunk + SC_CHUNK - 1) / SC_CHUNK * SC_CHUNK;
  if (chunk < SC_CHUNK) chunk = SC_CHUNK;
  const int64_t nchunks = (p.n + chunk - 1) / chunk;
  if (!column && nchunks >= 4 && p.nrows < cap && p.nrows * nchunks <= (1ll << 26)) {
    // few long rows: cut them into chunks so that every SM takes part
    ScPlan q = p;
    q.nchunks = nchunks;
    q.chunk = chunk;
    q.carry = (char *)scratch((size_t)(p.nrows * nchunks) * sizeof(O), s);
    if (!q.carry) return E.fail(PDLB200_ECUDA, "%s: cannot allocate scan scratch", name);
    int64_t g = (p.nrows * nchunks + 7) / 8;
    if (g > cap * 4) g = cap * 4;
    scan_chunk_kernel<T, O, PROD, false><<<(int)g, 256, 0, s>>>(q);
    int64_t g2 = (p.nrows + 7) / 8;
    scan_chunk_prefix_kernel<O, PROD><<<(int)g2, 256, 0, s>>>(q);
    scan_chunk_kernel<T, O, PROD, true><<<(int)g, 256, 0, s>>>(q);
    note_launch(name); note_launch(name);
  } else if (p.n >= 128 && !column) {
    // one warp per whole row: the chunk kernel's apply pass with a single chunk and no carry-in
    // (128-bit loads, 4 in flight per lane, when the row is unit-stride and 16-byte aligned)
    ScPlan q = p;
    q.nchunks = 1; q.chunk = (p.n + SC_CHUNK - 1) / SC_CHUNK * SC_CHUNK; q.carry = nullptr;
    int64_t g = (p.nrows + 7) / 8;
    if (g > cap * 4) g = cap * 4;
    scan_chunk_kernel<T, O, PROD, true><<<(int)g, 256, 0, s>>>(q);
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
