// minmaximum.cu — minmaximum, lib/PDL/Ufunc.pd:563-613: a(n); [o]cmin(); [o]cmax(); indx [o]cmin_ind();
// indx [o]cmax_ind() — the body of `minmax` (Ufunc.pd:738).  One pass over the row
// for both extremes.  Roofline: HBM, n*sizeof(T) read per row.
//
// Reference loop: BAD elements and NaNs are skipped; the first remaining element initialises both
// extremes; later elements replace them on STRICT < / > only.  Order-independent restatement used
// here: min = smallest value, lowest index among equals (and the same for max) — "first wins" in
// index order is exactly what the sequential strict compare produces.  +0/-0 compare equal, so the
// one met first stays, as in the reference.  A row with no usable element writes BAD to all four
// outputs and sets their badflags (Ufunc.pd:578-583): reported through pdlb200_trans.anybad.
//
// Work split as in inner.cu: one warp per (row, chunk), partials through scratch + a warp-per-row
// finishing pass when a row is cut; one thread per row for many short rows.
#include <cstring>
#include "common.cuh"

namespace pdlb200 {

struct MmxPlan {
  const char *a; char *o[4];
  char *part;
  int *flag;
  int64_t dims[MAXD], sa[MAXD], so[4][MAXD];
  int64_t n, inc_a, nrows, nchunks, chunk;
  uint64_t abad, obad[4];
  int nd, badmode, abadnan, want_flag;
};

template <class T> struct MmxAcc {
  T mn, mx; long long imn, imx; int have;
  __device__ void init() { mn = mx = T(0); imn = imx = 0; have = 0; }
  __device__ __forceinline__ void take(T v, long long i) {
    if (!have) { mn = mx = v; imn = imx = i; have = 1; return; }
    if (v < mn) { mn = v; imn = i; }
    if (v > mx) { mx = v; imx = i; }
  }
  // merge another partial whose indices may be lower or higher than ours
  __device__ __forceinline__ void merge(const MmxAcc &o) {
    if (!o.have) return;
    if (!have) { *this = o; return; }
    if (o.mn < mn || (o.mn == mn && o.imn < imn)) { mn = o.mn; imn = o.imn; }
    if (o.mx > mx || (o.mx == mx && o.imx < imx)) { mx = o.mx; imx = o.imx; }
  }
};

// Lane-local running state for the hot loop: 32-bit indices relative to the chunk start, no "have" branch.
// The extremes start from the identities (+inf / type max for min, -inf / type min for max) and are replaced
// on STRICT compares only, so the first occurrence in the lane's (increasing) order is kept.  A usable element
// always fires at least one of the two compares (it cannot equal both identities), so "saw a usable element"
// is imn >= 0 || imx >= 0; if only one side fired, every usable element equals the other side's identity and
// that extreme sits at the first usable element, which is where the side that did fire recorded its first hit.
// About 8 ALU instructions per element — too many for 4 bytes per element at HBM rate, so the vector loop runs it
// only behind mmx_vec_extremes' filter (≈ 2.5 instructions per element on the common path).
template <class T> struct MmxLoc {
  T mn, mx; int imn, imx;
  __device__ __forceinline__ void init() {
    if constexpr (tt<T>::is_int) {
      constexpr T hi = tt<T>::is_uns ? T(~T(0)) : T((typename tt<T>::wide_u(1) << (sizeof(T) * 8 - 1)) - 1);
      constexpr T lo = tt<T>::is_uns ? T(0) : T(-hi - 1);
      mn = hi; mx = lo;
    } else { mn = T(INFINITY); mx = T(-INFINITY); }
    imn = imx = -1;
  }
  __device__ __forceinline__ void step(T v, int i, bool usable) {   // selects, not branches
    const bool lt = usable && v < mn, gt = usable && v > mx;
    mn = lt ? v : mn; imn = lt ? i : imn;
    mx = gt ? v : mx; imx = gt ? i : imx;
  }
  __device__ __forceinline__ MmxAcc<T> finish(long long base) const {
    MmxAcc<T> a;
    a.have = (imn >= 0) || (imx >= 0);
    a.mn = mn; a.mx = mx;
    a.imn = base + (imn >= 0 ? imn : imx);
    a.imx = base + (imx >= 0 ? imx : imn);
    return a;
  }
};

template <class A> __device__ __forceinline__ A mmx_shfl_down(const A &v, int d) {
  A r;
  constexpr int NW = (sizeof(A) + 3) / 4;
  unsigned w[NW], o[NW];
  memcpy(w, &v, sizeof(A));
#pragma unroll
  for (int k = 0; k < NW; k++) o[k] = __shfl_down_sync(0xffffffffu, w[k], d);
  memcpy(&r, o, sizeof(A));
  return r;
}

__device__ __forceinline__ void mmx_offsets(const MmxPlan &p, int64_t row, int64_t &oa, int64_t (&oo)[4]) {
  oa = 0; oo[0] = oo[1] = oo[2] = oo[3] = 0;
  for (int d = 0; d < p.nd; d++) {
    const int64_t q = (d == p.nd - 1) ? 0 : row / p.dims[d];
    const int64_t i = row - q * p.dims[d];
    oa += i * p.sa[d];
#pragma unroll
    for (int k = 0; k < 4; k++) oo[k] += i * p.so[k][d];
    row = q;
  }
}

template <class T>
__device__ __forceinline__ void mmx_write(const MmxPlan &p, const int64_t (&oo)[4], const MmxAcc<T> &acc, bool &flagged) {
  T *cmin = reinterpret_cast<T *>(p.o[0]) + oo[0], *cmax = reinterpret_cast<T *>(p.o[1]) + oo[1];
  long long *imin = reinterpret_cast<long long *>(p.o[2]) + oo[2], *imax = reinterpret_cast<long long *>(p.o[3]) + oo[3];
  if (acc.have) { *cmin = acc.mn; *cmax = acc.mx; *imin = acc.imn; *imax = acc.imx; }
  else {
    *cmin = from_bits<T>(p.obad[0]); *cmax = from_bits<T>(p.obad[1]);
    *imin = (long long)p.obad[2]; *imax = (long long)p.obad[3];
    flagged = true;
  }
}

// CHK = bad mode with an ordinary (non-NaN) badvalue: the only case that needs a compare against it.  NaNs are
// never usable, so a NaN badvalue needs no test of its own, and integers have no NaN.
template <class T, bool CHK>
__device__ __forceinline__ bool mmx_usable(T v, T abad) {
  if constexpr (CHK) { if (v == abad) return false; }
  return !t_isnan(v);
}

template <class T, bool CHK>
__device__ __forceinline__ void mmx_vec_extremes(const Pack<T> &r, T abad, T &vmn, T &vmx) {
  constexpr int VEC = 16 / sizeof(T);
  if constexpr (tt<T>::is_int) {
    constexpr T hi = tt<T>::is_uns ? T(~T(0)) : T((typename tt<T>::wide_u(1) << (sizeof(T) * 8 - 1)) - 1);
    constexpr T lo = tt<T>::is_uns ? T(0) : T(-hi - 1);
    vmn = hi; vmx = lo;
#pragma unroll
    for (int k = 0; k < VEC; k++) {
      const T v = r.e[k];
      const bool bad = CHK && v == abad;
      const T a = bad ? hi : v, b = bad ? lo : v;
      vmn = a < vmn ? a : vmn; vmx = b > vmx ? b : vmx;
    }
  } else {
    T e[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) { e[k] = r.e[k]; if (CHK && e[k] == abad) e[k] = T(NAN); }
    vmn = vmx = e[0];
#pragma unroll
    for (int k = 1; k < VEC; k++) {
      if constexpr (sizeof(T) == 4) { vmn = fminf(vmn, e[k]); vmx = fmaxf(vmx, e[k]); }
      else { vmn = fmin(vmn, e[k]); vmx = fmax(vmx, e[k]); }
    }
  }
}

// "a row had no usable element": at most ONE plain store of 1 per warp and launch — p.flag may be a device word or, in
// the deferred mode, the caller's pinned host int32 (mapped into the device address space), where atomics and a
// store per flagged row would be slow
__device__ __forceinline__ void mmx_publish_flag(const MmxPlan &p, bool flagged) {
  if (flagged && p.flag) {
    const unsigned m = __activemask();                       // the flagged lanes that got here together: one of them stores
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) *reinterpret_cast<volatile int *>(p.flag) = 1;
  }
}

template <class T, bool CHK>
__global__ void __launch_bounds__(256, 4) minmaximum_warp_kernel(const __grid_constant__ MmxPlan p) {
  const T abad = from_bits<T>(p.abad);
  const int lane = threadIdx.x & 31;
  const int64_t nwork = p.nrows * p.nchunks;
  bool flagged = false;
  for (int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); w < nwork; w += (int64_t)gridDim.x * 8) {
    const int64_t row = w / p.nchunks, chunk = w - row * p.nchunks;
    int64_t oa = 0;
    {   // the four output offsets are decoded after the loop (lane 0 only): 8 fewer live registers in it
      int64_t r = row;
      for (int d = 0; d < p.nd; d++) { const int64_t q = (d == p.nd - 1) ? 0 : r / p.dims[d]; oa += (r - q * p.dims[d]) * p.sa[d]; r = q; }
    }
    const T *pa = reinterpret_cast<const T *>(p.a) + oa;
    const int64_t lo = chunk * p.chunk, hi = (lo + p.chunk < p.n) ? lo + p.chunk : p.n;
    MmxLoc<T> loc; loc.init();
    constexpr int U = 4, VEC = 16 / sizeof(T);
    int64_t n = lo + lane;
    if (p.inc_a == 1 && (((uintptr_t)(pa + lo)) & 15) == 0) {
      // lane-local order is increasing in n, so take()'s strict compares keep the first occurrence
      const int64_t nvec = (hi - lo) / VEC;
      const uint4 *qa = reinterpret_cast<const uint4 *>(pa + lo);
      for (int64_t v0 = 0; v0 < nvec; v0 += 32 * U) {
        Pack<T> ra[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const int64_t j = v0 + u * 32 + lane; if (j < nvec) ra[u].q = qa[j]; }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int64_t j = v0 + u * 32 + lane;
          if (j < nvec) {
            // filter: extremes of the whole 16-byte vector first (unusable elements replaced by values that cannot
            // win: NaN for floats — min/max return the other operand —, the identities for integers); the exact
            // per-element strict-compare update runs only when the vector can change an extreme, which after the
            // first few vectors is rare (the number of record-setting elements of a row grows like log n)
            T vmn, vmx;
            mmx_vec_extremes<T, CHK>(ra[u], abad, vmn, vmx);
            if (vmn < loc.mn || vmx > loc.mx) {
#pragma unroll
              for (int k = 0; k < VEC; k++) { const T v = ra[u].e[k]; loc.step(v, (int)(j * VEC + k), mmx_usable<T, CHK>(v, abad)); }
            }
          }
        }
      }
      n = lo + nvec * VEC + lane;
    }
    // strided / unaligned rows and the tail of a vectorised chunk: 4 independent loads in flight per lane
    for (; n + (U - 1) * 32 < hi; n += U * 32) {
      T v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = pa[(n + u * 32) * p.inc_a];
#pragma unroll
      for (int u = 0; u < U; u++) loc.step(v[u], (int)(n + u * 32 - lo), mmx_usable<T, CHK>(v[u], abad));
    }
    for (; n < hi; n += 32) { const T v = pa[n * p.inc_a]; loc.step(v, (int)(n - lo), mmx_usable<T, CHK>(v, abad)); }
    MmxAcc<T> acc = loc.finish(lo);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { const MmxAcc<T> o = mmx_shfl_down(acc, d); acc.merge(o); }
    if (lane == 0) {
      if (p.nchunks == 1) { int64_t oo[4]; mmx_offsets(p, row, oa, oo); mmx_write<T>(p, oo, acc, flagged); }
      else reinterpret_cast<MmxAcc<T> *>(p.part)[w] = acc;
    }
  }
  if (flagged && p.flag) *reinterpret_cast<volatile int *>(p.flag) = 1;      // lane 0 only ever sets it: one store per warp
}

// finishing pass: one CTA of 1024 threads per row; threads stride over the chunk partials (two independent merges in
// flight each: a flat 2^28-element ndarray leaves ~9400 partials, and 37 dependent 32-byte loads per thread of a
// 256-thread CTA cost more than 20 us), then warp shuffles + shared memory
template <class T>
__global__ void __launch_bounds__(1024) minmaximum_finish_kernel(const __grid_constant__ MmxPlan p) {
  __shared__ MmxAcc<T> sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bool flagged = false;
  for (int64_t row = blockIdx.x; row < p.nrows; row += gridDim.x) {
    const MmxAcc<T> *part = reinterpret_cast<const MmxAcc<T> *>(p.part) + row * p.nchunks;
    MmxAcc<T> acc, acc2; acc.init(); acc2.init();
    int64_t k = threadIdx.x;
    for (; k + 1024 < p.nchunks; k += 2048) {           // ascending chunk order within each accumulator
      const MmxAcc<T> x = part[k], y = part[k + 1024];
      acc.merge(x); acc2.merge(y);
    }
    if (k < p.nchunks) acc.merge(part[k]);
    acc.merge(acc2);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { const MmxAcc<T> o = mmx_shfl_down(acc, d); acc.merge(o); }
    if (lane == 0) sh[wid] = acc;
    __syncthreads();
    if (wid == 0) {
      acc = sh[lane];
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) { const MmxAcc<T> o = mmx_shfl_down(acc, d); acc.merge(o); }
      if (lane == 0) {
        int64_t oa, oo[4];
        mmx_offsets(p, row, oa, oo);
        mmx_write<T>(p, oo, acc, flagged);
      }
    }
    __syncthreads();
  }
  if (flagged && p.flag) *reinterpret_cast<volatile int *>(p.flag) = 1;
}

template <class T, bool CHK>
__global__ void __launch_bounds__(256) minmaximum_thread_kernel(const __grid_constant__ MmxPlan p) {
  const T abad = from_bits<T>(p.abad);
  bool flagged = false;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.nrows; row += (int64_t)gridDim.x * blockDim.x) {
    int64_t oa, oo[4];
    mmx_offsets(p, row, oa, oo);
    const T *pa = reinterpret_cast<const T *>(p.a) + oa;
    MmxAcc<T> acc; acc.init();
    for (int64_t n = 0; n < p.n; n++) { const T v = pa[n * p.inc_a]; if (mmx_usable<T, CHK>(v, abad)) acc.take(v, n); }
    mmx_write<T>(p, oo, acc, flagged);
  }
  mmx_publish_flag(p, flagged);
}

template <class T>
static int mmx_go(MmxPlan &p, cudaStream_t s, const Err &E) {
  const int64_t cap = (int64_t)sm_count() * 8;
  const bool per_thread = p.n < 64 && p.nrows >= 1024;
  const bool chk = p.badmode && !p.abadnan;
  if (!per_thread) {
    int64_t want = cap * 8 / (p.nrows > 0 ? p.nrows : 1);
    if (want < 1) want = 1;
    int64_t chunk = (p.n + want - 1) / want;
    chunk = (chunk + 1023) / 1024 * 1024;
    if (chunk < 1024) chunk = 1024;
    if (chunk > (1ll << 30)) chunk = 1ll << 30;      // lane-local indices are 32-bit offsets into the chunk
    p.chunk = chunk;
    p.nchunks = p.n > 0 ? (p.n + chunk - 1) / chunk : 1;
  }
  // one scratch block: [0,64) the flag word, then the partials (scratch() may move when it grows: ask once)
  if (p.want_flag || p.nchunks > 1) {
    char *base = (char *)scratch(64 + (p.nchunks > 1 ? (size_t)(p.nrows * p.nchunks) * sizeof(MmxAcc<T>) : 0), s);
    if (!base) return E.fail(PDLB200_ECUDA, "minmaximum: cannot allocate scratch");
    p.part = base + 64;
    if (p.want_flag && !p.flag) { p.flag = (int *)base; PDLB200_CUDA_OK(cudaMemsetAsync(p.flag, 0, sizeof(int), s), E); }
  }
  if (per_thread) {
    int64_t g = (p.nrows + 255) / 256;
    if (g > cap) g = cap;
    if (chk) minmaximum_thread_kernel<T, true><<<(int)g, 256, 0, s>>>(p);
    else minmaximum_thread_kernel<T, false><<<(int)g, 256, 0, s>>>(p);
  } else {
    int64_t g = (p.nrows * p.nchunks + 7) / 8;
    if (g > cap * 4) g = cap * 4;
    if (chk) minmaximum_warp_kernel<T, true><<<(int)g, 256, 0, s>>>(p);
    else minmaximum_warp_kernel<T, false><<<(int)g, 256, 0, s>>>(p);
    if (p.nchunks > 1) {
      int64_t g2 = p.nrows < cap * 4 ? p.nrows : cap * 4;
      minmaximum_finish_kernel<T><<<(int)g2, 1024, 0, s>>>(p);
      note_launch("minmaximum");
    }
  }
  note_launch("minmaximum");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

int launch_minmaximum(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 5) return E.fail(PDLB200_EINVAL, "minmaximum: expected 5 parameters");
  if (!t->anybad) return E.fail(PDLB200_EINVAL, "minmaximum: pdlb200_trans.anybad must point to an int32");
  if (t->pdls[3].type != PDLB200_IND || t->pdls[4].type != PDLB200_IND)
    return E.fail(PDLB200_EINVAL, "minmaximum: the index outputs are `indx`");
  const size_t sz = pdlb200_type_size(t->datatype);
  if (!sz) return E.fail(PDLB200_EUNSUPPORTED, "minmaximum: type %d is not on the device path", t->datatype);
  *t->anybad = 0;
  Collapsed c;
  collapse_dims(t, &c);
  if (c.total == 0) return PDLB200_OK;
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "minmaximum: %d non-mergeable broadcast dims exceed the device walker's %d", c.nd, MAXD);
  MmxPlan p{};
  p.n = t->ind[0]; p.inc_a = t->rinc[0];
  if (p.n < 0) return E.fail(PDLB200_EINVAL, "minmaximum: n = %lld", (long long)p.n);
  if (!t->pdls[0].data && p.n > 0) return E.fail(PDLB200_EINVAL, "minmaximum: parameter 0 got NULL data");
  p.a = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)sz;
  for (int k = 0; k < 4; k++) {
    const pdlb200_par &o = t->pdls[k + 1];
    if (!o.data) return E.fail(PDLB200_EINVAL, "minmaximum: parameter %d got NULL data", k + 1);
    p.o[k] = (char *)o.data + o.offs * (int64_t)(k < 2 ? sz : 8);
    p.obad[k] = o.badval;
    for (int d = 0; d < c.nd; d++) p.so[k][d] = c.st[k + 1][d];
  }
  p.nd = c.nd; p.nrows = c.total; p.nchunks = 1;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sa[d] = c.st[0][d]; }
  p.abad = t->pdls[0].badval;
  p.badmode = t->bvalflag != 0;
  p.abadnan = (t->pdls[0].flags & PDLB200_PAR_BADNAN) != 0;
  cudaStream_t s = (cudaStream_t)t->stream;
  // "no usable element" can only happen with BAD inputs, NaNs (floating point) or n == 0
  const bool may_flag = p.badmode || p.n == 0 || t->datatype >= PDLB200_F;
  p.want_flag = may_flag;
  // deferred mode: the kernels store straight into the caller's pinned int32 (already 0) — no memset, no copy, no sync
  const bool defer = may_flag && (t->tflags & PDLB200_TRANS_DEFER_ANYBAD);
  if (defer) p.flag = reinterpret_cast<int *>(t->anybad);
  int rc;
  switch (t->datatype) {
    case PDLB200_SB: rc = mmx_go<int8_t>(p, s, E); break;   case PDLB200_B:  rc = mmx_go<uint8_t>(p, s, E); break;
    case PDLB200_S:  rc = mmx_go<int16_t>(p, s, E); break;  case PDLB200_US: rc = mmx_go<uint16_t>(p, s, E); break;
    case PDLB200_L:  rc = mmx_go<int32_t>(p, s, E); break;  case PDLB200_UL: rc = mmx_go<uint32_t>(p, s, E); break;
    case PDLB200_IND: case PDLB200_LL: rc = mmx_go<int64_t>(p, s, E); break;
    case PDLB200_ULL: rc = mmx_go<uint64_t>(p, s, E); break;
    case PDLB200_F:  rc = mmx_go<float>(p, s, E); break;    case PDLB200_D:  rc = mmx_go<double>(p, s, E); break;
    default: return E.fail(PDLB200_EUNSUPPORTED, "minmaximum: type %d is not on the device path", t->datatype);
  }
  if (rc) return rc;
  if (may_flag) {
    if (defer) return PDLB200_OK;
    int host = 0;
    PDLB200_CUDA_OK(cudaMemcpyAsync(&host, scratch(64, s), sizeof(int), cudaMemcpyDeviceToHost, s), E);
    PDLB200_CUDA_OK(cudaStreamSynchronize(s), E);
    *t->anybad = host != 0;
  }
  return PDLB200_OK;
}
}  // namespace pdlb200
