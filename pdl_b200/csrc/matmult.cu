// matmult.cu — PDL::Primitive::matmult, a(t,h); b(w,t); [o]c(w,h)  (lib/PDL/Primitive.pd:191-264).
//
// c(w,h) = sum_t a(t,h) * b(w,t): in memory order (dim 0 fastest) that is the row-major
// product C[h][w] = A[h][t] . B[t][w], each operand with arbitrary element strides
// (sliced / transposed / dummy views are read in place) and optional broadcast (batch) dims.
//
// Two kernels:
//  * mm_exact_kernel — every type, BAD-aware.  Shared-memory tiled SIMT; each output is
//    accumulated strictly left-to-right over t with separate multiply and add (-fmad=false),
//    i.e. the reference's own summation order (Primitive.pd:236-240), so results are bit-exact
//    for integer types AND for float/double.  Bound: FP64/FP32 pipe at 2 instructions per MAC.
//  * mm_dmma_kernel (matmult_dmma.cu) — double, no BAD: FP64 tensor-core tiles; fused multiply-add
//    inside the MMA, so equal only within the stated tolerance (exact for exactly-representable inputs).
#include <cstdlib>
#include <cstring>
#include "matmult.cuh"

namespace pdlb200 {


constexpr int MM_BM = 64, MM_BN = 64, MM_BK = 16, MM_TM = 4, MM_TN = 4;

template <class T> struct mm_acc { using type = T; };
template <> struct mm_acc<int8_t>   { using type = uint32_t; };
template <> struct mm_acc<uint8_t>  { using type = uint32_t; };
template <> struct mm_acc<int16_t>  { using type = uint32_t; };
template <> struct mm_acc<uint16_t> { using type = uint32_t; };
template <> struct mm_acc<int32_t>  { using type = uint32_t; };
template <> struct mm_acc<uint32_t> { using type = uint32_t; };
template <> struct mm_acc<int64_t>  { using type = uint64_t; };
template <> struct mm_acc<uint64_t> { using type = uint64_t; };

// One output element in the reference's own order with the x86 NaN rules at every step: the slow, exact twin of the
// hot loop, run only for outputs whose fast sum came out NaN.
template <class T>
__device__ __noinline__ T mm_nan_elem(const T *a, const T *b, int64_t tn, int64_t iat, int64_t ibt) {
  T cc = T(0);
  for (int64_t t = 0; t < tn; t++) {
    const T x = a[t * iat], y = b[t * ibt];
    const T m = x86_nan2(x, y, x * y);
    cc = x86_nan2(cc, m, cc + m);
  }
  return cc;
}

// One output element of a BAD-mode product exactly as the reference computes it (Primitive.pd:224-244), x86 NaN rules
// at every step: run only for outputs of the fast BAD path that ended up NaN (their value, not their state, is what
// the fast path's plain arithmetic cannot reproduce).
template <class T>
__device__ __noinline__ T mm_bad_elem(const T *a, const T *b, int64_t tn, int64_t iat, int64_t ibt, T abad, T bbad, T cbad,
                                      bool abadnan, bool bbadnan, bool cbadnan, int64_t tsiz) {
  T cc = T(0);
  for (int64_t ts = 0; ts < tn; ts += tsiz) {
    if (is_bad(cc, cbad, cbadnan)) continue;
    T run = cc;
    bool isbad = false;
    const int64_t te = ts + tsiz < tn ? ts + tsiz : tn;
    for (int64_t t = ts; t < te; t++) {
      const T x = a[t * iat], y = b[t * ibt];
      if (is_bad(x, abad, abadnan) || is_bad(y, bbad, bbadnan)) { isbad = true; break; }
      const T m = x86_nan2(x, y, x * y);
      run = x86_nan2(run, m, run + m);
    }
    cc = isbad ? cbad : run;
  }
  return cc;
}

// BAD mode (Primitive.pd:227,235,241), per output and per t-tile of TSIZ elements: at the tile start the running value
// is tested against c's badvalue — if it IS bad the output stops with that value (state 1); a BAD a(t,h) or b(w,t)
// inside the tile makes the output BAD (state 2); either way it stays stopped.  Round 2: the kernel no longer tests a
// flag per multiply-add.  Staging leaves ONE flag per (row, sub-tile) of a and per (column, 4 k) of b; a thread
// updates the states of its 4x4 outputs once per sub-tile (16-bit masks), and then runs the sub-tile's multiply-adds
// unconditionally (nothing stopped: the common case), not at all (everything stopped: the common case once BAD values
// are frequent — a CTA whose outputs have all stopped leaves the t loop), or per output (mixed).  Float/double run plain
// multiply + add here as in good mode; outputs that end up NaN are recomputed by mm_bad_elem.
template <class T, bool BAD>
__global__ void __launch_bounds__(256)
mm_exact_kernel(const __grid_constant__ MmPlan p) {
  using A = typename mm_acc<T>::type;  // wrap-around accumulator; low bits == the reference's `cc += a*b` in T
  constexpr int64_t TSIZ = 8 * sizeof(double) / sizeof(T);  // the reference's tile edge (Primitive.pd:212)
  constexpr int SUB = TSIZ < MM_BK ? (int)TSIZ : MM_BK;      // state updates per staged chunk: every SUB k
  constexpr int NSUB = MM_BK / SUB;
  __shared__ T sA[MM_BK][MM_BM + 1];
  __shared__ T sB[MM_BK][MM_BN + 1];
  __shared__ unsigned char rfA[BAD ? NSUB : 1][BAD ? MM_BM : 1];  // a BAD a(t,h) in sub-tile s of row h
  __shared__ unsigned char cfB[BAD ? 4 : 1][BAD ? MM_BN : 1];     // a BAD b(w,t) among k in [4q, 4q+4) of column w
  __shared__ int s_live[2];         // "some output of this CTA is still live", double-buffered by chunk parity

  int64_t oa = 0, ob = 0, oc = 0;
  {
    int64_t row = p.z0 + blockIdx.z;
    for (int d = 0; d < p.nd; d++) {
      const int64_t q = (d == p.nd - 1) ? 0 : row / p.dims[d];
      const int64_t i = row - q * p.dims[d];
      oa += i * p.sa[d]; ob += i * p.sb[d]; oc += i * p.sc[d];
      row = q;
    }
  }
  const T *Ap = reinterpret_cast<const T *>(p.a) + oa;
  const T *Bp = reinterpret_cast<const T *>(p.b) + ob;
  T *Cp = reinterpret_cast<T *>(p.c) + oc;
  const T abad = from_bits<T>(p.abad), bbad = from_bits<T>(p.bbad), cbad = from_bits<T>(p.cbad);

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  const int lane = threadIdx.x & 31;
  const int64_t h0 = (int64_t)blockIdx.y * MM_BM, w0 = (int64_t)blockIdx.x * MM_BN;

  A acc[MM_TM][MM_TN];
  unsigned fz1 = 0, fz2 = 0;           // bit i*4+j: output (i,j) stopped in state 1 / state 2
#pragma unroll
  for (int i = 0; i < MM_TM; i++)
#pragma unroll
    for (int j = 0; j < MM_TN; j++) acc[i][j] = A(0);

  if (BAD && threadIdx.x == 0) s_live[0] = 0;
  int chunk = 0;
  for (int64_t t0 = 0; t0 < p.T; t0 += MM_BK, chunk++) {
    // stage A[h0..h0+BM) x [t0..t0+BK)  and  B[t0..t0+BK) x [w0..w0+BN); threads run along t for A
    // and along w for B, the unit-stride dims of PDL's default layout
#pragma unroll
    for (int r = 0; r < MM_BM * MM_BK / 256; r++) {
      const int e = threadIdx.x + 256 * r;
      const int k = e % MM_BK, m = e / MM_BK;
      const int64_t h = h0 + m, tt_ = t0 + k;
      T v = T(0); bool f = false;
      if (h < p.H && tt_ < p.T) {
        v = Ap[tt_ * p.iat + h * p.iah];
        if (BAD) f = is_bad(v, abad, p.abadnan != 0);
      }
      sA[k][m] = v;
      if (BAD) {
        // 16 consecutive lanes hold the 16 k of one row: the row's flags come out of one ballot
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if ((k % SUB) == 0) rfA[k / SUB][m] = ((bal >> ((lane & 16) + k)) & ((1u << SUB) - 1u)) != 0;
      }
    }
    {
      const int n = threadIdx.x % MM_BN, q = threadIdx.x / MM_BN;      // this thread: column n, k in [4q, 4q+4)
      const int64_t w = w0 + n;
      bool f = false;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int k = 4 * q + r;
        const int64_t tt_ = t0 + k;
        T v = T(0);
        if (w < p.W && tt_ < p.T) {
          v = Bp[w * p.ibw + tt_ * p.ibt];
          if (BAD) f = f || is_bad(v, bbad, p.bbadnan != 0);
        }
        sB[k][n] = v;
      }
      if (BAD) cfB[q][n] = f;
    }
    __syncthreads();
    if (BAD && threadIdx.x == 0) s_live[(chunk + 1) & 1] = 0;    // read last after the previous chunk's second barrier
    const int kmax = (p.T - t0) < MM_BK ? (int)(p.T - t0) : MM_BK;
    if constexpr (BAD) {
#pragma unroll
      for (int sub = 0; sub < NSUB; sub++) {
        if (sub * SUB >= kmax) break;
        unsigned fz = fz1 | fz2;
        if (fz != 0xffffu) {
          if (((t0 + sub * SUB) % TSIZ) == 0) {
            // the reference re-tests cc against c's badvalue at every t-tile start (Primitive.pd:227)
#pragma unroll
            for (int i = 0; i < MM_TM; i++)
#pragma unroll
              for (int j = 0; j < MM_TN; j++)
                if (!((fz >> (i * 4 + j)) & 1u) && is_bad((T)acc[i][j], cbad, p.cbadnan != 0)) fz1 |= 1u << (i * 4 + j);
          }
          unsigned rf = 0, cf = 0;
#pragma unroll
          for (int i = 0; i < MM_TM; i++) rf |= (unsigned)rfA[sub][ty * MM_TM + i] << i;
#pragma unroll
          for (int j = 0; j < MM_TN; j++) {
            unsigned c4 = 0;
#pragma unroll
            for (int q = sub * SUB / 4; q < (sub + 1) * SUB / 4; q++) c4 |= cfB[q][tx * MM_TN + j];
            cf |= c4 << j;
          }
          if (rf | cf) {
#pragma unroll
            for (int i = 0; i < MM_TM; i++)
#pragma unroll
              for (int j = 0; j < MM_TN; j++)
                if ((((rf >> i) | (cf >> j)) & 1u) && !((fz1 >> (i * 4 + j)) & 1u)) fz2 |= 1u << (i * 4 + j);
          }
          fz = fz1 | fz2;
        }
        if (fz == 0xffffu) continue;                       // every output of this thread has stopped
        const int kend = (kmax - sub * SUB) < SUB ? kmax - sub * SUB : SUB;
        if (fz == 0) {
#pragma unroll
          for (int kk = 0; kk < SUB; kk++) {
            if (kk < kend) {
              const int k = sub * SUB + kk;
              T av[MM_TM], bv[MM_TN];
#pragma unroll
              for (int i = 0; i < MM_TM; i++) av[i] = sA[k][ty * MM_TM + i];
#pragma unroll
              for (int j = 0; j < MM_TN; j++) bv[j] = sB[k][tx * MM_TN + j];
#pragma unroll
              for (int i = 0; i < MM_TM; i++)
#pragma unroll
                for (int j = 0; j < MM_TN; j++) {
                  if constexpr (tt<T>::is_int) acc[i][j] += (A)av[i] * (A)bv[j];
                  else acc[i][j] = acc[i][j] + av[i] * bv[j];
                }
            }
          }
        } else {
#pragma unroll
          for (int kk = 0; kk < SUB; kk++) {
            if (kk < kend) {
              const int k = sub * SUB + kk;
              T av[MM_TM], bv[MM_TN];
#pragma unroll
              for (int i = 0; i < MM_TM; i++) av[i] = sA[k][ty * MM_TM + i];
#pragma unroll
              for (int j = 0; j < MM_TN; j++) bv[j] = sB[k][tx * MM_TN + j];
#pragma unroll
              for (int i = 0; i < MM_TM; i++)
#pragma unroll
                for (int j = 0; j < MM_TN; j++) {
                  if ((fz >> (i * 4 + j)) & 1u) continue;
                  if constexpr (tt<T>::is_int) acc[i][j] += (A)av[i] * (A)bv[j];
                  else acc[i][j] = acc[i][j] + av[i] * bv[j];
                }
            }
          }
        }
      }
      // a CTA whose outputs have all stopped has nothing left to do
      if ((fz1 | fz2) != 0xffffu) s_live[chunk & 1] = 1;
      __syncthreads();
      if (!s_live[chunk & 1]) break;
    } else {
#pragma unroll
      for (int k = 0; k < MM_BK; k++) {
        if (k < kmax) {
          T av[MM_TM], bv[MM_TN];
#pragma unroll
          for (int i = 0; i < MM_TM; i++) av[i] = sA[k][ty * MM_TM + i];
#pragma unroll
          for (int j = 0; j < MM_TN; j++) bv[j] = sB[k][tx * MM_TN + j];
#pragma unroll
          for (int i = 0; i < MM_TM; i++)
#pragma unroll
            for (int j = 0; j < MM_TN; j++) {
              if constexpr (tt<T>::is_int) acc[i][j] += (A)av[i] * (A)bv[j];
              // good mode: plain multiply then add (two roundings, -fmad=false).  A NaN is sticky in a sum, so its
              // x86 sign/payload rules are applied afterwards, only to the outputs that ended up NaN (mm_nan_elem)
              else acc[i][j] = acc[i][j] + av[i] * bv[j];
            }
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < MM_TM; i++)
#pragma unroll
    for (int j = 0; j < MM_TN; j++) {
      const int64_t h = h0 + ty * MM_TM + i, w = w0 + tx * MM_TN + j;
      if (h < p.H && w < p.W) {
        T out = (T)acc[i][j];
        if (BAD && ((fz2 >> (i * 4 + j)) & 1u)) out = cbad;
        if constexpr (!tt<T>::is_int) {
          if (out != out) {
            if constexpr (BAD)
              out = mm_bad_elem<T>(Ap + h * p.iah, Bp + w * p.ibw, p.T, p.iat, p.ibt, abad, bbad, cbad, p.abadnan != 0,
                                   p.bbadnan != 0, p.cbadnan != 0, TSIZ);
            else out = mm_nan_elem<T>(Ap + h * p.iah, Bp + w * p.ibw, p.T, p.iat, p.ibt);
          }
        }
        Cp[w * p.icw + h * p.ich] = out;
      }
    }
}

template <class T>
static int mm_launch(const pdlb200_trans *t, const MmPlan &plan, const Err &E) {
  cudaStream_t s = (cudaStream_t)t->stream;
  MmPlan p = plan;
  // broadcast positions ride on gridDim.z (<= 65535): batches of many small matrices go out in slices
  for (p.z0 = 0; p.z0 < p.nbatch; p.z0 += MM_MAXZ) {
    const int64_t nz = (p.nbatch - p.z0 < MM_MAXZ) ? p.nbatch - p.z0 : MM_MAXZ;
    dim3 grid((unsigned)((p.W + MM_BN - 1) / MM_BN), (unsigned)((p.H + MM_BM - 1) / MM_BM), (unsigned)nz);
    if (t->bvalflag) mm_exact_kernel<T, true><<<grid, 256, 0, s>>>(p);
    else mm_exact_kernel<T, false><<<grid, 256, 0, s>>>(p);
    note_launch("matmult_exact");
    PDLB200_CUDA_OK(cudaGetLastError(), E);
  }
  return PDLB200_OK;
}

int launch_matmult(const pdlb200_trans *t, const Err &E) {
  if (t->npdls != 3) return E.fail(PDLB200_EINVAL, "matmult: expected 3 parameters, got %d", t->npdls);
  MmPlan p;
  memset(&p, 0, sizeof p);
  p.T = t->ind[0]; p.H = t->ind[1]; p.W = t->ind[2];
  if (p.T < 0 || p.H < 0 || p.W < 0) return E.fail(PDLB200_EINVAL, "matmult: negative dim size");
  p.iat = t->rinc[0]; p.iah = t->rinc[1]; p.ibw = t->rinc[2]; p.ibt = t->rinc[3]; p.icw = t->rinc[4]; p.ich = t->rinc[5];
  Collapsed c;
  collapse_dims(t, &c);
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "matmult: %d non-mergeable broadcast dims exceed %d", c.nd, MAXD);
  p.nd = c.nd; p.nbatch = c.total;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sa[d] = c.st[0][d]; p.sb[d] = c.st[1][d]; p.sc[d] = c.st[2][d]; }
  if (p.nbatch == 0 || p.H == 0 || p.W == 0) return PDLB200_OK;
  const size_t sz = pdlb200_type_size(t->datatype);
  for (int k = 0; k < 3; k++)
    if (!t->pdls[k].data) return E.fail(PDLB200_EINVAL, "matmult: parameter %d got NULL data", k);
  p.a = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)sz;
  p.b = (const char *)t->pdls[1].data + t->pdls[1].offs * (int64_t)sz;
  p.c = (char *)t->pdls[2].data + t->pdls[2].offs * (int64_t)sz;
  p.abad = t->pdls[0].badval; p.bbad = t->pdls[1].badval; p.cbad = t->pdls[2].badval;
  p.abadnan = (t->pdls[0].flags & PDLB200_PAR_BADNAN) != 0;
  p.bbadnan = (t->pdls[1].flags & PDLB200_PAR_BADNAN) != 0;
  p.cbadnan = (t->pdls[2].flags & PDLB200_PAR_BADNAN) != 0;

  const char *force = getenv("PDLB200_MATMULT");  // "exact" | "dmma" (default: dmma when eligible)
  const bool want_exact = force && !strcmp(force, "exact");
  if (t->datatype == PDLB200_D && !t->bvalflag && !want_exact) {
    int rc = launch_matmult_tma(t, p, E);         // TMA-staged tiles when the layout allows
    if (rc != PDLB200_EUNSUPPORTED) return rc;
    rc = launch_matmult_dmma(t, p, E);
    if (rc != PDLB200_EUNSUPPORTED) return rc;  // shape not eligible -> exact kernel
  }
  switch (t->datatype) {
    case PDLB200_SB:  return mm_launch<int8_t>(t, p, E);
    case PDLB200_B:   return mm_launch<uint8_t>(t, p, E);
    case PDLB200_S:   return mm_launch<int16_t>(t, p, E);
    case PDLB200_US:  return mm_launch<uint16_t>(t, p, E);
    case PDLB200_L:   return mm_launch<int32_t>(t, p, E);
    case PDLB200_UL:  return mm_launch<uint32_t>(t, p, E);
    case PDLB200_IND: case PDLB200_LL: return mm_launch<int64_t>(t, p, E);
    case PDLB200_ULL: return mm_launch<uint64_t>(t, p, E);
    case PDLB200_F:   return mm_launch<float>(t, p, E);
    case PDLB200_D:   return mm_launch<double>(t, p, E);
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "matmult: type %d is not on the device path", t->datatype);
}

}  // namespace pdlb200
