// nind.cu — minimum_n_ind / maximum_n_ind (lib/PDL/Ufunc.pd:502-561): a(n); indx [o]c(m).
//
// c(0..m-1) = the indices of the first m extreme elements of the row, found by m selection passes: pass k takes
// the extreme of the elements not selected by passes 0..k-1, with minimum_ind's rule (strict compare keeps the
// FIRST of equal values, a NaN `cur` is always replaced, BAD elements never qualify).  When fewer than m good
// elements exist the remaining slots are BAD.  The output's badflag follows the reference's generated code to the
// letter: `$PDLSTATESETGOOD(c)` sits at the top of a Code without an explicit broadcastloop, so it runs once per
// broadcast position and `$PDLSTATESETBAD(c)` of an earlier row is undone by a later one — the flag (`anybad`)
// says whether the LAST row (in broadcast order) had a slot it could not fill (Ufunc.pd:521-533).
// One CTA per row; every pass is a CTA-wide reduction of (value, first index, state) with the exact reducer of
// reduce.cuh, so each pass equals the reference's sequential loop however the row is cut.  Already-selected
// elements are marked in a shared-memory bitmap (rows up to 2^18 elements; longer rows scan the output list).
// HBM-bound for m = 1, L2-resident re-reads for the later passes: m * n * sizeof(T) bytes read per row.
#include <cstring>
#include "reduce.cuh"
namespace pdlb200 {

constexpr int NI_BITMAP_WORDS = 8192;            // 32 KB: one bit per element for n <= 262144

struct NiPlan {
  const char *a; int64_t *c;
  int64_t n, m, inc_n, inc_m, nrows;
  int64_t dims[MAXD], sa[MAXD], sc[MAXD];
  uint64_t abad, cbad;
  int *flag;
  int nd, abadnan, badmode;
};

template <class T, bool ISMAX>
__global__ void __launch_bounds__(RD_THREADS) nind_kernel(const __grid_constant__ NiPlan p) {
  using R = RMinMaxExact<T, int64_t, ISMAX, true>;
  using Acc = typename R::Acc;
  __shared__ Acc smem[RD_THREADS / 32 + 1];
  __shared__ uint32_t taken[NI_BITMAP_WORDS];
  const bool use_bitmap = p.n <= (int64_t)NI_BITMAP_WORDS * 32;
  const T abad = from_bits<T>(p.abad);
  for (int64_t row = blockIdx.x; row < p.nrows; row += gridDim.x) {
    int64_t oa = 0, oc = 0, rem = row;
    for (int d = 0; d < p.nd; d++) {
      const int64_t i = (d == p.nd - 1) ? rem : rem % p.dims[d];
      rem = (d == p.nd - 1) ? 0 : rem / p.dims[d];
      oa += i * p.sa[d]; oc += i * p.sc[d];
    }
    const T *a = reinterpret_cast<const T *>(p.a) + oa;
    int64_t *c = p.c + oc;
    if (use_bitmap) for (int w = threadIdx.x; w < (int)((p.n + 31) / 32); w += RD_THREADS) taken[w] = 0;
    __syncthreads();
    for (int64_t k = 0; k < p.m; k++) {
      typename R::Loc loc = R::linit();
      int64_t base = 0;                            // Loc indices are 32-bit: walk long rows in 2^30 pieces
      Acc mine = R::init();
      for (; base < p.n; base += (1ll << 30)) {
        const int64_t hi = (base + (1ll << 30) < p.n) ? base + (1ll << 30) : p.n;
        loc = R::linit();
        for (int64_t i = base + threadIdx.x; i < hi; i += RD_THREADS) {
          const T v = a[i * p.inc_n];
          if (p.badmode && is_bad(v, abad, p.abadnan != 0)) continue;
          bool sel = false;
          if (use_bitmap) sel = (taken[i >> 5] >> (i & 31)) & 1u;
          else for (int64_t q = 0; q < k && !sel; q++) sel = c[q * p.inc_m] == i;
          if (!sel) R::lpush(loc, v, (int32_t)(i - base));
        }
        mine = R::merge(mine, R::lift(loc, base));
      }
      const Acc tot = rd_group_reduce<R, 2>(mine, smem);
      if (threadIdx.x == 0) {
        if (tot.state == 0) { c[k * p.inc_m] = (int64_t)p.cbad; if (row == p.nrows - 1) *p.flag = 1; }
        else {
          c[k * p.inc_m] = tot.idx;
          if (use_bitmap) taken[tot.idx >> 5] |= 1u << (tot.idx & 31);
        }
      }
      __threadfence_block();
      __syncthreads();
    }
  }
}

template <class T>
static int ni_go(const pdlb200_trans *t, const NiPlan &p, const Err &E) {
  int64_t g = p.nrows;
  const int64_t cap = (int64_t)sm_count() * 4;
  if (g > cap) g = cap;
  if (t->op == PDLB200_OP_MAXIMUM_N_IND) nind_kernel<T, true><<<(int)g, RD_THREADS, 0, (cudaStream_t)t->stream>>>(p);
  else nind_kernel<T, false><<<(int)g, RD_THREADS, 0, (cudaStream_t)t->stream>>>(p);
  note_launch(t->op == PDLB200_OP_MAXIMUM_N_IND ? "maximum_n_ind" : "minimum_n_ind");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

int launch_nind(const pdlb200_trans *t, const Err &E) {
  const char *nm = pdlb200_op_name(t->op);
  if (t->npdls != 2) return E.fail(PDLB200_EINVAL, "%s: expected 2 parameters, got %d", nm, t->npdls);
  if (!t->anybad) return E.fail(PDLB200_EINVAL, "%s: the descriptor needs `anybad` (output badflag)", nm);
  *t->anybad = 0;
  if (t->pdls[1].type != PDLB200_IND && t->pdls[1].type != PDLB200_LL) return E.fail(PDLB200_EINVAL, "%s: output must be indx", nm);
  NiPlan p;
  memset(&p, 0, sizeof p);
  p.n = t->ind[0]; p.m = t->ind[1]; p.inc_n = t->rinc[0]; p.inc_m = t->rinc[1];
  if (p.n < 0 || p.m < 0) return E.fail(PDLB200_EINVAL, "%s: negative dim size", nm);
  if (p.m > p.n) return E.fail(PDLB200_EINVAL, "%s: m_size > n_size", nm);     // RedoDimsCode, Ufunc.pd:518
  Collapsed c;
  collapse_dims(t, &c);
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "%s: %d non-mergeable broadcast dims exceed the device walker's %d", nm, c.nd, MAXD);
  p.nrows = c.total; p.nd = c.nd;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sa[d] = c.st[0][d]; p.sc[d] = c.st[1][d]; }
  if (p.nrows == 0 || p.m == 0) return PDLB200_OK;
  if (!t->pdls[1].data || (p.n > 0 && !t->pdls[0].data)) return E.fail(PDLB200_EINVAL, "%s: parameter got NULL data", nm);
  const size_t sz = pdlb200_type_size(t->datatype);
  p.a = (const char *)t->pdls[0].data + t->pdls[0].offs * (int64_t)sz;
  p.c = (int64_t *)t->pdls[1].data + t->pdls[1].offs;
  p.abad = t->pdls[0].badval; p.cbad = t->pdls[1].badval;
  p.abadnan = (t->pdls[0].flags & PDLB200_PAR_BADNAN) != 0;
  p.badmode = t->bvalflag != 0;
  cudaStream_t s = (cudaStream_t)t->stream;
  p.flag = (int *)scratch(64, s);
  if (!p.flag) return E.fail(PDLB200_ECUDA, "%s: no scratch", nm);
  PDLB200_CUDA_OK(cudaMemsetAsync(p.flag, 0, sizeof(int), s), E);
  int rc;
  switch (t->datatype) {
    case PDLB200_SB: rc = ni_go<int8_t>(t, p, E); break;   case PDLB200_B:  rc = ni_go<uint8_t>(t, p, E); break;
    case PDLB200_S:  rc = ni_go<int16_t>(t, p, E); break;  case PDLB200_US: rc = ni_go<uint16_t>(t, p, E); break;
    case PDLB200_L:  rc = ni_go<int32_t>(t, p, E); break;  case PDLB200_UL: rc = ni_go<uint32_t>(t, p, E); break;
    case PDLB200_IND: case PDLB200_LL: rc = ni_go<int64_t>(t, p, E); break;
    case PDLB200_ULL: rc = ni_go<uint64_t>(t, p, E); break;
    case PDLB200_F:  rc = ni_go<float>(t, p, E); break;    case PDLB200_D:  rc = ni_go<double>(t, p, E); break;
    default: return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", nm, t->datatype);
  }
  if (rc) return rc;
  int flag = 0;
  PDLB200_CUDA_OK(cudaMemcpyAsync(&flag, p.flag, sizeof(int), cudaMemcpyDeviceToHost, s), E);
  PDLB200_CUDA_OK(cudaStreamSynchronize(s), E);
  *t->anybad = flag;
  return PDLB200_OK;
}

}  // namespace pdlb200
