// reduce_shard.cu — whole-array reductions of an ndarray partitioned across GPUs.
//
// The reference's whole-array wrappers are `$x->flat->Xover` (lib/PDL/Ufunc.pd:618-663); with the ndarray
// split along its outermost dim by the reference's own autopthread rule (lib/PDL/Core/pdlbroadcast.c:469-484)
// that reduction collapses the sharded dim.  Everything stays on the device:
//   PART_*  this rank's block -> one 32-byte record per row {value bits, good count, global index, state}
//           (the ordinary row reducers of reduce.cuh with a record-writing finish: same kernels, same
//           roofline — n*sizeof(T) bytes read per row);
//   gather  ncclAllGather of the records (32 B per row per rank), or peer stores;
//   COLL_*  the k records of a row merged in rank order with the reference's rules (Ufunc.pd:102-110 BAD
//           skipped / all BAD -> BAD, :427 average = sum / good count, :455-465 NaN loses to non-NaN, the
//           first index wins, all-NaN -> the last NaN), identical bits on every rank.
#include "reduce.cuh"
namespace pdlb200 {

template <class V> __device__ __forceinline__ uint64_t bits_of(V v) {
  if constexpr (sizeof(V) == 8) { union { V t; uint64_t u; } x; x.t = v; return x.u; }
  else if constexpr (sizeof(V) == 4) { union { V t; uint32_t u; } x; x.t = v; return x.u; }
  else if constexpr (sizeof(V) == 2) { union { V t; uint16_t u; } x; x.t = v; return x.u; }
  else { union { V t; uint8_t u; } x; x.t = v; return x.u; }
}

// sum + good count: RAvg's accumulators, a record instead of the quotient
template <class T, class O> struct RPartSum : RAvg<T, O> {
  using Acc = typename RAvg<T, O>::Acc;
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, int64_t *out) {
    out[0] = (int64_t)bits_of<O>(x.s); out[p.inc_r] = x.cnt; out[2 * p.inc_r] = 0; out[3 * p.inc_r] = x.cnt > 0;
  }
};
// extreme + first global index + state: RMinMax<.., WANT_IND>'s accumulators
template <class T, bool ISMAX> struct RPartMinMax : RMinMax<T, int64_t, ISMAX, true> {
  using Acc = typename RMinMax<T, int64_t, ISMAX, true>::Acc;
  static __device__ __forceinline__ void finish(const Acc &x, const RdPlan &p, int64_t *out) {
    const bool have = x.state != 0;
    out[0] = have ? (int64_t)bits_of<T>(x.cur) : 0; out[p.inc_r] = have;
    out[2 * p.inc_r] = have ? x.idx + p.goff : 0; out[3 * p.inc_r] = x.state;
  }
};

#define PS_CASES(OUTT, NAME) \
  case PDLB200_SB:  return rd_launch_typed<RPartSum<int8_t,   OUTT(int8_t)>,   int8_t,   int64_t>(t, NAME, E); \
  case PDLB200_B:   return rd_launch_typed<RPartSum<uint8_t,  OUTT(uint8_t)>,  uint8_t,  int64_t>(t, NAME, E); \
  case PDLB200_S:   return rd_launch_typed<RPartSum<int16_t,  OUTT(int16_t)>,  int16_t,  int64_t>(t, NAME, E); \
  case PDLB200_US:  return rd_launch_typed<RPartSum<uint16_t, OUTT(uint16_t)>, uint16_t, int64_t>(t, NAME, E); \
  case PDLB200_L:   return rd_launch_typed<RPartSum<int32_t,  OUTT(int32_t)>,  int32_t,  int64_t>(t, NAME, E); \
  case PDLB200_UL:  return rd_launch_typed<RPartSum<uint32_t, OUTT(uint32_t)>, uint32_t, int64_t>(t, NAME, E); \
  case PDLB200_IND: case PDLB200_LL: return rd_launch_typed<RPartSum<int64_t, OUTT(int64_t)>, int64_t, int64_t>(t, NAME, E); \
  case PDLB200_ULL: return rd_launch_typed<RPartSum<uint64_t, OUTT(uint64_t)>, uint64_t, int64_t>(t, NAME, E); \
  case PDLB200_F:   return rd_launch_typed<RPartSum<float,    OUTT(float)>,    float,    int64_t>(t, NAME, E); \
  case PDLB200_D:   return rd_launch_typed<RPartSum<double,   OUTT(double)>,   double,   int64_t>(t, NAME, E);
#define PM_CASES(ISMAX, NAME) \
  case PDLB200_SB:  return rd_launch_typed<RPartMinMax<int8_t,   ISMAX>, int8_t,   int64_t>(t, NAME, E); \
  case PDLB200_B:   return rd_launch_typed<RPartMinMax<uint8_t,  ISMAX>, uint8_t,  int64_t>(t, NAME, E); \
  case PDLB200_S:   return rd_launch_typed<RPartMinMax<int16_t,  ISMAX>, int16_t,  int64_t>(t, NAME, E); \
  case PDLB200_US:  return rd_launch_typed<RPartMinMax<uint16_t, ISMAX>, uint16_t, int64_t>(t, NAME, E); \
  case PDLB200_L:   return rd_launch_typed<RPartMinMax<int32_t,  ISMAX>, int32_t,  int64_t>(t, NAME, E); \
  case PDLB200_UL:  return rd_launch_typed<RPartMinMax<uint32_t, ISMAX>, uint32_t, int64_t>(t, NAME, E); \
  case PDLB200_IND: case PDLB200_LL: return rd_launch_typed<RPartMinMax<int64_t, ISMAX>, int64_t, int64_t>(t, NAME, E); \
  case PDLB200_ULL: return rd_launch_typed<RPartMinMax<uint64_t, ISMAX>, uint64_t, int64_t>(t, NAME, E); \
  case PDLB200_F:   return rd_launch_typed<RPartMinMax<float,    ISMAX>, float,    int64_t>(t, NAME, E); \
  case PDLB200_D:   return rd_launch_typed<RPartMinMax<double,   ISMAX>, double,   int64_t>(t, NAME, E);
#define OUT_PLUS(T) typename tt<T>::plus
#define OUT_DBL(T) double

int launch_partial(const pdlb200_trans *t, const Err &E) {
  if (t->pdls[1].type != PDLB200_LL && t->pdls[1].type != PDLB200_IND)
    return E.fail(PDLB200_EINVAL, "%s: the record parameter must be longlong", pdlb200_op_name(t->op));
  switch (t->op) {
    case PDLB200_OP_PART_SUM:  switch (t->datatype) { PS_CASES(OUT_PLUS, "part_sum")  default: break; } break;
    case PDLB200_OP_PART_DSUM: switch (t->datatype) { PS_CASES(OUT_DBL,  "part_dsum") default: break; } break;
    case PDLB200_OP_PART_MIN:  switch (t->datatype) { PM_CASES(false, "part_min") default: break; } break;
    case PDLB200_OP_PART_MAX:  switch (t->datatype) { PM_CASES(true,  "part_max") default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}

// ---- COLL_*: merge the k records of every row, one thread per row ---------------------------------------
struct CollPlan {
  const int64_t *rec; char *out;
  int64_t nrows, k, inc_r, inc_k;
  int64_t dims[MAXD], sr[MAXD], so[MAXD];
  uint64_t obad;
  int nd, kind, vtype, osize, badmode;
};

template <class V> __device__ __forceinline__ uint64_t coll_row(const CollPlan &p, const int64_t *rec) {
  const bool want_ind = p.kind == PDLB200_OP_COLL_MIN_IND || p.kind == PDLB200_OP_COLL_MAX_IND;
  if (p.kind == PDLB200_OP_COLL_SUM || p.kind == PDLB200_OP_COLL_AVG) {
    V tot = V(0); int64_t cnt = 0; bool any = false;
    for (int64_t r = 0; r < p.k; r++) {
      const int64_t *q = rec + r * p.inc_k;
      if (!q[3 * p.inc_r]) continue;
      tot = wrap_add<V>(tot, from_bits<V>((uint64_t)q[0])); cnt += q[p.inc_r]; any = true;
    }
    if (p.kind == PDLB200_OP_COLL_SUM) return (p.badmode && !any) ? p.obad : bits_of<V>(tot);
    if (cnt == 0) {
      if (p.badmode) return p.obad;
      if constexpr (tt<V>::is_int) return 0; else return bits_of<V>(nan_of<V>());
    }
    if constexpr (!tt<V>::is_int) return bits_of<V>(tot / (V)cnt);
    else if constexpr (sizeof(V) == 8 && tt<V>::is_uns) return (uint64_t)tot / (uint64_t)cnt;
    else return bits_of<V>((V)((int64_t)tot / cnt));
  }
  const bool ismax = p.kind == PDLB200_OP_COLL_MAX || p.kind == PDLB200_OP_COLL_MAX_IND;
  V cur = V(0); int64_t idx = -1, state = 0;
  for (int64_t r = 0; r < p.k; r++) {
    const int64_t *q = rec + r * p.inc_k;
    const int64_t st = q[3 * p.inc_r];
    if (!st) continue;
    const V v = from_bits<V>((uint64_t)q[0]); const int64_t i = q[2 * p.inc_r];
    bool take;
    if (!state) take = true;
    else if (state == 1 && st == 1) take = (ismax ? (v > cur) : (v < cur)) || (v == cur && i < idx);
    else if (state == 2 && st == 1) take = true;
    else if (state == 2 && st == 2) take = i > idx;      // every good value is NaN: the reference ends on the LAST one
    else take = false;
    if (take) { cur = v; idx = i; state = st; }
  }
  if (!state) return p.obad;
  return want_ind ? (uint64_t)idx : bits_of<V>(cur);
}

__global__ void __launch_bounds__(128) collapse_records_kernel(const __grid_constant__ CollPlan p) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.nrows; row += (int64_t)gridDim.x * blockDim.x) {
    int64_t orr = 0, oo = 0, rem = row;
    for (int d = 0; d < p.nd; d++) { const int64_t i = (d == p.nd - 1) ? rem : rem % p.dims[d]; rem /= p.dims[d]; orr += i * p.sr[d]; oo += i * p.so[d]; }
    const int64_t *rec = p.rec + orr;
    uint64_t bits;
    switch (p.vtype) {
      case PDLB200_SB: bits = coll_row<int8_t>(p, rec); break;   case PDLB200_B:  bits = coll_row<uint8_t>(p, rec); break;
      case PDLB200_S:  bits = coll_row<int16_t>(p, rec); break;  case PDLB200_US: bits = coll_row<uint16_t>(p, rec); break;
      case PDLB200_L:  bits = coll_row<int32_t>(p, rec); break;  case PDLB200_UL: bits = coll_row<uint32_t>(p, rec); break;
      case PDLB200_ULL: bits = coll_row<uint64_t>(p, rec); break;
      case PDLB200_F:  bits = coll_row<float>(p, rec); break;    case PDLB200_D:  bits = coll_row<double>(p, rec); break;
      default: bits = coll_row<int64_t>(p, rec); break;
    }
    char *o = p.out + oo * p.osize;
    if (p.osize == 8) *reinterpret_cast<uint64_t *>(o) = bits;
    else if (p.osize == 4) *reinterpret_cast<uint32_t *>(o) = (uint32_t)bits;
    else if (p.osize == 2) *reinterpret_cast<uint16_t *>(o) = (uint16_t)bits;
    else *reinterpret_cast<uint8_t *>(o) = (uint8_t)bits;
  }
}

int launch_collapse(const pdlb200_trans *t, const Err &E) {
  const char *nm = pdlb200_op_name(t->op);
  if (t->npdls != 2) return E.fail(PDLB200_EINVAL, "%s: expected 2 parameters, got %d", nm, t->npdls);
  if (t->pdls[0].type != PDLB200_LL && t->pdls[0].type != PDLB200_IND) return E.fail(PDLB200_EINVAL, "%s: the record parameter must be longlong", nm);
  if (t->ind[0] < 0) return E.fail(PDLB200_EINVAL, "%s: %lld ranks", nm, (long long)t->ind[0]);
  const bool want_ind = t->op == PDLB200_OP_COLL_MIN_IND || t->op == PDLB200_OP_COLL_MAX_IND;
  const int otype = t->pdls[1].type;
  if (want_ind ? (otype != PDLB200_IND && otype != PDLB200_LL) : (otype != t->datatype && !(otype == PDLB200_IND && t->datatype == PDLB200_LL) && !(otype == PDLB200_LL && t->datatype == PDLB200_IND)))
    return E.fail(PDLB200_EINVAL, "%s: output type %d does not match the record value type %d", nm, otype, t->datatype);
  Collapsed c;
  collapse_dims(t, &c);
  if (c.nd > MAXD) return E.fail(PDLB200_EUNSUPPORTED, "%s: %d non-mergeable broadcast dims exceed the device walker's %d", nm, c.nd, MAXD);
  if (c.total == 0) return PDLB200_OK;
  if (!t->pdls[1].data || (t->ind[0] > 0 && !t->pdls[0].data)) return E.fail(PDLB200_EINVAL, "%s: parameter got NULL data", nm);
  CollPlan p;
  memset(&p, 0, sizeof p);
  p.rec = (const int64_t *)t->pdls[0].data + t->pdls[0].offs;
  p.osize = (int)pdlb200_type_size(otype);
  p.out = (char *)t->pdls[1].data + t->pdls[1].offs * p.osize;
  p.nrows = c.total; p.k = t->ind[0]; p.inc_r = t->rinc[0]; p.inc_k = t->rinc[1];
  p.nd = c.nd;
  for (int d = 0; d < c.nd; d++) { p.dims[d] = c.dims[d]; p.sr[d] = c.st[0][d]; p.so[d] = c.st[1][d]; }
  p.obad = t->pdls[1].badval; p.kind = t->op; p.vtype = t->datatype; p.badmode = t->bvalflag != 0;
  int64_t g = (p.nrows + 127) / 128;
  if (g > (int64_t)sm_count() * 8) g = (int64_t)sm_count() * 8;
  collapse_records_kernel<<<(int)g, 128, 0, (cudaStream_t)t->stream>>>(p);
  note_launch("collapse_records");
  PDLB200_CUDA_OK(cudaGetLastError(), E);
  return PDLB200_OK;
}

}  // namespace pdlb200
