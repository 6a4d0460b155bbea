// reduce_minmax_ind.cu — minimum_ind maximum_ind (lib/PDL/Ufunc.pd:477-500).
#include <cstdlib>
#include "reduce.cuh"
namespace pdlb200 {
#define MM_CASES(ISMAX, WANT, NAME) \
  case PDLB200_SB:  return mm<int8_t,   ISMAX, WANT>(t, NAME, E); \
  case PDLB200_B:   return mm<uint8_t,  ISMAX, WANT>(t, NAME, E); \
  case PDLB200_S:   return mm<int16_t,  ISMAX, WANT>(t, NAME, E); \
  case PDLB200_US:  return mm<uint16_t, ISMAX, WANT>(t, NAME, E); \
  case PDLB200_L:   return mm<int32_t,  ISMAX, WANT>(t, NAME, E); \
  case PDLB200_UL:  return mm<uint32_t, ISMAX, WANT>(t, NAME, E); \
  case PDLB200_IND: case PDLB200_LL: return mm<int64_t, ISMAX, WANT>(t, NAME, E); \
  case PDLB200_ULL: return mm<uint64_t, ISMAX, WANT>(t, NAME, E); \
  case PDLB200_F:   return mm<float,    ISMAX, WANT>(t, NAME, E); \
  case PDLB200_D:   return mm<double,   ISMAX, WANT>(t, NAME, E);
template <class T, bool ISMAX, bool WANT>
static int mm(const pdlb200_trans *t, const char *name, const Err &E) {
  if constexpr (WANT && sizeof(T) <= 2) {
    // 8/16-bit rows: packed value reduction + first-index search (two phases) unless the rows are cut into chunks
    RdPlan p; RdLaunch l;
    const int rc = rd_build_plan(t, sizeof(T), sizeof(int64_t), sizeof(typename RMinMaxIntInd<T, ISMAX>::Acc), &p, &l, E);
    if (rc) return rc;
    if (p.nchunks == 1 && !getenv("PDLB200_IND_PER_ELEMENT"))
      return rd_launch_typed<RMinMaxIntInd<T, ISMAX>, T, int64_t>(t, name, E);
  }
  if constexpr (WANT) return rd_launch_typed<RMinMax<T, int64_t, ISMAX, true>, T, int64_t>(t, name, E);
  else if constexpr (tt<T>::is_int) return rd_launch_typed<RMinMaxInt<T, ISMAX>, T, T>(t, name, E);
  else return rd_launch_typed<RMinMax<T, T, ISMAX, false>, T, T>(t, name, E);
}
int reduce_minmax_ind_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_MINIMUM_IND: switch (t->datatype) { MM_CASES(false, true,  "reduce_minimum_ind") default: break; } break;
    case PDLB200_OP_MAXIMUM_IND: switch (t->datatype) { MM_CASES(true,  true,  "reduce_maximum_ind") default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
