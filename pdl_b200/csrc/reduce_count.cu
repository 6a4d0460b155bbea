// reduce_count.cu — nbadover ngoodover (lib/PDL/Bad.pd:418-480) and the dispatcher over the reduction families.
#include "reduce.cuh"
namespace pdlb200 {
#define CT_ALL(GOOD, NAME) \
  case PDLB200_SB:  return rd_launch_typed<RCount<int8_t,   GOOD>, int8_t,   int64_t>(t, NAME, E); \
  case PDLB200_B:   return rd_launch_typed<RCount<uint8_t,  GOOD>, uint8_t,  int64_t>(t, NAME, E); \
  case PDLB200_S:   return rd_launch_typed<RCount<int16_t,  GOOD>, int16_t,  int64_t>(t, NAME, E); \
  case PDLB200_US:  return rd_launch_typed<RCount<uint16_t, GOOD>, uint16_t, int64_t>(t, NAME, E); \
  case PDLB200_L:   return rd_launch_typed<RCount<int32_t,  GOOD>, int32_t,  int64_t>(t, NAME, E); \
  case PDLB200_UL:  return rd_launch_typed<RCount<uint32_t, GOOD>, uint32_t, int64_t>(t, NAME, E); \
  case PDLB200_IND: case PDLB200_LL: return rd_launch_typed<RCount<int64_t, GOOD>, int64_t, int64_t>(t, NAME, E); \
  case PDLB200_ULL: return rd_launch_typed<RCount<uint64_t, GOOD>, uint64_t, int64_t>(t, NAME, E); \
  case PDLB200_F:   return rd_launch_typed<RCount<float,    GOOD>, float,    int64_t>(t, NAME, E); \
  case PDLB200_D:   return rd_launch_typed<RCount<double,   GOOD>, double,   int64_t>(t, NAME, E);
int reduce_count_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_NBADOVER:  switch (t->datatype) { CT_ALL(false, "reduce_nbadover")  default: break; } break;
    case PDLB200_OP_NGOODOVER: switch (t->datatype) { CT_ALL(true,  "reduce_ngoodover") default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
int reduce_sum_family(const pdlb200_trans *, const Err &);
int reduce_prod_family(const pdlb200_trans *, const Err &);
int reduce_avg_family(const pdlb200_trans *, const Err &);
int reduce_minmax_family(const pdlb200_trans *, const Err &);
int reduce_minmax_ind_family(const pdlb200_trans *, const Err &);
int reduce_logic_family(const pdlb200_trans *, const Err &);
int reduce_bitwise_family(const pdlb200_trans *, const Err &);
int launch_reduce(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_SUMOVER: case PDLB200_OP_DSUMOVER: return reduce_sum_family(t, E);
    case PDLB200_OP_PRODOVER: case PDLB200_OP_DPRODOVER: return reduce_prod_family(t, E);
    case PDLB200_OP_AVERAGE: case PDLB200_OP_DAVERAGE: return reduce_avg_family(t, E);
    case PDLB200_OP_MINIMUM: case PDLB200_OP_MAXIMUM: return reduce_minmax_family(t, E);
    case PDLB200_OP_MINIMUM_IND: case PDLB200_OP_MAXIMUM_IND: return reduce_minmax_ind_family(t, E);
    case PDLB200_OP_ANDOVER: case PDLB200_OP_OROVER: case PDLB200_OP_ZCOVER: case PDLB200_OP_XOROVER: return reduce_logic_family(t, E);
    case PDLB200_OP_BANDOVER: case PDLB200_OP_BOROVER: case PDLB200_OP_BXOROVER: return reduce_bitwise_family(t, E);
    case PDLB200_OP_NBADOVER: case PDLB200_OP_NGOODOVER: return reduce_count_family(t, E);
    default: break;
  }
  return E.fail(PDLB200_EINVAL, "%s is not a reduction", pdlb200_op_name(t->op));
}
}  // namespace pdlb200
