// reduce_bits.cu — andover orover zcover xorover (all types) and bandover borover
// bxorover (integer types), lib/PDL/Ufunc.pd:143-187.  Output type == input type.
#include "reduce.cuh"
namespace pdlb200 {
template <int KIND> struct BitsOf { template <class T, class O> using R = RBits<T, KIND>; };
#define BT_INT(KIND, NAME) \
  case PDLB200_SB:  return rd_launch_typed<RBits<int8_t,   KIND>, int8_t,   int8_t>(t, NAME, E); \
  case PDLB200_B:   return rd_launch_typed<RBits<uint8_t,  KIND>, uint8_t,  uint8_t>(t, NAME, E); \
  case PDLB200_S:   return rd_launch_typed<RBits<int16_t,  KIND>, int16_t,  int16_t>(t, NAME, E); \
  case PDLB200_US:  return rd_launch_typed<RBits<uint16_t, KIND>, uint16_t, uint16_t>(t, NAME, E); \
  case PDLB200_L:   return rd_launch_typed<RBits<int32_t,  KIND>, int32_t,  int32_t>(t, NAME, E); \
  case PDLB200_UL:  return rd_launch_typed<RBits<uint32_t, KIND>, uint32_t, uint32_t>(t, NAME, E); \
  case PDLB200_IND: case PDLB200_LL: return rd_launch_typed<RBits<int64_t, KIND>, int64_t, int64_t>(t, NAME, E); \
  case PDLB200_ULL: return rd_launch_typed<RBits<uint64_t, KIND>, uint64_t, uint64_t>(t, NAME, E);
#define BT_FLT(KIND, NAME) \
  case PDLB200_F:   return rd_launch_typed<RBits<float,  KIND>, float,  float>(t, NAME, E); \
  case PDLB200_D:   return rd_launch_typed<RBits<double, KIND>, double, double>(t, NAME, E);
int reduce_bits_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_ANDOVER:  switch (t->datatype) { BT_INT(0, "reduce_andover") BT_FLT(0, "reduce_andover") default: break; } break;
    case PDLB200_OP_OROVER:   switch (t->datatype) { BT_INT(1, "reduce_orover")  BT_FLT(1, "reduce_orover")  default: break; } break;
    case PDLB200_OP_ZCOVER:   switch (t->datatype) { BT_INT(2, "reduce_zcover")  BT_FLT(2, "reduce_zcover")  default: break; } break;
    case PDLB200_OP_XOROVER:  switch (t->datatype) { BT_INT(3, "reduce_xorover") BT_FLT(3, "reduce_xorover") default: break; } break;
    case PDLB200_OP_BANDOVER: switch (t->datatype) { BT_INT(4, "reduce_bandover") default: break; } break;
    case PDLB200_OP_BOROVER:  switch (t->datatype) { BT_INT(5, "reduce_borover")  default: break; } break;
    case PDLB200_OP_BXOROVER: switch (t->datatype) { BT_INT(6, "reduce_bxorover") default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
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
int reduce_minmax_family(const pdlb200_trans *, const Err &);
int launch_reduce(const pdlb200_trans *t, const Err &E) {
  if (t->op <= PDLB200_OP_DAVERAGE) return reduce_sum_family(t, E);
  if (t->op <= PDLB200_OP_MAXIMUM_IND) return reduce_minmax_family(t, E);
  if (t->op >= PDLB200_OP_NBADOVER) return reduce_count_family(t, E);
  return reduce_bits_family(t, E);
}
}  // namespace pdlb200
