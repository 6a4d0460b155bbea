// reduce_bitwise.cu — bandover borover bxorover (integer types), lib/PDL/Ufunc.pd:143-187.
#include "reduce.cuh"
namespace pdlb200 {
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
int reduce_bitwise_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_BANDOVER: switch (t->datatype) { BT_INT(4, "reduce_bandover") default: break; } break;
    case PDLB200_OP_BOROVER:  switch (t->datatype) { BT_INT(5, "reduce_borover")  default: break; } break;
    case PDLB200_OP_BXOROVER: switch (t->datatype) { BT_INT(6, "reduce_bxorover") default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
