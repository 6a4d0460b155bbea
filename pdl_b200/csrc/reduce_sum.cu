// reduce_sum.cu — sumover prodover dsumover dprodover (lib/PDL/Ufunc.pd:88-118),
// average daverage (:413-444).  Output type: int+ = max(long, T) or double.
#include "reduce.cuh"
namespace pdlb200 {
#define RD_CASES(RED, OUTT, NAME) \
  case PDLB200_SB:  return rd_launch_typed<RED<int8_t,   OUTT(int8_t)>,   int8_t,   OUTT(int8_t)>(t, NAME, E); \
  case PDLB200_B:   return rd_launch_typed<RED<uint8_t,  OUTT(uint8_t)>,  uint8_t,  OUTT(uint8_t)>(t, NAME, E); \
  case PDLB200_S:   return rd_launch_typed<RED<int16_t,  OUTT(int16_t)>,  int16_t,  OUTT(int16_t)>(t, NAME, E); \
  case PDLB200_US:  return rd_launch_typed<RED<uint16_t, OUTT(uint16_t)>, uint16_t, OUTT(uint16_t)>(t, NAME, E); \
  case PDLB200_L:   return rd_launch_typed<RED<int32_t,  OUTT(int32_t)>,  int32_t,  OUTT(int32_t)>(t, NAME, E); \
  case PDLB200_UL:  return rd_launch_typed<RED<uint32_t, OUTT(uint32_t)>, uint32_t, OUTT(uint32_t)>(t, NAME, E); \
  case PDLB200_IND: case PDLB200_LL: return rd_launch_typed<RED<int64_t, OUTT(int64_t)>, int64_t, OUTT(int64_t)>(t, NAME, E); \
  case PDLB200_ULL: return rd_launch_typed<RED<uint64_t, OUTT(uint64_t)>, uint64_t, OUTT(uint64_t)>(t, NAME, E); \
  case PDLB200_F:   return rd_launch_typed<RED<float,    OUTT(float)>,    float,    OUTT(float)>(t, NAME, E); \
  case PDLB200_D:   return rd_launch_typed<RED<double,   OUTT(double)>,   double,   OUTT(double)>(t, NAME, E);
#define OUT_PLUS(T) typename tt<T>::plus
#define OUT_DBL(T) double
int reduce_sum_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_SUMOVER:   switch (t->datatype) { RD_CASES(RSum,  OUT_PLUS, "reduce_sumover")   default: break; } break;
    case PDLB200_OP_PRODOVER:  switch (t->datatype) { RD_CASES(RProd, OUT_PLUS, "reduce_prodover")  default: break; } break;
    case PDLB200_OP_DSUMOVER:  switch (t->datatype) { RD_CASES(RSum,  OUT_DBL,  "reduce_dsumover")  default: break; } break;
    case PDLB200_OP_DPRODOVER: switch (t->datatype) { RD_CASES(RProd, OUT_DBL,  "reduce_dprodover") default: break; } break;
    case PDLB200_OP_AVERAGE:   switch (t->datatype) { RD_CASES(RAvg,  OUT_PLUS, "reduce_average")   default: break; } break;
    case PDLB200_OP_DAVERAGE:  switch (t->datatype) { RD_CASES(RAvg,  OUT_DBL,  "reduce_daverage")  default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
