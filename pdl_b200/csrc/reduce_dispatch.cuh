// reduce_dispatch.cuh — type-dispatch macros shared by the reduce_*.cu translation units.  The reduction
// families are split over several .cu files only to keep the per-file compile time (and the wall time of a
// from-scratch parallel build) down; every file instantiates reduce.cuh for its own ops.
#pragma once
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
}  // namespace pdlb200
