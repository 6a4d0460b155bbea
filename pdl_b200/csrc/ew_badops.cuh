// ew_badops.cuh — per-element bodies of lib/PDL/Bad.pd:343-416,584-905: the
// bad-value producers and consumers that sit either side of the hot path in every script
// (setbadif -> sumover, average -> setbadtoval ...).  These functors use the bad-aware protocol of
// elementwise.cuh: the kernel hands them the PDL_ISBAD2 test results instead of applying the
// "any BAD in -> BAD out" rule itself.  All of them are HBM-bound copies with a select.
#pragma once
#include <cstring>
#include "elementwise.cuh"

namespace pdlb200 {

#define PDLB200_BADOP template <class TI, class TB, class TO> static __device__ __forceinline__ TO \
  g(TI a, bool abad, TB b, bool bbad, TO cbad, uint64_t param, int &flag)

template <class T> __device__ __forceinline__ bool t_isfinite(T v) {
  if constexpr (tt<T>::is_int) return true; else return isfinite(v);
}

// isbad / isgood, Bad.pd:343-371,387-416: $b() = PDL_IF_BAD($ISBAD(a()),0) / PDL_IF_BAD($ISGOOD(a()),1)
struct OpIsbad  { static constexpr bool kBadAware = true; PDLB200_BADOP { return (TO)(abad ? 1 : 0); } };
struct OpIsgood { static constexpr bool kBadAware = true; PDLB200_BADOP { return (TO)(abad ? 0 : 1); } };
// isnan, Bad.pd:373-385: no HandleBad, the same body in both modes
struct OpIsnan  { static constexpr bool kBadAware = true; PDLB200_BADOP { return (TO)(t_isnan(a) ? 1 : 0); } };
// setbadif, Bad.pd:584-637: if (ISBAD(mask) || mask) SETBAD(b) else b = a      (a BAD is copied as it is)
struct OpSetbadif { static constexpr bool kBadAware = true; PDLB200_BADOP { return (bbad || b != TB(0)) ? cbad : (TO)a; } };
// setvaltobad, Bad.pd:639-677: a == (T)value -> BAD; param carries (T)value
struct OpSetvaltobad { static constexpr bool kBadAware = true; PDLB200_BADOP { return (a == from_bits<TI>(param)) ? cbad : (TO)a; } };
// setnantobad / setinftobad / setnonfinitetobad, Bad.pd:679-775: flag only when a BAD was written
struct OpSetnantobad { static constexpr bool kBadAware = true; PDLB200_BADOP {
  if (t_isnan(a)) { flag = 1; return cbad; } return (TO)a; } };
struct OpSetinftobad { static constexpr bool kBadAware = true; PDLB200_BADOP {
  if (!t_isfinite(a) && !t_isnan(a)) { flag = 1; return cbad; } return (TO)a; } };
struct OpSetnonfinitetobad { static constexpr bool kBadAware = true; PDLB200_BADOP {
  if (!t_isfinite(a)) { flag = 1; return cbad; } return (TO)a; } };
// setbadtonan, Bad.pd:777-806: $ISBAD(a()) is tested in BOTH code copies (no PDL_IF_BAD), so the launcher
// always runs the BAD-testing kernel; NAN is C's positive quiet NaN
struct OpSetbadtonan { static constexpr bool kBadAware = true; PDLB200_BADOP {
  if constexpr (sizeof(TO) == 4) return abad ? __uint_as_float(0x7fc00000u) : (TO)a;
  else return abad ? __longlong_as_double(0x7ff8000000000000ll) : (TO)a; } };
// setbadtoval, Bad.pd:808-840: param carries (T)newval
struct OpSetbadtoval { static constexpr bool kBadAware = true; PDLB200_BADOP { return abad ? from_bits<TO>(param) : (TO)a; } };
// badmask, Bad.pd:842-860: c = (isfinite((double)a) && ISGOOD(a)) ? a : b
struct OpBadmask { static constexpr bool kBadAware = true; PDLB200_BADOP { return (t_isfinite(a) && !abad) ? (TO)a : (TO)b; } };
// copybad, Bad.pd:862-905: ISBAD(mask) -> BAD else b = a
struct OpCopybad { static constexpr bool kBadAware = true; PDLB200_BADOP { return bbad ? cbad : (TO)a; } };

// (T)value exactly as the reference's C cast does it (host compiler, same ISA)
template <class T> static inline uint64_t cast_bits(double v) {
  T x = (T)v; uint64_t b = 0; memcpy(&b, &x, sizeof x); return b;
}

#define PDLB200_BAD_CASES(CALL) \
  case PDLB200_SB:  CALL(int8_t)  case PDLB200_B:   CALL(uint8_t) case PDLB200_S:  CALL(int16_t) \
  case PDLB200_US:  CALL(uint16_t) case PDLB200_L:  CALL(int32_t) case PDLB200_UL: CALL(uint32_t) \
  case PDLB200_IND: case PDLB200_LL: CALL(int64_t) case PDLB200_ULL: CALL(uint64_t) \
  case PDLB200_F:   CALL(float)   case PDLB200_D:   CALL(double)

}  // namespace pdlb200
