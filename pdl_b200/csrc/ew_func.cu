// ew_func.cu — power atan2 modulo spaceship (lib/PDL/Ops.pd:321-324).  bifunc's BAD
// test has no per-parameter state check (Ops.pd:210), hence state_checked_bad=false.
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {
#define OP_SWITCH_ALL(OP, NAME) switch (t->datatype) { \
  PDLB200_EW_CASES_INT(OP, 2, false, NAME) PDLB200_EW_CASES_FLT(OP, 2, false, NAME) default: break; } break;
#define OP_SWITCH_FLT(OP, NAME) switch (t->datatype) { PDLB200_EW_CASES_FLT(OP, 2, false, NAME) default: break; } break;
int ew_func(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_POWER:     OP_SWITCH_FLT(OpPower, "ew_power")
    case PDLB200_OP_ATAN2:     OP_SWITCH_FLT(OpAtan2, "ew_atan2")
    case PDLB200_OP_MODULO:    OP_SWITCH_ALL(OpModulo, "ew_modulo")
    case PDLB200_OP_SPACESHIP: OP_SWITCH_ALL(OpSpaceship, "ew_spaceship")
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
