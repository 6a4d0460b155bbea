// ew_unary.cu — sqrt sin cos not exp log log10 (lib/PDL/Ops.pd:328-333,365-380),
// _rabs (:345-361), assgn (:382-397), abs2 (:491-503): a(); [o]b().
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {
#define OP_SWITCH_ALL(OP, NAME) switch (t->datatype) { \
  PDLB200_EW_CASES_INT(OP, 1, false, NAME) PDLB200_EW_CASES_FLT(OP, 1, false, NAME) default: break; } break;
#define OP_SWITCH_FLT(OP, NAME) switch (t->datatype) { PDLB200_EW_CASES_FLT(OP, 1, false, NAME) default: break; } break;
int ew_unary(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_SQRT:  OP_SWITCH_ALL(OpSqrt,  "ew_sqrt")
    case PDLB200_OP_SIN:   OP_SWITCH_ALL(OpSin,   "ew_sin")
    case PDLB200_OP_COS:   OP_SWITCH_ALL(OpCos,   "ew_cos")
    case PDLB200_OP_NOT:   OP_SWITCH_ALL(OpNot,   "ew_not")
    case PDLB200_OP_EXP:   OP_SWITCH_FLT(OpExp,   "ew_exp")
    case PDLB200_OP_LOG:   OP_SWITCH_FLT(OpLog,   "ew_log")
    case PDLB200_OP_LOG10: OP_SWITCH_ALL(OpLog10, "ew_log10")
    case PDLB200_OP_RABS:  OP_SWITCH_ALL(OpRabs,  "ew_rabs")
    case PDLB200_OP_ASSGN: OP_SWITCH_ALL(OpAssgn, "ew_assgn")
    case PDLB200_OP_ABS2:  OP_SWITCH_ALL(OpAbs2,  "ew_abs2")
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
