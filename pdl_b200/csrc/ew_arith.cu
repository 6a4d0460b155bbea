// ew_arith.cu — plus mult minus divide (lib/PDL/Ops.pd:288-291), all real device types.
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {
#define OP_SWITCH(OP, NAME) switch (t->datatype) { \
  PDLB200_EW_CASES_INT(OP, 2, true, NAME) PDLB200_EW_CASES_FLT(OP, 2, true, NAME) default: break; } break;
int ew_arith(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_PLUS:   OP_SWITCH(OpPlus,   "ew_plus")
    case PDLB200_OP_MULT:   OP_SWITCH(OpMult,   "ew_mult")
    case PDLB200_OP_MINUS:  OP_SWITCH(OpMinus,  "ew_minus")
    case PDLB200_OP_DIVIDE: OP_SWITCH(OpDivide, "ew_divide")
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
