// ew_cmp.cu — gt lt le ge eq ne (lib/PDL/Ops.pd:297-302): result 0/1 in the op type.
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {
#define OP_SWITCH(OP, NAME) switch (t->datatype) { \
  PDLB200_EW_CASES_INT(OP, 2, true, NAME) PDLB200_EW_CASES_FLT(OP, 2, true, NAME) default: break; } break;
int ew_cmp(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_GT: OP_SWITCH(OpGt, "ew_gt")
    case PDLB200_OP_LT: OP_SWITCH(OpLt, "ew_lt")
    case PDLB200_OP_LE: OP_SWITCH(OpLe, "ew_le")
    case PDLB200_OP_GE: OP_SWITCH(OpGe, "ew_ge")
    case PDLB200_OP_EQ: OP_SWITCH(OpEq, "ew_eq")
    case PDLB200_OP_NE: OP_SWITCH(OpNe, "ew_ne")
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
