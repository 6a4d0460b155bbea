// ew_bits.cu — shiftleft shiftright or2 and2 xor (lib/PDL/Ops.pd:306-313) and
// bitnot (:327): integer types only (GenericTypes => $T).
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {
#define OP_SWITCH(OP, NIN, SC, NAME) switch (t->datatype) { PDLB200_EW_CASES_INT(OP, NIN, SC, NAME) default: break; } break;
int ew_bits(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_SHIFTLEFT:  OP_SWITCH(OpShl, 2, true, "ew_shiftleft")
    case PDLB200_OP_SHIFTRIGHT: OP_SWITCH(OpShr, 2, true, "ew_shiftright")
    case PDLB200_OP_OR2:        OP_SWITCH(OpOr,  2, true, "ew_or2")
    case PDLB200_OP_AND2:       OP_SWITCH(OpAnd, 2, true, "ew_and2")
    case PDLB200_OP_XOR:        OP_SWITCH(OpXor, 2, true, "ew_xor")
    case PDLB200_OP_BITNOT:     OP_SWITCH(OpBitnot, 1, false, "ew_bitnot")
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not an integer device type", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
