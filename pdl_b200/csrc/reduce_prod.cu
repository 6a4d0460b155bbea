// reduce_prod.cu — prodover dprodover (lib/PDL/Ufunc.pd:88-118), incl. the `tmp == 0 -> break` early exit.
#include "reduce_dispatch.cuh"
namespace pdlb200 {
int reduce_prod_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_PRODOVER:  switch (t->datatype) { RD_CASES(RProd, OUT_PLUS, "reduce_prodover")  default: break; } break;
    case PDLB200_OP_DPRODOVER: switch (t->datatype) { RD_CASES(RProd, OUT_DBL,  "reduce_dprodover") default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
