// reduce_sum.cu — sumover dsumover (lib/PDL/Ufunc.pd:88-118).  Output type: int+ = max(long, T) or double.
#include "reduce_dispatch.cuh"
namespace pdlb200 {
int reduce_sum_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_SUMOVER:   switch (t->datatype) { RD_CASES(RSum,  OUT_PLUS, "reduce_sumover")   default: break; } break;
    case PDLB200_OP_DSUMOVER:  switch (t->datatype) { RD_CASES(RSum,  OUT_DBL,  "reduce_dsumover")  default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
