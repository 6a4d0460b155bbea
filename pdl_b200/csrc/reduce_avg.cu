// reduce_avg.cu — average daverage (lib/PDL/Ufunc.pd:413-444).
#include "reduce_dispatch.cuh"
namespace pdlb200 {
int reduce_avg_family(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_AVERAGE:   switch (t->datatype) { RD_CASES(RAvg,  OUT_PLUS, "reduce_average")   default: break; } break;
    case PDLB200_OP_DAVERAGE:  switch (t->datatype) { RD_CASES(RAvg,  OUT_DBL,  "reduce_daverage")  default: break; } break;
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
