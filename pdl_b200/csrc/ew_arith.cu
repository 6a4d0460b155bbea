// ew_arith.cu — plus mult minus divide (lib/PDL/Ops.pd:288-291), all real device types.
#include "elementwise.cuh"
#include "ew_ops.cuh"
namespace pdlb200 {
#define OP_SWITCH(OP, NAME) switch (t->datatype) { \
  PDLB200_EW_CASES_INT(OP, 2, true, NAME) PDLB200_EW_CASES_FLT(OP, 2, true, NAME) default: break; } break;
// outer (lib/PDL/Primitive.pd:78-96): `loop(n,m) %{ c = a * b %}` is mult's body over two more dims, so the
// named dims n, m are put in front of the broadcast dims (a moves along n only, b along m only) and the mult
// kernels run it.  One difference kept: outer tests $ISBAD without looking at the ndarray's state flag.
static int ew_outer(const pdlb200_trans *t0, const Err &E) {
  if (t0->npdls != 3) return E.fail(PDLB200_EINVAL, "outer: expected 3 parameters");
  if (t0->ndims + 2 > PDLB200_MAXDIMS) return E.fail(PDLB200_EUNSUPPORTED, "outer: too many broadcast dims");
  pdlb200_trans w = *t0;
  w.ndims = t0->ndims + 2;
  w.dims[0] = t0->ind[0]; w.dims[1] = t0->ind[1];
  w.incs[0] = t0->rinc[0]; w.incs[1] = 0;           w.incs[2] = t0->rinc[2];
  w.incs[3] = 0;           w.incs[4] = t0->rinc[1]; w.incs[5] = t0->rinc[3];
  for (int d = 0; d < t0->ndims; d++) {
    w.dims[d + 2] = t0->dims[d];
    for (int p = 0; p < 3; p++) w.incs[(d + 2) * 3 + p] = t0->incs[d * 3 + p];
  }
  const pdlb200_trans *t = &w;
  switch (t->datatype) {
    PDLB200_EW_CASES_INT(OpMult, 2, false, "ew_outer") PDLB200_EW_CASES_FLT(OpMult, 2, false, "ew_outer") default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "outer: type %d is not on the device path", t->datatype);
}

int ew_arith(const pdlb200_trans *t, const Err &E) {
  if (t->op == PDLB200_OP_OUTER) return ew_outer(t, E);
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
