// ew_bad_query.cu — isbad / isgood / isnan (lib/PDL/Bad.pd:343-416): a(); int [o]b().
#include "ew_badops.cuh"
namespace pdlb200 {
int ew_bad_query(const pdlb200_trans *t, const Err &E) {
  if (t->pdls[1].type != PDLB200_L)
    return E.fail(PDLB200_EINVAL, "%s: the output parameter is `int` (long), got type %d", pdlb200_op_name(t->op), t->pdls[1].type);
  switch (t->op) {
#define Q(T) return ew_launch_typed<OpIsbad, T, int32_t, 1>(t, false, "ew_isbad", E);
    case PDLB200_OP_ISBAD: switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
#define Q(T) return ew_launch_typed<OpIsgood, T, int32_t, 1>(t, false, "ew_isgood", E);
    case PDLB200_OP_ISGOOD: switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
#define Q(T) return ew_launch_typed<OpIsnan, T, int32_t, 1>(t, false, "ew_isnan", E);
    case PDLB200_OP_ISNAN: switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
