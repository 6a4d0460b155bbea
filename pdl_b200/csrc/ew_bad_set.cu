// ew_bad_set.cu — setbadif, setvaltobad, setbadtoval (lib/PDL/Bad.pd:584-677,808-840), all types.
#include "ew_badops.cuh"
namespace pdlb200 {
int ew_bad_set(const pdlb200_trans *t, const Err &E) {
  switch (t->op) {
    case PDLB200_OP_SETBADIF:
      if (t->npdls != 3 || t->pdls[1].type != PDLB200_L)
        return E.fail(PDLB200_EINVAL, "setbadif: the mask parameter is `int` (long), got type %d", t->pdls[1].type);
#define Q(T) return ew_launch_typed<OpSetbadif, T, T, 2, int32_t>(t, false, "ew_setbadif", E);
      switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
#define Q(T) return ew_launch_typed<OpSetvaltobad, T, T, 1>(t, false, "ew_setvaltobad", E, cast_bits<T>(t->param));
    case PDLB200_OP_SETVALTOBAD: switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
#define Q(T) return ew_launch_typed<OpSetbadtoval, T, T, 1>(t, false, "ew_setbadtoval", E, cast_bits<T>(t->param));
    case PDLB200_OP_SETBADTOVAL: switch (t->datatype) { PDLB200_BAD_CASES(Q) default: break; } break;
#undef Q
    default: break;
  }
  return E.fail(PDLB200_EUNSUPPORTED, "%s: type %d is not on the device path", pdlb200_op_name(t->op), t->datatype);
}
}  // namespace pdlb200
